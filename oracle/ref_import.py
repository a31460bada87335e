"""Import the UNMODIFIED reference (/root/reference) in this container, with import-time shims only.

TEST INFRASTRUCTURE: used by oracle/make_golden.py and by CPU tests that pin the oracle against the
live reference.  /root/reference does not exist on the GPU box; nothing in the product path, the
`-m gpu` tests, smoke() or bench.py imports this module.

Shims: torchvision.models.utils (removed upstream) -> torch.hub loader;
models.anchor_center / models.centerness (imported by A2/models/__init__.py but absent) -> empty stubs;
is_main_process -> False so the missing pretrained_models/resnet50-*.pth is not read.
"""
import importlib
import os
import sys
import types

REF_ROOT = "/root/reference/src"
STAGE_DIRS = {1: "CountDETR_lvis_1st_stage", 2: "CountDETR_147_2nd_stage"}


def available():
    return os.path.isdir(REF_ROOT)


def _purge():
    for k in list(sys.modules):
        if k == "models" or k.startswith("models.") or k == "util" or k.startswith("util."):
            del sys.modules[k]


def load(stage):
    """Returns the reference `models` package of the given stage (fresh import)."""
    import torch
    path = os.path.join(REF_ROOT, STAGE_DIRS[stage])
    _purge()
    sys.path = [p for p in sys.path if not p.startswith(REF_ROOT)]
    sys.path.insert(0, path)
    if "torchvision.models.utils" not in sys.modules:
        m = types.ModuleType("torchvision.models.utils")
        m.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = m
    if stage == 2:
        for name in ("models.anchor_center", "models.centerness"):
            stub = types.ModuleType(name)
            stub.build = lambda args: None
            stub.build_anchor_center = lambda args: None
            stub.build_centerness = lambda args: None
            sys.modules[name] = stub
    # skip pretrained download/load: patch before models.backbone binds it
    misc = importlib.import_module("util.misc")
    misc.is_main_process = lambda: False
    models = importlib.import_module("models")
    sys.modules["models.backbone"].is_main_process = lambda: False
    return models


def default_args(stage, **over):
    import argparse
    a = argparse.Namespace(
        device="cpu", backbone="resnet50", dilation=True, lr_backbone=1e-5, masks=False,
        num_feature_levels=1, hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6,
        dim_feedforward=1024, dropout=0.0, num_query_position=300, num_query_pattern=1,
        spatial_prior="learned", attention_type="RCDA", frozen_weights=None,
        aux_loss=False, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0,
        set_cost_class=2.0, set_cost_bbox=5.0, set_cost_giou=2.0,
        cls_loss_coef=2.0, bbox_loss_coef=5.0, giou_loss_coef=2.0, variance_loss_coef=2.0,
        focal_alpha=0.25, mask_loss_coef=1.0, dice_loss_coef=1.0, dataset_file="fsc147")
    for k, v in over.items():
        setattr(a, k, v)
    return a
