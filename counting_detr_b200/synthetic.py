"""Portable, seed-deterministic synthetic weights and inputs for benchmarks, smoke tests and parity tests.

There is no checkpoint and no dataset on the GPU box, and torch/numpy RNG streams are not a documented
cross-machine contract, so every synthetic tensor comes from a counter-based integer hash (splitmix64)
evaluated with numpy uint64 arithmetic: bit-identical on every machine.  The same function feeds the
reference model when golden vectors are generated (load_state_dict strict=True), the CPU oracle and
the CUDA path.

Key names and shapes follow the reference state_dict.  Distributions are chosen so that
activations stay O(1) through 16 un-normalised bottlenecks, the encoder keeps spatial variation, and the
heads the reference zero/constant-initialises (A2/models/transformer.py:94-103) are NOT degenerate
(otherwise box/variance parity would be vacuous).  Input laws: SURVEY.md §8d.
"""
import argparse
import zlib
from dataclasses import dataclass

import numpy as np
import torch

RESNET50_BLOCKS = (3, 4, 6, 3)
RESNET50_PLANES = (64, 128, 256, 512)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


@dataclass
class SynthCfg:
    stage: int = 2
    hidden_dim: int = 256
    nheads: int = 8
    enc_layers: int = 6
    dec_layers: int = 6
    dim_feedforward: int = 1024
    num_query_position: int = 300
    num_query_pattern: int = 1
    spatial_prior: str = "learned"
    aux_loss: bool = False

    @property
    def pattern_key(self):
        # A1/models/transformer.py:66 vs A2/models/transformer.py:67
        return "transformer.modify_pattern.weight" if self.stage == 1 else "transformer.pattern.weight"


def default_args(stage, **over):
    """The fields build_model(args) reads (A2/main.py:17-135, A1/main.py:27-132), reference defaults with
    --num_query_pattern 1 --no_aux_loss as in the reference's own scripts."""
    a = argparse.Namespace(
        device="cuda", backbone="resnet50", dilation=True, lr_backbone=1e-5, masks=False,
        num_feature_levels=1, hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6,
        dim_feedforward=1024, dropout=0.0, num_query_position=300, num_query_pattern=1,
        spatial_prior="learned", attention_type="RCDA", frozen_weights=None, dataset_file="fsc147")
    if stage == 2:
        a.__dict__.update(aux_loss=False, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, cls_loss_coef=2.0,
                          bbox_loss_coef=5.0, giou_loss_coef=2.0, variance_loss_coef=2.0, focal_alpha=0.25,
                          mask_loss_coef=1.0, dice_loss_coef=1.0)
    else:
        a.__dict__.update(aux_loss=False, set_cost_class=2.0, set_cost_bbox=5.0, set_cost_giou=2.0,
                          cls_loss_coef=2.0, bbox_loss_coef=5.0, giou_loss_coef=2.0, focal_alpha=0.25)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(name, shape, lo=0.0, hi=1.0, seed=0):
    """U[lo,hi) float32 tensor, a pure function of (name, seed, shape)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        key = np.uint64(zlib.crc32(name.encode()) & 0xFFFFFFFF) * np.uint64(0x100000001B3) + np.uint64(seed)
        base = _splitmix64(np.uint64(key))
        idx = np.arange(n, dtype=np.uint64)
        bits = _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + base) & _M64)
    u = (bits >> np.uint64(40)).astype(np.float64) / float(1 << 24)      # 24-bit mantissa, exact in fp32
    v = (lo + (hi - lo) * u).astype(np.float32)
    return torch.from_numpy(v.reshape(shape))


def _sym(name, shape, bound, seed):
    return uniform(name, shape, -bound, bound, seed)


def make_state_dict(cfg, seed=0):
    sd = {}
    E, F_ = cfg.hidden_dim, cfg.dim_feedforward

    def conv(name, cout, cin, k, gain=1.0):
        fan_in = cin * k * k
        sd[name] = _sym(name, (cout, cin, k, k), gain * (3.0 / fan_in) ** 0.5 * 1.4, seed)

    def bn(name, c, wscale=1.0):
        sd[name + ".weight"] = uniform(name + ".weight", (c,), 0.7 * wscale, 1.3 * wscale, seed)
        sd[name + ".bias"] = _sym(name + ".bias", (c,), 0.1, seed)
        sd[name + ".running_mean"] = _sym(name + ".running_mean", (c,), 0.1, seed)
        sd[name + ".running_var"] = uniform(name + ".running_var", (c,), 0.6, 1.4, seed)

    def linear(name, out_f, in_f, wbound=None, bbound=None):
        wb = (1.0 / in_f) ** 0.5 * 1.7 if wbound is None else wbound
        sd[name + ".weight"] = _sym(name + ".weight", (out_f, in_f), wb, seed)
        sd[name + ".bias"] = _sym(name + ".bias", (out_f,), 0.05 if bbound is None else bbound, seed)

    def norm(name, c):
        sd[name + ".weight"] = uniform(name + ".weight", (c,), 0.8, 1.2, seed)
        sd[name + ".bias"] = _sym(name + ".bias", (c,), 0.1, seed)

    p = "backbone.body"
    conv(p + ".conv1.weight", 64, 3, 7)
    bn(p + ".bn1", 64)
    inpl = 64
    for li, (nb, planes) in enumerate(zip(RESNET50_BLOCKS, RESNET50_PLANES)):
        for bi in range(nb):
            q = f"{p}.layer{li + 1}.{bi}"
            conv(q + ".conv1.weight", planes, inpl, 1)
            bn(q + ".bn1", planes)
            conv(q + ".conv2.weight", planes, planes, 3)
            bn(q + ".bn2", planes)
            conv(q + ".conv3.weight", planes * 4, planes, 1)
            bn(q + ".bn3", planes * 4, wscale=0.5)
            if bi == 0:
                conv(q + ".downsample.0.weight", planes * 4, inpl, 1)
                bn(q + ".downsample.1", planes * 4, wscale=0.8)
            inpl = planes * 4
    # 1x1 projections + GroupNorm (stage 2 registers both, A2/models/anchor_detr.py:68-84)
    projs = ["input_proj.0"] + (["aggr_input_proj.0"] if cfg.stage == 2 else [])
    for name in projs:
        cin = 4096 if name.startswith("aggr") else 2048
        sd[name + ".0.weight"] = _sym(name + ".0.weight", (E, cin, 1, 1), (3.0 / cin) ** 0.5, seed)
        sd[name + ".0.bias"] = _sym(name + ".0.bias", (E,), 0.05, seed)
        norm(name + ".1", E)
    t = "transformer"

    def attn_rcda(name):
        w = _sym(name + ".in_proj_weight", (5 * E, E), (1.0 / E) ** 0.5 * 3.0, seed)
        w[4 * E:] *= 0.25                       # value rows: keep the residual branch below the trunk
        sd[name + ".in_proj_weight"] = w
        sd[name + ".in_proj_bias"] = _sym(name + ".in_proj_bias", (5 * E,), 0.05, seed)
        linear(name + ".out_proj", E, E, wbound=(1.0 / E) ** 0.5 * 1.2)

    def ffn(name):
        linear(name + ".linear1", F_, E)
        linear(name + ".linear2", E, F_, wbound=(1.0 / F_) ** 0.5 * 0.8)
        norm(name + ".norm2", E)

    for i in range(cfg.enc_layers):
        q = f"{t}.encoder_layers.{i}"
        attn_rcda(q + ".self_attn")
        norm(q + ".norm1", E)
        ffn(q + ".ffn")
    for i in range(cfg.dec_layers):
        q = f"{t}.decoder_layers.{i}"
        attn_rcda(q + ".cross_attn")
        norm(q + ".norm1", E)
        w = _sym(q + ".self_attn.in_proj_weight", (3 * E, E), (1.0 / E) ** 0.5 * 3.0, seed)
        w[2 * E:] *= 0.25
        sd[q + ".self_attn.in_proj_weight"] = w
        sd[q + ".self_attn.in_proj_bias"] = _sym(q + ".self_attn.in_proj_bias", (3 * E,), 0.05, seed)
        linear(q + ".self_attn.out_proj", E, E, wbound=(1.0 / E) ** 0.5 * 1.2)
        norm(q + ".norm2", E)
        ffn(q + ".ffn")
    sd[cfg.pattern_key] = _sym(cfg.pattern_key, (cfg.num_query_pattern, E), 1.0, seed)
    if cfg.spatial_prior == "learned":
        sd[t + ".position.weight"] = uniform(t + ".position.weight", (cfg.num_query_position, 2), 0.02, 0.98, seed)
    for name in ("adapt_pos2d", "adapt_pos1d"):
        linear(f"{t}.{name}.0", E, E, wbound=(1.0 / E) ** 0.5 * 2.5)
        linear(f"{t}.{name}.2", E, E, wbound=(1.0 / E) ** 0.5 * 4.0)
    # heads: ONE module shared by all decoder layers but emitted dec_layers times in the state_dict
    # (transformer.py:104-107).  Values identical across k.
    heads = {}
    heads["cls_embed.weight"] = _sym("cls_embed.weight", (2, E), 0.2, seed)
    heads["cls_embed.bias"] = (_sym("cls_embed.bias", (2 if cfg.stage == 2 else 1,), 0.3, seed) - 2.0)
    for hname, nout in (("bbox_embed", 4),) + ((("bbox_variance", 2),) if cfg.stage == 2 else ()):
        for j in range(3):
            o = E if j < 2 else nout
            heads[f"{hname}.layers.{j}.weight"] = _sym(f"{hname}.layers.{j}.weight", (o, E), (1.0 / E) ** 0.5 * (1.7 if j < 2 else 1.0), seed)
            heads[f"{hname}.layers.{j}.bias"] = _sym(f"{hname}.layers.{j}.bias", (o,), 0.05, seed)
    heads["bbox_embed.layers.2.bias"][2:] -= 2.0           # keeps boxes small, like the reference's -2.0 init
    if cfg.stage == 2:
        heads["bbox_variance.layers.2.bias"] += 0.6        # keeps sigma > 0 (log of a negative -> NaN, §8a-8)
    for k in range(cfg.dec_layers):
        for name, v in heads.items():
            mod, rest = name.split(".", 1)
            sd[f"{t}.{mod}.{k}.{rest}"] = v
    return sd


# ------------------------------------------------------------------ synthetic inputs (SURVEY.md §8d laws)
def make_inputs(B, S, T=50, seed=0, stage=2, Q=None):
    img = (uniform("image", (B, 3, S, S), -1.0, 1.0, seed) * 1.7)
    x1y1 = uniform("rect_xy", (B, 3, 2), 0.05, 0.6, seed)
    wh = uniform("rect_wh", (B, 3, 2), 0.05, 0.3, seed)
    rects = torch.cat([x1y1, (x1y1 + wh).clamp(max=0.999)], -1)
    out = {"image": img, "rects": rects}
    if stage == 2:
        tg = []
        for b in range(B):
            c = uniform(f"tgt_c{b}", (T, 2), 0.1, 0.9, seed)
            s = uniform(f"tgt_s{b}", (T, 2), 0.02, 0.12, seed)
            tg.append({"boxes": torch.cat([c, s], -1), "labels": torch.zeros(T, dtype=torch.int64)})
        out["targets"] = tg
    else:
        n = Q if Q is not None else T
        out["points"] = uniform("pts", (B, n, 2), 0.1, 0.9, seed)
        out["whs"] = uniform("whs", (B, n, 2), 0.02, 0.12, seed)
    return out
