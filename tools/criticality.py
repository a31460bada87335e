"""Which kernels sit on the critical path of the captured step?  For each family of entry points the step is
re-captured with those launches SKIPPED (results are wrong, the schedule is otherwise identical) and re-timed: the
drop of the step time is what the family costs on the critical path (its serial kernel time may be much larger when
it runs on a parallel branch).  python tools/criticality.py [c3|c4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from counting_detr_b200 import _lib as L, synthetic as SY
from counting_detr_b200.models import build_model
from counting_detr_b200.step import CapturedStep

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
model, crit, _ = build_model(SY.default_args(st, num_query_position=Q, device="cuda"))
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
model.to(dev).train(); crit.train()
inp = SY.make_inputs(B, S, T=T, stage=st, Q=Q)
img = inp["image"].to(dev); rects = inp["rects"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]]

FAMILIES = [
    ("baseline", []),
    ("rcda_bwd_v", ["cdetr_rcda_bwd_v_tc"]),
    ("rcda_bwd_q", ["cdetr_rcda_bwd_q_tc"]),
    ("rcda_bwd_k", ["cdetr_rcda_bwd_k"]),
    ("rcda_bwd all", ["cdetr_rcda_bwd_v_tc", "cdetr_rcda_bwd_q_tc", "cdetr_rcda_bwd_k"]),
    ("rcda_fwd", ["cdetr_rcda_fwd_tc"]),
    ("mha_fwd", ["cdetr_mha_fwd"]),
    ("mha_bwd", ["cdetr_mha_bwd"]),
    ("all attention", ["cdetr_rcda_bwd_v_tc", "cdetr_rcda_bwd_q_tc", "cdetr_rcda_bwd_k", "cdetr_rcda_fwd_tc", "cdetr_mha_fwd",
                       "cdetr_mha_bwd"]),
    ("layernorm_fwd", ["cdetr_layernorm_fwd"]),
    ("layernorm_bwd", ["cdetr_layernorm_bwd"]),
    ("colsum", ["cdetr_colsum"]),
    ("add_bcast", ["cdetr_add_bcast"]),
    ("combine_bcast", ["cdetr_combine_bcast"]),
    ("reduce_axis", ["cdetr_reduce_axis"]),
    ("matcher+loss", ["cdetr_match_cost", "cdetr_lsap", "cdetr_set_loss_fwd", "cdetr_set_loss_bwd"]),
    ("groupnorm", ["cdetr_groupnorm_fwd", "cdetr_groupnorm_bwd"]),
    ("box_head", ["cdetr_box_head_fwd", "cdetr_box_head_bwd"]),
    ("sine_embed", ["cdetr_sine_embed", "cdetr_sine_embed_bwd"]),
    ("unpack_conv_grad", ["cdetr_unpack_conv_grad"]),
    ("exemplar", ["cdetr_exemplar_concat", "cdetr_exemplar_concat_bwd"]),
    ("im2col/col2im/pool", ["cdetr_stem_im2col", "cdetr_im2col3x3", "cdetr_col2im3x3", "cdetr_maxpool3x3s2", "cdetr_subsample2",
                            "cdetr_upsample2_zero"]),
]
known = set(L._SIGS)
base = None
for tag, names in FAMILIES:
    names = [n for n in names if n in known]
    L.SKIP = set(names) if names else None
    step = CapturedStep(model, crit)
    for _ in range(3):
        step(img, targets, rects=rects)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step(img, targets, rects=rects)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    base = ms if base is None else base
    print(f"{tag:22s} {ms:8.3f} ms/step   on the critical path: {base - ms:6.3f} ms   ({step.launches_per_step} launches)", flush=True)
    del step
L.SKIP = None
