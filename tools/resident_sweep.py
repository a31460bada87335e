"""Resident-B schedule (weight slab of an n-tile kept in shared memory, only A streams) vs the streaming schedules on
the L2-bound short-K shapes (CDETR_GEMM_RESIDENT read per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
from epi_debug import timed
dev = "cuda"
for (M, N, K, flav) in [(16384, 1024, 256, "split"), (16384, 1024, 256, "split+add+mask"), (16384, 256, 256, "f32"), (16384, 2304, 256, "split"),
                        (16384, 4096, 256, "split"), (65536, 512, 128, "split+add"), (65536, 512, 256, "split+add+mask"), (262144, 256, 64, "split+add"),
                        (4800, 1024, 256, "split"), (16384, 2048, 512, "split+add")]:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    kw = dict(bias=torch.randn(N, device=dev))
    if flav.startswith("split"):
        kw["out_split"] = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
    else:
        kw["out_f32"] = torch.empty(M, N, device=dev)
    if "add" in flav:
        kw["add_split"] = L.to_split(torch.randn(M, N, device=dev))
    if "mask" in flav:
        kw["mask"] = L.to_split(torch.randn(M, N, device=dev))
    else:
        kw["relu"] = True
    res = {}
    for tag, env, bn in [("auto", {}, 0), ("resident bn128", {"CDETR_GEMM_RESIDENT": "1", "CDETR_GEMM_PAIR": "0"}, 128),
                         ("resident bn64", {"CDETR_GEMM_RESIDENT": "1", "CDETR_GEMM_PAIR": "0"}, 64)]:
        for k, v in env.items():
            os.environ[k] = v
        try:
            res[tag] = timed(lambda: L.gemm(A, B, M, N, K, block_n=bn, **kw))
        except Exception as e:
            res[tag] = float("nan")
        for k in env:
            os.environ.pop(k)
    print(f"M={M} N={N} K={K} {flav}: " + " | ".join(f"{t} {v:.1f} us" for t, v in res.items()), flush=True)
