"""Tile-shape / pipeline-depth sweep of the TN GEMM on the shapes that dominate the C3 step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
shapes = [(16384, 256, 256), (16384, 1024, 256), (16384, 256, 1024), (4800, 256, 256), (16384, 2304, 256), (65536, 512, 128),
          (65536, 128, 512), (16384, 512, 2048), (16384, 2048, 512), (16384, 512, 4608), (262144, 256, 64), (262144, 64, 256),
          (262144, 64, 576), (65536, 128, 1152), (16384, 256, 2304), (16384, 4608, 512)]
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (M, N, K) in shapes:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    out = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
    res = []
    for bn in (64, 128, 256):
        if bn > max(N, 64) and bn != 64: continue
        for st in (1, 2, 3, 4):
            os.environ["CDETR_GEMM_STAGES"] = str(st)
            try:
                for _ in range(2): L.gemm(A, B, M, N, K, out_split=out, bias=bias, relu=True, block_n=bn)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ts = []
                for _ in range(5):
                    flush.zero_()
                    e0.record(); L.gemm(A, B, M, N, K, out_split=out, bias=bias, relu=True, block_n=bn); e1.record()
                    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
                res.append((sorted(ts)[2], bn, st))
            except Exception as ex:
                pass
    os.environ.pop("CDETR_GEMM_STAGES", None)
    L.gemm(A, B, M, N, K, out_split=out, bias=bias, relu=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        flush.zero_(); e0.record(); L.gemm(A, B, M, N, K, out_split=out, bias=bias, relu=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    res.sort()
    print(f"M={M} N={N} K={K}: auto={sorted(ts)[2]:.1f}us | best " + "  ".join(f"bn{bn}/s{st}:{t:.1f}" for t, bn, st in res[:5]) + f" | worst {res[-1][0]:.1f}", flush=True)
