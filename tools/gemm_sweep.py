"""Tile-shape / pipeline-depth / CTAs-per-SM sweep of the TN GEMM on the shapes that dominate the C3 step.

Each configuration is timed as a CUDA graph of REPS launches (device-bound: no host launch gaps; operands of the
small shapes are L2-resident, as they are in the real step right after their producer kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
REPS = 8
shapes = [  # (M, N, K, epilogue)
    (512, 256, 256, "f32"), (4800, 256, 256, "f32"), (4800, 1024, 256, "split"), (4800, 256, 1024, "f32"),
    (16384, 256, 256, "f32"), (16384, 1024, 256, "split"), (16384, 1024, 256, "split+add+mask"), (16384, 256, 1024, "f32"),
    (16384, 2304, 256, "split"), (65536, 512, 128, "split+add"), (65536, 128, 512, "split"), (16384, 512, 2048, "split"),
    (16384, 2048, 512, "split+add"), (16384, 512, 4608, "split"), (262144, 256, 64, "split+add"), (262144, 64, 256, "split"),
    (65536, 128, 1152, "split"), (16384, 256, 2304, "split"), (16384, 4096, 256, "split")]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for _ in range(REPS):
            fn()
    torch.cuda.synchronize()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / REPS)
    return sorted(ts)[1]


for (M, N, K, ep) in shapes:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    kw = dict(bias=torch.randn(N, device=dev))
    if "split" in ep:
        kw["out_split"] = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16); kw["relu"] = True
    else:
        kw["out_f32"] = torch.empty(M, N, device=dev)
    if "add" in ep:
        kw["add_split"] = L.to_split(torch.randn(M, N, device=dev))
    if "mask" in ep:
        kw["mask"] = L.to_split(torch.randn(M, N, device=dev))
    res = []
    for bn in (64, 128, 256):
        if bn > max(N, 64):
            continue
        for st in (0, 2, 3, 4):           # 0 = the library's own choice for this tile width
            if st:
                os.environ["CDETR_GEMM_STAGES"] = str(st)
            else:
                os.environ.pop("CDETR_GEMM_STAGES", None)
            try:
                t = timed(lambda: L.gemm(A, B, M, N, K, block_n=bn, **kw))
                res.append((t, bn, 1, st))
            except Exception:
                pass
    os.environ.pop("CDETR_GEMM_STAGES", None)
    auto = timed(lambda: L.gemm(A, B, M, N, K, **kw))
    res.sort()
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} {ep}: auto={auto:.1f}us ({fl/auto/1e6:.0f} TF/s alg) | best "
          + "  ".join(f"bn{bn}/c{c}/s{st}:{t:.1f}" for t, bn, c, st in res[:6]) + f" | worst {res[-1][0]:.1f}", flush=True)
    if os.environ.get("SWEEP_FULL"):
        print("      " + " ".join(f"bn{bn}/c{c}/s{st}:{t:.1f}" for t, bn, c, st in sorted(res, key=lambda r: (r[1], str(r[2]), r[3]))), flush=True)

# ---- weight-gradient shapes (mode 1: D[M,N] = A[K,M]^T B[K,N], split-K accumulation through TMA reduce-add)
nt_shapes = [(512, 4608, 16384, 2), (256, 2304, 16384, 8), (1024, 256, 16384, 18), (256, 1024, 16384, 18), (256, 256, 16384, 64),
             (2048, 512, 16384, 4), (512, 2048, 16384, 4), (128, 1152, 65536, 32), (512, 128, 65536, 74), (256, 256, 4800, 18)]
for (M, N, K, sk0) in nt_shapes:
    A = L.to_split(torch.randn(K, M, device=dev)); B = L.to_split(torch.randn(K, N, device=dev))
    out = torch.zeros(M, N, device=dev)
    res = []
    for bn in (64, 128, 256):
        if bn > N:
            continue
        for sk in sorted({max(1, sk0 // 2), sk0, sk0 * 2}):
            try:
                t = timed(lambda: L.gemm(A, B, M, N, K, mode=1, out_f32=out, accumulate=True, split_k=sk, block_n=bn))
                res.append((t, bn, sk))
            except Exception:
                pass
    res.sort()
    fl = 2.0 * M * N * K
    print(f"NT M={M} N={N} K={K}: " + "  ".join(f"bn{bn}/sk{sk}:{t:.1f}" for t, bn, sk in sorted(res, key=lambda r: (r[1], r[2])))
          + f" | best {fl/res[0][0]/1e6:.0f} TF/s alg", flush=True)
