"""GPU-side cost of small GEMM launches: REPS identical launches captured in one CUDA graph (no host launch cost),
replayed and timed with events.  Variants through the library's tuning environment variables."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
REPS = 50
shapes = [(512, 256, 256, 0), (4800, 256, 256, 0), (16384, 256, 256, 0), (16384, 256, 256, 1), (256, 256, 4800, 2), (256, 256, 16384, 2),
          (256, 256, 512, 2)]


def bench(M, N, K, kind):
    if kind == 2:
        A = L.to_split(torch.randn(K, M, device=dev)); B = L.to_split(torch.randn(K, N, device=dev))
        out = torch.zeros(M, N, device=dev)
        tiles = -(-M // 128) * -(-N // 128)
        sk = max(1, min(-(-K // 64) // 4, (2 * 148) // tiles))
        fn = lambda: L.gemm(A, B, M, N, K, mode=1, out_f32=out, accumulate=True, split_k=sk, block_n=128)
    else:
        A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
        out = torch.empty(M, N, device=dev); outs = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev)
        fn = (lambda: L.gemm(A, B, M, N, K, out_f32=out, bias=bias)) if kind == 0 else (lambda: L.gemm(A, B, M, N, K, out_split=outs, bias=bias, relu=True))
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3


# reference: an empty-ish kernel in a graph
x = torch.zeros(1024, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    L.call("cdetr_scale", x, 1024, 1.0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(REPS):
            L.call("cdetr_scale", x, 1024, 1.0)
g.replay(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"tiny elementwise kernel in graph: {e0.elapsed_time(e1) / REPS * 1e3:.2f} us/launch")
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CDETR_"))
for M, N, K, kind in shapes:
    print(f"[{tag}] M={M} N={N} K={K} kind={kind}: {bench(M, N, K, kind):.2f} us/launch (graph replay)", flush=True)

# ---- in-kernel timeline of CTA 0 (cdetr_gemm_debug_timeline): where do the ~10 us of a tiny GEMM go?
import ctypes
NL = 6
buf = torch.zeros(8 * NL * 8, dtype=torch.int64, device=dev)
names = ["entry", "setup", "ops_landed", "mma_issued", "acc_seen", "epi_done", "all_done", "tmem_freed"]
for M, N, K, kind in [(512, 256, 256, 0), (16384, 256, 256, 0), (16384, 1024, 256, 1)]:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    out = torch.empty(M, N, device=dev); outs = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
    fn = (lambda: L.gemm(A, B, M, N, K, out_f32=out, bias=bias)) if kind == 0 else (lambda: L.gemm(A, B, M, N, K, out_split=outs, bias=bias, relu=True))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    L.lib().cdetr_gemm_debug_timeline(ctypes.c_void_p(buf.data_ptr()), NL)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(NL):
                fn()
    L.lib().cdetr_gemm_debug_timeline(None, 0)
    g.replay(); torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    t = buf.view(-1, 8)[:NL].cpu()
    print(f"timeline M={M} N={N} K={K} kind={kind} (ns relative to entry of launch; gap = entry - previous launch's tmem_freed)")
    for i in range(1, NL):
        rel = [int(t[i, k] - t[i, 0]) for k in range(8)]
        gap = int(t[i, 0] - t[i - 1, 7])
        print(f"  launch {i}: gap={gap:6d} " + " ".join(f"{n}={v}" for n, v in zip(names[1:], rel[1:])))
