"""CTA-pair (cta_group::2, 256x256 tiles) vs single-CTA tiles on the forward / dgrad shapes of the C3 step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
REPS = 8
shapes = [(16384, 512, 4608, "split", None), (16384, 512, 4608, "split+mask", None), (16384, 512, 2048, "split+mask", None),
          (16384, 2048, 1024, "split", None), (16384, 1024, 2048, "split", None), (16384, 256, 2304, "split", None),
          (16384, 256, 1024, "f32+add", None), (16384, 1024, 512, "split+add+mask", None), (16384, 2048, 512, "split+add", None),
          (16384, 4096, 256, "split", None), (16384, 2304, 256, "split", None), (16384, 1024, 256, "split", None),
          (16384, 256, 256, "f32", None), (65536, 512, 256, "split+add+mask", None), (65536, 256, 512, "split", None),
          (16384, 512, 4608, "split", (32, 32, 512, 2, 1)), (16384, 256, 2304, "split", (32, 32, 256, 1, 1)),
          (16384, 512, 4608, "split+mask", (32, 32, 512, 2, -1))]


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


for (M, N, K, ep, conv) in shapes:
    if conv is None:
        A = L.to_split(torch.randn(M, K, device=dev))
    else:
        A = L.to_split(torch.randn(M, conv[2], device=dev))
    B = L.to_split(torch.randn(N, K, device=dev))
    kw = dict(bias=torch.randn(N, device=dev))
    if ep.startswith("split"):
        kw["out_split"] = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
    else:
        kw["out_f32"] = torch.empty(M, N, device=dev)
    if "add" in ep:
        if ep.startswith("split"):
            kw["add_split"] = L.to_split(torch.randn(M, N, device=dev))
        else:
            kw["add_f32"] = torch.randn(M, N, device=dev)
    if "mask" in ep:
        kw["mask"] = L.to_split(torch.randn(M, N, device=dev))
    else:
        kw["relu"] = True
    if conv is not None:
        kw["conv"] = conv
    res = {}
    for mode in ("0", "1"):
        os.environ["CDETR_GEMM_PAIR"] = mode
        res[mode] = timed(lambda: L.gemm(A, B, M, N, K, **kw))
    os.environ.pop("CDETR_GEMM_PAIR")
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} {ep}{' conv' if conv else ''}: single {res['0']:.1f} us ({3*fl/res['0']/1e6:.0f} TF/s issued) | pair {res['1']:.1f} us "
          f"({3*fl/res['1']/1e6:.0f} TF/s issued)  x{res['0']/res['1']:.2f}", flush=True)
