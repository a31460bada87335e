"""Where does a short-K GEMM spend its time?  Same launches with parts of the epilogue disabled (results are wrong in
modes 1-3; timing only).  CDETR_GEMM_EPI_DEBUG: 0 full, 1 no TMA stores, 2 no staging/fence/stores, 3 TMEM read only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
REPS = 8


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for _ in range(REPS):
            fn()
    torch.cuda.synchronize(); g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / REPS)
    return sorted(ts)[1]


if __name__ == "__main__":
  for (M, N, K, bn) in [(16384, 1024, 256, 0), (16384, 1024, 256, 128), (16384, 256, 256, 0), (65536, 512, 128, 0), (16384, 2048, 512, 0)]:
      A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
      outs = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16); bias = torch.randn(N, device=dev)
      res = []
      for mode in ("0", "1", "2", "3"):
          os.environ["CDETR_GEMM_EPI_DEBUG"] = mode
          res.append(timed(lambda: L.gemm(A, B, M, N, K, out_split=outs, bias=bias, relu=True, block_n=bn)))
      os.environ.pop("CDETR_GEMM_EPI_DEBUG")
      print(f"M={M} N={N} K={K} bn={bn or 'auto'}: full {res[0]:.1f} us | no stores {res[1]:.1f} | no staging {res[2]:.1f} | TMEM read only {res[3]:.1f}", flush=True)
