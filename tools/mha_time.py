"""Decoder self-attention core timings at the C3 shape (B=16, L=300): tensor-core path vs CDETR_MHA_LEGACY=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"; E = 256; nh = 8
for (Bz, Lq) in [(16, 300), (8, 300)]:
    qkv = torch.randn(Bz * Lq, 3 * E, device=dev)
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
    zs = lambda: torch.zeros(2, Bz * Lq, E, device=dev, dtype=torch.bfloat16)
    o = zs(); lse = torch.empty(Bz, nh, Lq, device=dev); dO = torch.randn(Bz * Lq, E, device=dev)
    dsum = torch.empty(Bz, nh, Lq, device=dev); dq, dk, dv = zs(), zs(), zs()
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tf = timeit(lambda: L.call("cdetr_mha_fwd", Bz, Lq, E, nh, q, k, v, 3 * E, o, lse))
    tb = timeit(lambda: L.call("cdetr_mha_bwd", Bz, Lq, E, nh, q, k, v, 3 * E, o, lse, dO, dsum, dq, dk, dv))
    print(f"[legacy={os.environ.get('CDETR_MHA_LEGACY', '0')}] B={Bz} L={Lq}: fwd {tf:.1f} us, bwd {tb:.1f} us", flush=True)
