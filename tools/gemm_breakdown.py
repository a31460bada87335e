"""Per-shape time of every GEMM launch in one train step.

Pass 1 records every cdetr_gemm call of one eager step (shape + epilogue flavour, CUDA events around each launch:
includes launch gaps, so tiny GEMMs look slower than they are).  Pass 2 replays each distinct call REPS times
back-to-back between two events (no launch gaps; operands of small GEMMs become L2-resident, as they mostly are
in the real step where the producer kernel has just written them)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from counting_detr_b200 import _lib as L, synthetic as SY
from counting_detr_b200.models import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
REPS = 10
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
model, crit, _ = build_model(SY.default_args(st, num_query_position=Q, device="cuda"))
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
model.to(dev).train()
model._auto_graph = False
inp = SY.make_inputs(B, S, T=T, stage=st, Q=Q)
img = inp["image"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]] if st == 2 else {"points": inp["points"].to(dev), "whs": inp["whs"].to(dev)}

def step():
    model.zero_grad(set_to_none=True)
    if st == 2:
        out, _ = model(img, None, inp["rects"])
    else:
        out = model(img, targets["points"])
    ld = crit(out, targets)
    sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()

eng = model.engine()
eng.side_stream, eng.aux_streams = None, []
for _ in range(2): step()
torch.cuda.synchronize()
orig = L.gemm
calls = []
def traced(a, b, M, N, K, mode=0, **kw):
    calls.append((a, b, M, N, K, mode, dict(kw)))
    return orig(a, b, M, N, K, mode=mode, **kw)
L.gemm = traced
import counting_detr_b200.engine as EN
EN.L.gemm = traced
L.GEMM_TRACE = []
step(); torch.cuda.synchronize()
trace, L.GEMM_TRACE = L.GEMM_TRACE, None
L.gemm = orig; EN.L.gemm = orig

def flavour(mode, kw):
    return (mode, kw.get("split_k", 1), kw.get("block_n", 0),
            ("split" if kw.get("out_split") is not None else "") + ("f32" if kw.get("out_f32") is not None else ""),
            "+add" if kw.get("add_split") is not None or kw.get("add_f32") is not None else "",
            "+mask" if kw.get("mask") is not None else "", "+acc" if kw.get("accumulate") else "")

agg = collections.OrderedDict()
for (a, b, M, N, K, mode, kw), (_, _, _, e0, e1) in zip(calls, trace):
    key = (M, N, K) + flavour(mode, kw)
    ent = agg.setdefault(key, dict(n=0, us_eager=0.0, call=(a, b, M, N, K, mode, kw)))
    ent["n"] += 1; ent["us_eager"] += e0.elapsed_time(e1) * 1e3
for key, ent in agg.items():
    a, b, M, N, K, mode, kw = ent["call"]
    for _ in range(2): orig(a, b, M, N, K, mode=mode, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS): orig(a, b, M, N, K, mode=mode, **kw)
    e1.record(); torch.cuda.synchronize()
    ent["us"] = e0.elapsed_time(e1) * 1e3 / REPS
tot = sum(e["us"] * e["n"] for e in agg.values()); tot_e = sum(e["us_eager"] for e in agg.values())
fl_tot = sum(2.0 * k[0] * k[1] * k[2] * e["n"] for k, e in agg.items())
print(f"{name}: {sum(e['n'] for e in agg.values())} gemm launches, {len(agg)} distinct; back-to-back total {tot/1e3:.2f} ms "
      f"(eager per-launch events {tot_e/1e3:.2f} ms), {fl_tot/1e9:.0f} GF algorithmic -> {fl_tot/tot/1e6:.1f} TF/s alg")
for key, e in sorted(agg.items(), key=lambda kv: -kv[1]["us"] * kv[1]["n"])[:70]:
    M, N, K = key[:3]; fl = 2.0 * M * N * K
    byt = 4.0 * (M * K + N * K) + (4.0 * M * N * (("split" in key[6]) + ("f32" in key[6]) + (key[7] != "") + 0.5 * (key[8] != "")))
    print(f"{e['us']*e['n']:8.0f} us {100*e['us']*e['n']/tot:5.1f}% n={e['n']:3d} avg={e['us']:7.1f} us (eager {e['us_eager']/e['n']:6.1f}) "
          f"{fl/e['us']/1e6:7.1f} TF/s alg {byt/e['us']/1e3:6.0f} GB/s  mode={key[3]} M={M} N={N} K={K} splitk={key[4]} bn={key[5]} out={key[6]}{key[7]}{key[8]}{key[9]}")
