"""Poor man's timeline of one CUDA-graph replay of the train step: every GEMM launch records CTA-0 entry / exit
device timestamps (cdetr_gemm_debug_timeline), so the gaps between GEMMs show where the non-GEMM kernels sit on the
critical path and how much the side streams overlap."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from counting_detr_b200 import _lib as L, synthetic as SY
from counting_detr_b200.models import build_model
import counting_detr_b200.engine as EN

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
model, crit, _ = build_model(SY.default_args(st, num_query_position=Q, device="cuda"))
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
model.to(dev).train()
model._auto_graph = False                      # this tool captures the step itself (per-GEMM device timestamps)
inp = SY.make_inputs(B, S, T=T, stage=st, Q=Q)
img = inp["image"].to(dev)
rects = inp["rects"].to(dev) if "rects" in inp else None
targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]] if st == 2 else {"points": inp["points"].to(dev), "whs": inp["whs"].to(dev)}


def step():
    model.zero_grad(set_to_none=True)
    if st == 2:
        out, _ = model(img, None, rects)
    else:
        out = model(img, targets["points"])
    ld = crit(out, targets)
    loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
    loss.backward()
    return loss.detach()


for _ in range(3):
    step()
torch.cuda.synchronize()
CAP = 1024
buf = torch.zeros(CAP * 8, dtype=torch.int64, device=dev)
calls = []
orig = L.gemm


def traced(a, b, M, N, K, mode=0, **kw):
    calls.append((M, N, K, mode, torch.cuda.current_stream().cuda_stream,
                  ("s" if kw.get("out_split") is not None else "") + ("f" if kw.get("out_f32") is not None else "")
                  + ("+add" if kw.get("add_split") is not None or kw.get("add_f32") is not None else "")
                  + ("+mask" if kw.get("mask") is not None else "") + ("+conv" if kw.get("conv") is not None else "")))
    return orig(a, b, M, N, K, mode=mode, **kw)


s = torch.cuda.Stream(priority=-1)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
    torch.cuda.synchronize()
    L.gemm = traced; EN.L.gemm = traced
    L.lib().cdetr_gemm_debug_timeline(ctypes.c_void_p(buf.data_ptr()), CAP)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    L.lib().cdetr_gemm_debug_timeline(None, 0)
    L.gemm = orig; EN.L.gemm = orig
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
t = buf.view(-1, 8)[:len(calls)].cpu()
t0 = int(t[:, 0].min())
rows = []
streams = {}
for i, c in enumerate(calls):
    sid = streams.setdefault(c[4], len(streams))
    rows.append((int(t[i, 0]) - t0, int(t[i, 7]) - t0, sid) + c[:4] + (c[5],))
rows.sort()
span = max(r[1] for r in rows)
print(f"replay {e0.elapsed_time(e1):.2f} ms; {len(rows)} GEMMs; first GEMM entry -> last GEMM exit {span/1e6:.2f} ms; streams {len(streams)}")
# union of busy intervals per stream and overall
def union(iv):
    iv = sorted(iv); tot = 0; cs, ce = iv[0]
    for a, b in iv[1:]:
        if a > ce:
            tot += ce - cs; cs, ce = a, b
        else:
            ce = max(ce, b)
    return tot + ce - cs
print(f"union of GEMM-busy time (any stream): {union([(r[0], r[1]) for r in rows])/1e6:.2f} ms")
for sid in range(len(streams)):
    iv = [(r[0], r[1]) for r in rows if r[2] == sid]
    print(f"  stream {sid}: {len(iv)} GEMMs, busy {union(iv)/1e6:.2f} ms, sum {sum(b - a for a, b in iv)/1e6:.2f} ms")
# main-stream chain: gaps between consecutive main-stream GEMMs
main = [r for r in rows if r[2] == 0]
gaps = []
for a, b in zip(main[:-1], main[1:]):
    gaps.append((b[0] - a[1], a, b))
print(f"main stream: sum of gaps between consecutive GEMMs {sum(max(g_[0], 0) for g_ in gaps)/1e6:.2f} ms")
print("largest gaps on the main stream (us): after -> before")
for gp, a, b in sorted(gaps, key=lambda x: -x[0])[:25]:
    print(f"  {gp/1e3:8.1f} us  at t={a[1]/1e6:6.2f} ms  after M={a[3]} N={a[4]} K={a[5]} m{a[6]} {a[7]}  -> M={b[3]} N={b[4]} K={b[5]} m{b[6]} {b[7]}")
# coarse phases: time histogram of main-stream occupancy per ms
print("per-ms: main-stream GEMM busy us | all-stream union us")
nms = int(span / 1e6) + 1
for k in range(nms):
    lo, hi = k * 1e6, (k + 1) * 1e6
    def clip(iv):
        return [(max(a, lo), min(b, hi)) for a, b in iv if b > lo and a < hi]
    m = clip([(r[0], r[1]) for r in main]); al = clip([(r[0], r[1]) for r in rows])
    print(f"  {k:3d} ms: {union(m)/1e3 if m else 0:7.1f} | {union(al)/1e3 if al else 0:7.1f}")
