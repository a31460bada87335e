"""Where the end-to-end step spends its host time (bench.py's e2e loop, pinned host batches -> DevicePrefetcher ->
CapturedStep -> loss.item()): perf_counter around every piece of the loop and around the pieces of CapturedStep.__call__.
Run on the GPU box: python tools/e2e_breakdown.py [c3|c4] [steps]."""
import os
import sys
import time
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from counting_detr_b200 import synthetic as SY
from counting_detr_b200.data import DevicePrefetcher
from counting_detr_b200.models import build_model
from counting_detr_b200.step import CapturedStep

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
margs = SY.default_args(st, num_query_position=Q, device=str(dev))
model, crit, _ = build_model(margs)
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
model.to(dev).train(); crit.train()
host = []
for i in range(2):
    inp = SY.make_inputs(B, S, T=T, seed=i, stage=st, Q=Q)
    host.append({"image": inp["image"].pin_memory(), "rects": inp["rects"].pin_memory(),
                 "targets": [{"boxes": t["boxes"].pin_memory(), "labels": t["labels"]} for t in inp["targets"]]})

acc = defaultdict(float)


def timed_method(obj, attr, tag):
    fn = getattr(obj, attr)

    def wrap(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        acc[tag] += time.perf_counter() - t0
        return r
    setattr(obj, attr, wrap)


step = CapturedStep(model, crit)
for i in range(4):
    b = {k: (v.to(dev) if isinstance(v, torch.Tensor) else [{kk: vv.to(dev) for kk, vv in t.items()} for t in v])
         for k, v in host[i & 1].items()}
    step(b["image"], b["targets"], rects=b["rects"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    step._graph.replay()
e1.record(); torch.cuda.synchronize()
print(f"{name}: bare graph replay {e0.elapsed_time(e1) / steps:.3f} ms/step")
t0 = time.perf_counter()
for i in range(steps):
    step._graph.replay()
t_launch = (time.perf_counter() - t0) / steps
torch.cuda.synchronize()
print(f"host time of graph.replay() while the queue is busy: {t_launch * 1e6:.0f} us")
# latency of one replay on an idle device: host call -> results visible
lat = []
for i in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step._graph.replay()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    lat.append((t1 - t0, time.perf_counter() - t0))
print(f"idle device: replay() returns after {min(l[0] for l in lat) * 1e6:.0f} us, step visible after "
      f"{min(l[1] for l in lat) * 1e3:.3f} ms")

timed_method(step, "_sig", "call._sig")
timed_method(step, "_fill", "call._fill")
timed_method(step, "_prep", "call._prep")
timed_method(model, "_current_version", "call._current_version")
timed_method(step._graph, "replay", "call.replay")


def loop(n, sync):
    pf = DevicePrefetcher((host[i & 1] for i in range(n)), dev, defer=sync)
    it = iter(pf)
    k = 0
    pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    while True:
        t0 = time.perf_counter()
        try:
            b = next(it)
        except StopIteration:
            break
        t1 = time.perf_counter()
        _, total = step(b["image"], b["targets"], rects=b["rects"])
        t2 = time.perf_counter()
        if sync:
            pf.kick()
            t3 = time.perf_counter()
            total.item()
            t4 = time.perf_counter()
        else:
            t3 = time.perf_counter()
            pin[k & 1].copy_(total, non_blocking=True); evs[k & 1].record()
            if k >= 1:
                evs[(k & 1) ^ 1].synchronize()
            t4 = time.perf_counter()
        acc["next(prefetcher)"] += t1 - t0; acc["step()"] += t2 - t1; acc["kick()"] += t3 - t2; acc["loss read"] += t4 - t3
        k += 1


for sync in (True, False):
    loop(3, sync)
    torch.cuda.synchronize()
    acc.clear()
    e0.record()
    t0 = time.perf_counter()
    loop(steps, sync)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    print(f"--- {'sync loss.item() every step' if sync else 'lagged loss read'}: {e0.elapsed_time(e1) / steps:.3f} ms/step "
          f"(wall {wall * 1e3:.3f})")
    for k_, v in sorted(acc.items()):
        print(f"    {k_:28s} {v / steps * 1e6:9.1f} us/step")
