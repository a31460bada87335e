"""Per-shape time of every GEMM launch in one C3 train step (CUDA events around each launch)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from counting_detr_b200 import _lib as L, synthetic as SY
from counting_detr_b200.models import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
model, crit, _ = build_model(SY.default_args(st, num_query_position=Q, device="cuda"))
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
model.to(dev).train()
inp = SY.make_inputs(B, S, T=T, stage=st, Q=Q)
img = inp["image"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]]

def step():
    model.zero_grad(set_to_none=True)
    out, _ = model(img, None, inp["rects"])
    ld = crit(out, targets)
    sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()

for _ in range(2): step()
torch.cuda.synchronize()
# record mode per call too
orig = L.gemm
modes = []
def traced(a, b, M, N, K, mode=0, **kw):
    modes.append((mode, kw.get("split_k", 1), kw.get("block_n", 0), "split" if kw.get("out_split") is not None else "f32",
                  "+add" if kw.get("add_split") is not None or kw.get("add_f32") is not None else "", "+mask" if kw.get("mask") is not None else ""))
    return orig(a, b, M, N, K, mode=mode, **kw)
L.gemm = traced
import counting_detr_b200.engine as EN
EN.L.gemm = traced
L.GEMM_TRACE = []
step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (M, N, K, e0, e1), md in zip(L.GEMM_TRACE, modes):
    key = (md[0], M, N, K) + md[1:]
    agg[key][0] += 1; agg[key][1] += e0.elapsed_time(e1) * 1e3; agg[key][2] += 2.0 * M * N * K
tot = sum(v[1] for v in agg.values())
print(f"total gemm {tot/1e3:.2f} ms, {sum(v[2] for v in agg.values())/1e9:.0f} GF")
for key, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{us:8.0f} us {100*us/tot:5.1f}% n={n:3d} avg={us/n:7.1f} us  {fl/us/1e6:7.1f} TF/s alg  mode={key[0]} M={key[1]} N={key[2]} K={key[3]} splitk={key[4]} bn={key[5]} out={key[6]}{key[7]}{key[8]}")
