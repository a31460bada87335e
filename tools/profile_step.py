"""One eager train step of a bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from counting_detr_b200.models import build_model
from counting_detr_b200 import synthetic as SY

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
st, B, S, Q, T = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
nl = int(os.environ.get("PROFILE_LAYERS", "6"))   # 1 = one encoder + one decoder layer (same shapes, every kernel type once)
args = SY.default_args(st, num_query_position=Q, device="cuda", enc_layers=nl, dec_layers=nl)
model, crit, _ = build_model(args)
model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q, enc_layers=nl, dec_layers=nl), 0), strict=True)
model.to(dev).train()
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
    loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
