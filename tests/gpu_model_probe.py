"""End-to-end GPU probe: build_model on cuda with the oracle's synthetic weights, forward + criterion +
backward, compared against the CPU oracle (outputs, losses, matching indices, parameter gradients)."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as OM, weights as OW, criterion as OC, ref_import as R
from counting_detr_b200.models import build_model

ap = argparse.ArgumentParser()
ap.add_argument("--S", type=int, default=128); ap.add_argument("--B", type=int, default=2)
ap.add_argument("--Q", type=int, default=50); ap.add_argument("--T", type=int, default=7)
ap.add_argument("--stage", type=int, default=2); ap.add_argument("--no-oracle", action="store_true")
a = ap.parse_args()
torch.set_num_threads(os.cpu_count())
cfg = OM.Config(stage=a.stage, num_query_position=a.Q)
sd = OW.make_state_dict(cfg, 0)
inp = OW.make_inputs(a.B, a.S, T=a.T, stage=a.stage, Q=a.Q)
args = R.default_args(a.stage, num_query_position=a.Q, device="cuda")
if a.stage == 1:
    for k in ("cost_class", "variance_loss_coef"):
        delattr(args, k)
model, crit, pp = build_model(args)
print("missing/unexpected:", model.load_state_dict(sd, strict=True))
model.cuda().train(); crit.train()
dev = "cuda"
t0 = time.time()
if a.stage == 2:
    out, ref = model(inp["image"].to(dev), None, inp["rects"].to(dev))
    targets = [{k: v.to(dev) for k, v in t.items()} for t in inp["targets"]]
    ld = crit(out, targets)
else:
    out = model(inp["image"].to(dev), inp["points"].to(dev))
    targets = {"points": inp["points"].to(dev), "whs": inp["whs"].to(dev)}
    ld = crit(out, targets)
loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
loss.backward()
torch.cuda.synchronize()
print(f"gpu fwd+loss+bwd (first call, incl. alloc) {time.time()-t0:.2f}s  loss={loss.item():.6f}")
print({k: round(v.item(), 6) for k, v in ld.items()})
if a.no_oracle:
    def step():
        model.zero_grad(set_to_none=True)
        if a.stage == 2:
            out, ref = model(img_d, None, rects_h)
        else:
            out = model(img_d, pts_d)
        ld = crit(out, targets)
        loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
        loss.backward()
        return loss
    img_d = inp["image"].to(dev); rects_h = inp.get("rects"); pts_d = inp["points"].to(dev) if a.stage == 1 else None
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(5): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"eager step: {ms:.2f} ms/step (wall {1e3*(time.time()-t0)/5:.2f} ms) -> {a.B/ms*1e3:.1f} img/s; loss {l.item():.5f}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    sys.exit(0)
# ---- oracle on CPU with autograd
sdg = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and ".bn" not in k and "downsample.1" not in k)
       for k, v in sd.items()}
# shared heads: alias the index-0 tensors so gradients accumulate like the reference's shared module
for k in list(sdg):
    for h in ("cls_embed", "bbox_embed", "bbox_variance"):
        if f"transformer.{h}." in k and f"transformer.{h}.0." not in k:
            idx = k.split(f"transformer.{h}.")[1].split(".")[0]
            sdg[k] = sdg[k.replace(f"{h}.{idx}.", f"{h}.0.", 1)]
t0 = time.time()
fails = 0
def cmp(name, got, ref_, tol=1e-3):
    global fails
    got = got.detach().cpu().double(); ref_ = ref_.detach().double()
    err = (got - ref_).abs().max().item(); scale = ref_.abs().max().item() + 1e-12
    ok = err / scale < tol or err < 1e-5
    fails += not ok
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs={err:.3e} rel_to_max={err/scale:.3e}", flush=True)
if a.stage == 2:
    (oo, oref), inter = OM.forward(sdg, cfg, inp["image"], rects=inp["rects"], return_intermediates=True)
    for k in ("pred_logits", "pred_boxes", "pred_vars"):
        cmp(k, out[k], oo[k])
    ol, oidx = OC.set_criterion(oo, inp["targets"])
    gidx = crit.matcher(out, targets)
    same = all(torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) for x, y in zip(gidx, oidx))
    print("OK  " if same else "FAIL", "matching indices identical to oracle:", same); fails += not same
    for k in ol:
        cmp("loss " + k, ld[k], ol[k])
    oloss = sum(ol[k] * OC.STAGE2_WEIGHT_DICT[k] for k in OC.STAGE2_WEIGHT_DICT)
else:
    oo = OM.forward(sdg, cfg, inp["image"], points=None)
    for k in ("pred_logits", "pred_wh", "pred_points"):
        cmp(k, out[k], oo[k])
    ol = OC.bounding_box_criterion(oo, {"points": inp["points"], "whs": inp["whs"]})
    for k in ol:
        cmp("loss " + k, ld[k], ol[k])
    oloss = sum(ol[k] * OC.STAGE1_WEIGHT_DICT[k] for k in ol)
oloss.backward()
print(f"oracle fwd+bwd {time.time()-t0:.2f}s")
# fp64 oracle = ground truth; ReLU-boundary flips make single elements of fp32 gradients differ between ANY two
# fp32 implementations, so judge the GPU path by its distance to fp64 next to the fp32 CPU oracle's distance.
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
sd64 = {k: v.clone().requires_grad_(sdg[k].requires_grad) if v.is_floating_point() else v for k, v in sd64.items()}
for k in list(sd64):
    for h in ("cls_embed", "bbox_embed", "bbox_variance"):
        if f"transformer.{h}." in k and f"transformer.{h}.0." not in k:
            idx = k.split(f"transformer.{h}.")[1].split(".")[0]
            sd64[k] = sd64[k.replace(f"{h}.{idx}.", f"{h}.0.", 1)]
import oracle.model as _om
_orig_arange = torch.arange
if a.stage == 2:
    (o64, _), _ = OM.forward(sd64, cfg, inp["image"].double(), rects=inp["rects"], return_intermediates=True)
    tg64 = [{"boxes": t["boxes"].double(), "labels": t["labels"]} for t in inp["targets"]]
    l64, _ = OC.set_criterion(o64, tg64, indices=oidx)
    loss64 = sum(l64[k] * OC.STAGE2_WEIGHT_DICT[k] for k in OC.STAGE2_WEIGHT_DICT)
else:
    o64 = OM.forward(sd64, cfg, inp["image"].double(), points=None)
    l64 = OC.bounding_box_criterion(o64, {"points": inp["points"].double(), "whs": inp["whs"].double()})
    loss64 = sum(l64[k] * OC.STAGE1_WEIGHT_DICT[k] for k in l64)
loss64.backward()
rows = []
for n, p in model.named_parameters():
    if not p.requires_grad:
        continue
    g64 = sd64[n].grad
    if g64 is None:
        g64 = torch.zeros_like(sd64[n])
    if p.grad is None:
        if g64.abs().max() > 0:
            print("MISSING grad", n); fails += 1
        continue
    g32 = sdg[n].grad if sdg[n].grad is not None else torch.zeros_like(sdg[n])
    nrm = g64.norm().item() + 1e-30
    e_gpu = (p.grad.cpu().double() - g64).norm().item() / nrm
    e_cpu = (g32.double() - g64).norm().item() / nrm
    rows.append((e_gpu, e_cpu, n, nrm))
rows.sort(reverse=True)
for e_gpu, e_cpu, n, nrm in rows[:15]:
    print(f"grad {n}: |gpu-f64|/|f64|={e_gpu:.3e}  |cpu32-f64|/|f64|={e_cpu:.3e}  norm={nrm:.3e}")
import statistics
print("median rel err gpu", statistics.median(r[0] for r in rows), "cpu32", statistics.median(r[1] for r in rows))
# tolerance: every tensor within 2e-2 (norm-relative) of the fp64 ground truth and the median within 1e-3.
# Gradients of a ReLU network are discontinuous in the activations: an activation within the forward error of
# zero flips its mask, so single elements differ between ANY two finite-precision runs (the fp32 CPU oracle
# itself is 1e-3..5e-3 from fp64 on the early backbone layers at these sizes).
nbad = sum(1 for e_gpu, e_cpu, *_ in rows if e_gpu > 2e-2)
med = statistics.median(r[0] for r in rows)
print("params with |gpu-f64|/|f64| > 2e-2:", nbad, "of", len(rows), " median:", med); fails += nbad + (med > 1e-3)
print("FAILS", fails)
