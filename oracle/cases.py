"""Run one entry of oracle.make_golden.CASES through the CPU oracle and compare result dicts.

TEST INFRASTRUCTURE (tests/, __graft_entry__.smoke() and bench.py's CPU arm only).  `oracle_case` produces the same
dict layout as oracle.make_golden.run_case (which runs the UNMODIFIED reference), so one comparator serves
reference-vs-oracle (CPU suite) and CUDA-vs-reference-golden / CUDA-vs-oracle (GPU suite).
"""
import torch

from counting_detr_b200 import synthetic as SY
from oracle import criterion as OC, model as OM
from oracle.make_golden import CASES, case_inputs, grad_fingerprints

HEADS = ("cls_embed", "bbox_embed", "bbox_variance")


def oracle_state_dict(cfg, seed, requires_grad=True):
    """Synthetic state dict with the shared heads aliased to index 0 (one module object in the reference,
    A2/models/transformer.py:104-107) and the reference's requires_grad pattern (A2/models/backbone.py:93-95)."""
    frozen = ("backbone.body.conv1", "backbone.body.layer1", "running_", ".bn", "downsample.1")
    sd = {}
    for k, v in SY.make_state_dict(cfg, seed).items():
        sd[k] = v.clone().requires_grad_(requires_grad and v.is_floating_point() and not any(f in k for f in frozen))
    for k in list(sd):
        for h in HEADS:
            if f"transformer.{h}." in k and f"transformer.{h}.0." not in k:
                i = k.split(f"transformer.{h}.")[1].split(".")[0]
                sd[k] = sd[k.replace(f"{h}.{i}.", f"{h}.0.", 1)]
    return sd


def oracle_case(name, seed=0):
    c = CASES[name]
    st = c["stage"]
    cfg = OM.Config(stage=st, num_query_position=c["Q"], num_query_pattern=c["P"], spatial_prior=c["prior"])
    sd = oracle_state_dict(SY.SynthCfg(stage=st, num_query_position=c["Q"], num_query_pattern=c["P"],
                                       spatial_prior=c["prior"]), seed, c["train"])
    inp = case_inputs(c, seed)
    image, image_mask = inp["image"], None
    if c.get("sizes"):
        image, image_mask = OM.pad_images(inp["images"])
    res = {"config": dict(c, name=name, seed=seed)}
    if st == 2:
        out, ref = OM.forward(sd, cfg, image, rects=inp["rects"], points=inp.get("points"), image_mask=image_mask)
        res["reference_points"] = ref.detach()
    else:
        out = OM.forward(sd, cfg, image, points=inp["points"] if c["prior"] == "defined" else None,
                         image_mask=image_mask)
    res["outputs"] = {k: v.detach() for k, v in out.items()}
    if c["train"]:
        if st == 2:
            losses, idx = OC.set_criterion(out, inp["targets"])
            res["indices"] = idx
            wd = OC.STAGE2_WEIGHT_DICT
        else:
            losses = OC.bounding_box_criterion(out, {"points": inp["points"], "whs": inp["whs"]})
            wd = OC.STAGE1_WEIGHT_DICT
        total = sum(losses[k] * w for k, w in wd.items())
        total.backward()
        res["losses"] = {k: v.detach() for k, v in losses.items()}
        res["total_loss"] = total.detach()
        # shared heads: one gradient under index 0 (named_parameters() of the reference dedups the aliases)
        alias = lambda k: any(f"transformer.{h}." in k and f"transformer.{h}.0." not in k for h in HEADS)
        res["grad_fp"], res["grads_small"] = grad_fingerprints({k: v.grad for k, v in sd.items()
                                                                if v.grad is not None and not alias(k)})
    return res


def compare(got, gold, tol_out=1e-3, tol_loss=1e-3, tol_grad_norm=2e-2, tol_grad_small=2e-2, min_grad=1e-12):
    """Returns a list of human-readable failures (empty = parity).  Tolerances are relative: outputs to the golden
    tensor's max magnitude, losses to their value, gradients norm-relative per tensor.  Indices bit-exact."""
    fails, worst = [], {}
    for k, gv in gold["outputs"].items():
        v = got["outputs"][k].detach().cpu().double()
        err = (v - gv.double()).abs().max().item() / (gv.double().abs().max().item() + 1e-30)
        worst["out." + k] = err
        if not err <= tol_out:
            fails.append(f"output {k}: rel err {err:.3e} > {tol_out}")
    if "reference_points" in gold and "reference_points" in got:
        if not torch.equal(got["reference_points"].cpu(), gold["reference_points"]):
            fails.append("reference_points differ")
    if "indices" in gold:
        for b, ((a, bb), (ga, gb)) in enumerate(zip(got["indices"], gold["indices"])):
            if not (torch.equal(torch.as_tensor(a).cpu(), ga) and torch.equal(torch.as_tensor(bb).cpu(), gb)):
                fails.append(f"matching indices differ for image {b}")
    for k, gv in gold.get("losses", {}).items():
        v = float(got["losses"][k])
        err = abs(v - float(gv)) / (abs(float(gv)) + 1e-30)
        worst["loss." + k] = err
        if not (err <= tol_loss or abs(v - float(gv)) <= 1e-6):
            fails.append(f"loss {k}: {v} vs {float(gv)} (rel {err:.3e})")
    if "grad_fp" in gold and "grad_fp" in got:
        gnorm_err = gdot_err = 0.0
        for k, fp in gold["grad_fp"].items():
            if k not in got["grad_fp"]:
                if fp[0].item() > min_grad:
                    fails.append(f"gradient of {k} missing")
                continue
            n_ref, d_ref = fp[0].item(), fp[1].item()
            n_got, d_got = got["grad_fp"][k][0].item(), got["grad_fp"][k][1].item()
            if n_ref <= min_grad:
                continue
            e_norm = abs(n_got - n_ref) / n_ref
            # probe ~ U(-1,1)^n: (g_got - g_ref) . probe ~ |g_got - g_ref| / sqrt(3) for an error of random direction, so the
            # dot difference estimates the norm of the difference (direction check for the tensors not stored whole)
            e_dot = abs(d_got - d_ref) * 3 ** 0.5 / n_ref
            gnorm_err = max(gnorm_err, e_norm)
            gdot_err = max(gdot_err, e_dot)
            if not e_norm <= tol_grad_norm:
                fails.append(f"grad norm {k}: {n_got:.6e} vs {n_ref:.6e} (rel {e_norm:.3e})")
            elif not e_dot <= 4 * tol_grad_norm:
                fails.append(f"grad probe {k}: dot {d_got:.6e} vs {d_ref:.6e} (est. rel diff {e_dot:.3e})")
        worst["grad.probe"] = gdot_err
        worst["grad.norm"] = gnorm_err
        gsmall = 0.0
        for k, gv in gold.get("grads_small", {}).items():
            if k not in got.get("grads_small", {}) or gv.double().norm().item() <= min_grad:
                continue
            v = got["grads_small"][k].detach().cpu().double().reshape(gv.shape)
            e = (v - gv.double()).norm().item() / gv.double().norm().item()
            gsmall = max(gsmall, e)
            if not e <= tol_grad_small:
                fails.append(f"grad {k}: norm-relative error {e:.3e} > {tol_grad_small}")
        worst["grad.small"] = gsmall
    return fails, worst
