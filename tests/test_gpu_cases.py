"""GPU parity at BASELINE.json's own sizes and for every model variant the reference supports (-m gpu).

Each case of oracle.make_golden.CASES runs through the public API of the CUDA path (build_model -> model -> criterion
-> backward) and is compared with the fixture the UNMODIFIED reference produced for the same seeded weights / inputs
(tests/golden/<case>.pt, generated in the authoring container by oracle/make_golden.py):
  outputs and every loss within 1e-3 relative (the north-star tolerance), matching indices bit-exact, reference
  points bit-exact, per-parameter gradient norms within 2e-2 and every parameter of <= 4096 elements within 2e-2
  norm-relative (ReLU-boundary flips bound the gradient agreement of two fp32 implementations; tests/gpu_model_probe.py
  triangulates that against fp64).
A second test repeats C3's shapes on other seeds against the CPU oracle run on the spot (no fixture involved).
"""
import gc
import json
import os

import pytest
import torch

from oracle.make_golden import CASES, case_inputs, grad_fingerprints

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_cases.jsonl")


def cuda_case(name, seed=0, B=None):
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model
    c = dict(CASES[name])
    if B is not None:
        c["B"] = B
    st = c["stage"]
    args = SY.default_args(st, num_query_position=c["Q"], num_query_pattern=c["P"], spatial_prior=c["prior"], device="cuda")
    model, crit, _ = build_model(args)
    cfg = SY.SynthCfg(stage=st, num_query_position=c["Q"], num_query_pattern=c["P"], spatial_prior=c["prior"])
    model.load_state_dict(SY.make_state_dict(cfg, seed), strict=True)
    model.cuda().train(); crit.train()
    inp = case_inputs(c, seed)
    samples = [im.cuda() for im in inp["images"]] if c.get("sizes") else inp["image"].cuda()
    res = {"config": dict(c, name=name, seed=seed)}
    if st == 2:
        pts = inp["points"].numpy() if c["prior"] == "defined" else None
        out, ref = model(samples, pts, inp["rects"].cuda())          # rects on the device, as the reference's engine passes them
        res["reference_points"] = ref.detach().cpu()
        targets = [{k: v.cuda() for k, v in t.items()} for t in inp["targets"]]
    else:
        out = model(samples, inp["points"].cuda())
        targets = {"points": inp["points"].cuda(), "whs": inp["whs"].cuda()}
    res["outputs"] = {k: v.detach().cpu() for k, v in out.items()}
    if c["train"]:
        losses = crit(out, targets)
        total = sum(losses[k] * crit.weight_dict[k] for k in losses if k in crit.weight_dict)
        total.backward()
        if st == 2:
            oq, ot, on = [t.cpu() for t in crit.last_indices]         # the assignment the loss itself used
            res["indices"] = [(oq[b, : on[b]], ot[b, : on[b]]) for b in range(c["B"])]
        res["losses"] = {k: v.detach().cpu() for k, v in losses.items()}
        res["total_loss"] = total.detach().cpu()
        grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
        res["grad_fp"], res["grads_small"] = grad_fingerprints(grads)
    del model, crit
    gc.collect(); torch.cuda.empty_cache()
    return res


def _report(kind, name, worst, fails):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps({"check": kind, "case": name, "worst_rel_err": {k: float(f"{v:.3e}") for k, v in worst.items()},
                            "fails": fails[:5]}) + "\n")


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_golden_at_size(name, golden_dir):
    from oracle.cases import compare
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    got = cuda_case(name, gold["config"]["seed"])
    fails, worst = compare(got, gold, tol_out=1e-3, tol_loss=1e-3, tol_grad_norm=2e-2, tol_grad_small=2e-2)
    _report("cuda_vs_reference_golden", name, worst, fails)
    assert not fails, (fails[:10], worst)


@pytest.mark.parametrize("name,B,seed", [("c3_stage2_S512_B16_Q300", 4, 1), ("c3_stage2_S512_B16_Q300", 4, 2),
                                         ("c4_stage2_S800_B2_Q500", 1, 3), ("c2_stage1_S512_B8_Q300", 2, 4)])
def test_cuda_matches_oracle_other_seeds(name, B, seed):
    """Same shapes per image as C3 / C4 / C2, other seeds (weights AND inputs), against the CPU oracle run here."""
    from oracle import cases as OCS
    saved = dict(CASES[name])
    CASES[name]["B"] = B
    try:
        gold = OCS.oracle_case(name, seed)
        got = cuda_case(name, seed)
    finally:
        CASES[name].update(saved)
    fails, worst = OCS.compare(got, gold, tol_out=1e-3, tol_loss=1e-3, tol_grad_norm=2e-2, tol_grad_small=2e-2)
    _report("cuda_vs_oracle", f"{name}[B={B},seed={seed}]", worst, fails)
    assert not fails, (fails[:10], worst)


def test_padding_mask_changes_the_result(golden_dir):
    """The padded case must really exercise the mask: the same padded pixels WITHOUT the mask give different logits."""
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model, _pad_images
    name = "padded_stage2_S256_B2_Q100"
    c = CASES[name]
    model, _, _ = build_model(SY.default_args(2, num_query_position=c["Q"], device="cuda"))
    model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=2, num_query_position=c["Q"]), 0), strict=True)
    model.cuda().eval()
    inp = case_inputs(c, 0)
    imgs = [im.cuda() for im in inp["images"]]
    with torch.no_grad():
        a, _ = model(imgs, None, inp["rects"])
        padded, mask = _pad_images(imgs)
        b, _ = model(padded, None, inp["rects"])
    assert mask is not None and mask.any()
    assert (a["pred_logits"] - b["pred_logits"]).abs().max().item() > 1e-3


def test_bad_exemplar_rect_poisons_the_loss():
    """A rect whose centre leaves the feature map raises IndexError in the reference; here (no host read) the losses
    turn NaN, which the reference's loop treats as fatal, and nothing is read out of bounds."""
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model
    model, crit, _ = build_model(SY.default_args(2, num_query_position=50, device="cuda"))
    model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=2, num_query_position=50), 0), strict=True)
    model.cuda().train()
    inp = SY.make_inputs(2, 128, T=7, stage=2)
    targets = [{k: v.cuda() for k, v in t.items()} for t in inp["targets"]]
    rects = inp["rects"].clone()
    rects[0, 1] = torch.tensor([0.9, 0.9, 1.3, 1.2])
    out, _ = model(inp["image"].cuda(), None, rects.cuda())
    ld = crit(out, targets)
    assert all(torch.isnan(v).item() for v in ld.values())
    out, _ = model(inp["image"].cuda(), None, inp["rects"].cuda())      # flag re-armed: the next step is clean
    ld = crit(out, targets)
    assert all(torch.isfinite(v).item() for v in ld.values())


def test_cuda_core_rcda_fallback_still_matches_c4_golden(golden_dir, monkeypatch):
    """CDETR_RCDA_LEGACY=1: the CUDA-core RCDA kernels (rcda.cu; maps beyond 64 x 64 would use them) on C4's shapes."""
    from oracle.cases import compare
    monkeypatch.setenv("CDETR_RCDA_LEGACY", "1")
    name = "c4_stage2_S800_B2_Q500"
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    got = cuda_case(name, gold["config"]["seed"])
    fails, worst = compare(got, gold, tol_out=1e-3, tol_loss=1e-3, tol_grad_norm=2e-2, tol_grad_small=2e-2)
    _report("cuda_vs_reference_golden[rcda_legacy]", name, worst, fails)
    assert not fails, (fails[:10], worst)
