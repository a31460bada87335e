"""CPU tests (-m "not gpu"): the oracle against the golden vectors produced by the real reference, against
the live reference import when /root/reference is present, and the LSAP restatements against scipy."""
import os

import numpy as np
import pytest
import torch

from counting_detr_b200 import synthetic as SY
from oracle import criterion as OC, model as OM, ref_import as R

torch.set_num_threads(8)


def _close(a, b, tol):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() <= tol * (b.abs().max().item() + 1e-12) + 1e-7


def test_oracle_matches_stage2_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "stage2_S128_B2_Q50_T7.pt"))
    c = g["config"]
    cfg = OM.Config(stage=2, num_query_position=c["Q"])
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in SY.make_state_dict(cfg, c["seed"]).items()}
    for k in list(sd):   # shared heads alias index 0 (one module in the reference)
        for h in ("cls_embed", "bbox_embed", "bbox_variance"):
            if f"transformer.{h}." in k and f"transformer.{h}.0." not in k:
                i = k.split(f"transformer.{h}.")[1].split(".")[0]
                sd[k] = sd[k.replace(f"{h}.{i}.", f"{h}.0.", 1)]
    inp = SY.make_inputs(c["B"], c["S"], T=c["T"], seed=c["seed"], stage=2)
    out, ref = OM.forward(sd, cfg, inp["image"], rects=inp["rects"])
    for k in ("pred_logits", "pred_boxes", "pred_vars"):
        assert _close(out[k], g["outputs"][k], 2e-5), k
    assert torch.equal(ref, g["reference_points"])
    losses, idx = OC.set_criterion(out, inp["targets"])
    for (a, b), (ga, gb) in zip(idx, g["indices"]):
        assert torch.equal(a, ga) and torch.equal(b, gb)          # bit-exact assignment
    for k, v in g["losses"].items():
        assert _close(losses[k], v, 2e-5), k
    total = sum(losses[k] * w for k, w in OC.STAGE2_WEIGHT_DICT.items())
    assert _close(total, g["total_loss"], 2e-5)
    total.backward()
    for k, gv in g["grads"].items():
        got = sd[k].grad
        if got.shape != gv.shape:
            got = got[: gv.shape[0], : gv.shape[1]]
        assert _close(got, gv, 2e-3), k


def test_oracle_matches_stage1_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "stage1_S128_B1_Q20.pt"))
    c = g["config"]
    cfg = OM.Config(stage=1, num_query_position=c["Q"])
    sd = SY.make_state_dict(cfg, c["seed"])
    inp = SY.make_inputs(c["B"], c["S"], stage=1, Q=c["Q"], seed=c["seed"])
    out = OM.forward(sd, cfg, inp["image"])
    for k in ("pred_logits", "pred_wh", "pred_points"):
        assert _close(out[k], g["outputs"][k], 2e-5), k
    losses = OC.bounding_box_criterion(out, {"points": inp["points"], "whs": inp["whs"]})
    for k, v in g["losses"].items():
        assert _close(losses[k], v, 2e-5), k


def test_lsap_restatements_match_scipy_golden(golden_dir):
    g = torch.load(os.path.join(golden_dir, "lsap_scipy_cases.pt"))
    for case in g["cases"]:
        c = case["cost"].numpy()
        a, b = OC.lsap_c(c)
        assert np.array_equal(a, case["rows"].numpy()) and np.array_equal(b, case["cols"].numpy())
        if c.size <= 4096:
            a, b = OC.lsap_py(c)
            assert np.array_equal(a, case["rows"].numpy()) and np.array_equal(b, case["cols"].numpy())


def test_lsap_c_matches_installed_scipy():
    from scipy.optimize import linear_sum_assignment as lsa
    rng = np.random.RandomState(1)
    for nr, nc in [(300, 50), (500, 100), (37, 91), (400, 400)]:
        c = rng.rand(nr, nc)
        a, b = lsa(c); a2, b2 = OC.lsap_c(c)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
    for _ in range(200):      # tie-heavy integer costs exercise scipy's tie rule
        nr, nc = rng.randint(1, 12), rng.randint(1, 12)
        c = rng.randint(0, 3, size=(nr, nc)).astype(float)
        a, b = lsa(c); a2, b2 = OC.lsap_c(c)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)


def test_lsap_empty_and_single():
    a, b = OC.lsap_c(np.zeros((1, 1)))
    assert a.tolist() == [0] and b.tolist() == [0]


@pytest.mark.skipif(not R.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("stage", [2, 1])
def test_oracle_matches_live_reference(stage):
    Q = 60
    models = R.load(stage)
    model, crit, pp = models.build_model(R.default_args(stage, num_query_position=Q))
    cfg = OM.Config(stage=stage, num_query_position=Q)
    sd = SY.make_state_dict(cfg, 3)
    model.load_state_dict(sd, strict=True)
    model.eval()
    inp = SY.make_inputs(1, 96, T=5, seed=3, stage=stage, Q=Q)
    with torch.no_grad():
        if stage == 2:
            ro, rref = model(inp["image"], None, inp["rects"])
            oo, oref = OM.forward(sd, cfg, inp["image"], rects=inp["rects"])
            rl = crit(ro, inp["targets"]); ol, oidx = OC.set_criterion(oo, inp["targets"])
            ridx = crit.matcher(ro, inp["targets"])
            assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(ridx, oidx))
            r = pp["bbox"](ro, torch.tensor([[480.0, 640.0]]))
            o = OC.post_process(oo, torch.tensor([[480.0, 640.0]]), k=min(100, Q * 2))
            if Q * 2 >= 100:
                assert all(_close(a[k].float(), b[k].float(), 1e-5) for a, b in zip(r, o) for k in a)
        else:
            ro = model(inp["image"], inp["points"]); oo = OM.forward(sd, cfg, inp["image"])
            tg = {"points": inp["points"], "whs": inp["whs"]}
            rl = crit(ro, tg); ol = OC.bounding_box_criterion(oo, tg)
    for k in ro:
        assert _close(ro[k], oo[k], 1e-5), k
    for k in rl:
        assert _close(rl[k], ol[k], 1e-5), k


# ------------------------------------------------------------------ cases at BASELINE sizes / model variants
from oracle.make_golden import CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_case_golden(name, golden_dir):
    """The oracle against outputs of the UNMODIFIED reference (oracle/make_golden.py run_case) at BASELINE.json's own
    sizes (C1a/b, C2, C3, C4) and for the grid / defined priors, 3 query patterns and a padded NestedTensor batch.
    Tolerances: 2e-5 relative on outputs and losses (two fp32 CPU implementations of the same graph), matching
    indices bit-exact, gradients 2e-3 norm-relative per tensor."""
    from oracle.cases import compare, oracle_case
    path = os.path.join(golden_dir, name + ".pt")
    gold = torch.load(path)
    got = oracle_case(name, gold["config"]["seed"])
    fails, worst = compare(got, gold, tol_out=2e-5, tol_loss=2e-5, tol_grad_norm=2e-3, tol_grad_small=2e-3)
    assert not fails, (fails[:10], worst)
