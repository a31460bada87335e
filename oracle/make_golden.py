"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, import-time shims only)
on the portable synthetic weights/inputs.  Run in the authoring container:  python -m oracle.make_golden

TEST INFRASTRUCTURE.  The fixtures pin the oracle (and through it the CUDA path) where /root/reference
is absent.  CPU fp32, torch.set_num_threads(8) (thread count changes results by <=1.2e-6).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counting_detr_b200 import synthetic as SY  # noqa: E402
from oracle import ref_import as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
GRAD_KEYS = ["transformer.decoder_layers.5.ffn.norm2.weight", "transformer.encoder_layers.0.norm1.bias",
             "transformer.decoder_layers.0.self_attn.in_proj_bias", "transformer.adapt_pos1d.2.bias",
             "backbone.body.layer4.2.conv3.weight", "backbone.body.layer2.0.conv1.weight"]


def stage2(S=128, B=2, Q=50, T=7, seed=0):
    torch.manual_seed(0)
    m = R.load(2)
    model, crit, _ = m.build_model(R.default_args(2, num_query_position=Q))
    cfg = SY.SynthCfg(stage=2, num_query_position=Q)
    model.load_state_dict(SY.make_state_dict(cfg, seed), strict=True)
    model.train(); crit.train()
    inp = SY.make_inputs(B, S, T=T, seed=seed, stage=2)
    out, ref = model(inp["image"], None, inp["rects"])
    losses = crit(out, inp["targets"])
    idx = crit.matcher(out, inp["targets"])
    total = sum(losses[k] * crit.weight_dict[k] for k in losses if k in crit.weight_dict)
    total.backward()
    named = dict(model.named_parameters())
    g = {k: named[k].grad.clone() for k in GRAD_KEYS}
    g["backbone.body.layer4.2.conv3.weight"] = g["backbone.body.layer4.2.conv3.weight"][:8, :16].clone()
    g["backbone.body.layer2.0.conv1.weight"] = g["backbone.body.layer2.0.conv1.weight"][:8, :16].clone()
    return {"config": dict(stage=2, S=S, B=B, Q=Q, T=T, seed=seed, threads=torch.get_num_threads()),
            "outputs": {k: v.detach().clone() for k, v in out.items()}, "reference_points": ref.detach().clone(),
            "losses": {k: v.detach().clone() for k, v in losses.items()},
            "indices": [(a.clone(), b.clone()) for a, b in idx], "total_loss": total.detach().clone(), "grads": g}


def stage1(S=128, B=1, Q=20, seed=0):
    torch.manual_seed(0)
    m = R.load(1)
    model, crit, _ = m.build_model(R.default_args(1, num_query_position=Q))
    cfg = SY.SynthCfg(stage=1, num_query_position=Q)
    model.load_state_dict(SY.make_state_dict(cfg, seed), strict=True)
    model.train()
    inp = SY.make_inputs(B, S, stage=1, Q=Q, seed=seed)
    out = model(inp["image"], inp["points"])
    losses = crit(out, {"points": inp["points"], "whs": inp["whs"]})
    total = sum(losses[k] * crit.weight_dict[k] for k in losses)
    total.backward()
    named = dict(model.named_parameters())
    g = {k: named[k].grad.clone() for k in GRAD_KEYS[:4]}
    return {"config": dict(stage=1, S=S, B=B, Q=Q, seed=seed, threads=torch.get_num_threads()),
            "outputs": {k: v.detach().clone() for k, v in out.items()},
            "losses": {k: v.detach().clone() for k, v in losses.items()}, "total_loss": total.detach().clone(), "grads": g}


# ---------------------------------------------------------------------------------------------------------
# Cases at BASELINE.json's own sizes and for the model variants the reference supports (spatial priors, query
# patterns, padded NestedTensor batches).  One table drives the generator, the CPU oracle tests and the GPU tests.
# `sizes` (padded case): per-image [h, w] crops of the S x S synthetic images, batched with the reference's own
# nested_tensor_from_tensor_list.  `train`: also run criterion + backward and store gradient fingerprints.
# ---------------------------------------------------------------------------------------------------------
CASES = {
    # BASELINE configs at their own size
    "c3_stage2_S512_B16_Q300": dict(stage=2, S=512, B=16, Q=300, T=50, prior="learned", P=1, train=True),
    "c2_stage1_S512_B8_Q300": dict(stage=1, S=512, B=8, Q=300, prior="learned", P=1, train=True),
    "c4_stage2_S800_B2_Q500": dict(stage=2, S=800, B=2, Q=500, T=50, prior="learned", P=1, train=True),
    "c1a_stage1_defined_Q3": dict(stage=1, S=512, B=1, Q=3, prior="defined", P=1, train=True),
    "c1b_stage1_defined_Q50_fwd": dict(stage=1, S=512, B=1, Q=50, prior="defined", P=1, train=False),
    # model variants
    "grid_stage2_S256_B2_Q289": dict(stage=2, S=256, B=2, Q=300, T=20, prior="grid", P=1, train=True),
    "pattern3_stage2_S256_B2_Q3x100": dict(stage=2, S=256, B=2, Q=100, T=20, prior="learned", P=3, train=True),
    "padded_stage2_S256_B2_Q100": dict(stage=2, S=256, B=2, Q=100, T=20, prior="learned", P=1, train=True,
                                       sizes=[[256, 192], [224, 256]]),
    "defined_stage2_S256_B1_Q40": dict(stage=2, S=256, B=1, Q=40, T=20, prior="defined", P=1, train=True),
}
PROBE_SEED = 99          # fingerprint vectors: SY.uniform(name, shape, -1, 1, PROBE_SEED)
FULL_GRAD_MAX = 4096     # parameters up to this size are stored whole


def case_inputs(c, seed=0):
    """Synthetic inputs of a case (shared with tests/): dict with image or images (+ sizes), rects, targets/points."""
    inp = SY.make_inputs(c["B"], c["S"], T=c.get("T", 50), seed=seed, stage=c["stage"], Q=c["Q"])
    if c.get("sizes"):
        inp["images"] = [inp["image"][b, :, :h, :w].contiguous() for b, (h, w) in enumerate(c["sizes"])]
    if c["prior"] == "defined" and c["stage"] == 2:
        inp["points"] = SY.uniform("pts", (c["Q"], 2), 0.1, 0.9, seed)      # ndarray [Q,2] in the reference's call
    return inp


def grad_fingerprints(named_grads):
    """name -> (norm, dot with a fixed pseudo-random probe); small tensors also whole."""
    fp, full = {}, {}
    for k, g in named_grads.items():
        probe = SY.uniform("probe." + k, tuple(g.shape), -1.0, 1.0, PROBE_SEED)
        fp[k] = torch.tensor([g.double().norm().item(), (g.double() * probe.double()).sum().item()], dtype=torch.float64)
        if g.numel() <= FULL_GRAD_MAX:
            full[k] = g.clone()
    return fp, full


def run_case(name, seed=0):
    c = CASES[name]
    st = c["stage"]
    torch.manual_seed(0)
    torch.Tensor.cuda = lambda self, *a, **k: self       # grid / defined priors call .cuda() (transformer.py:122,129)
    m = R.load(st)
    args = R.default_args(st, num_query_position=c["Q"], num_query_pattern=c["P"], spatial_prior=c["prior"])
    model, crit, _ = m.build_model(args)
    cfg = SY.SynthCfg(stage=st, num_query_position=c["Q"], num_query_pattern=c["P"], spatial_prior=c["prior"])
    model.load_state_dict(SY.make_state_dict(cfg, seed), strict=True)
    model.train(); crit.train()
    inp = case_inputs(c, seed)
    samples = inp["images"] if c.get("sizes") else inp["image"]
    res = {"config": dict(c, name=name, seed=seed, threads=torch.get_num_threads())}
    if st == 2:
        pts = inp["points"].numpy() if c["prior"] == "defined" else None
        out, ref = model(samples, pts, inp["rects"])
        res["reference_points"] = ref.detach().clone()
        targets = inp["targets"]
    else:
        out = model(samples, inp["points"])
        targets = {"points": inp["points"], "whs": inp["whs"]}
    res["outputs"] = {k: v.detach().clone() for k, v in out.items()}
    if c["train"]:
        losses = crit(out, targets)
        if st == 2:
            res["indices"] = [(a.clone(), b.clone()) for a, b in crit.matcher(out, targets)]
        total = sum(losses[k] * crit.weight_dict[k] for k in losses if k in crit.weight_dict)
        total.backward()
        res["losses"] = {k: v.detach().clone() for k, v in losses.items()}
        res["total_loss"] = total.detach().clone()
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        res["grad_fp"], res["grads_small"] = grad_fingerprints(grads)
    return res


def matcher_cases():
    """cost matrices (fp32, as the reference hands them to scipy) + scipy's own answers: random, ragged,
    duplicated-target ties, small-integer ties."""
    from scipy.optimize import linear_sum_assignment as lsa
    rng = np.random.RandomState(0)
    cases = []
    for (nr, nc) in [(300, 50), (50, 300), (64, 64), (1, 9), (9, 1), (120, 7)]:
        c = rng.rand(nr, nc).astype(np.float32)
        cases.append(c)
    c = rng.rand(100, 20).astype(np.float32); c[:, 10:] = c[:, :10]; cases.append(c)        # duplicated targets
    c = rng.rand(30, 30).astype(np.float32); c[15:] = c[:15]; cases.append(c)                # duplicated queries
    for _ in range(12):
        nr, nc = rng.randint(1, 14), rng.randint(1, 14)
        cases.append(rng.randint(0, 4, size=(nr, nc)).astype(np.float32))
    return [{"cost": torch.from_numpy(c), "rows": torch.from_numpy(lsa(c)[0].astype(np.int64)),
             "cols": torch.from_numpy(lsa(c)[1].astype(np.int64))} for c in cases]


if __name__ == "__main__":
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if not only:
        torch.save(stage2(), os.path.join(OUT, "stage2_S128_B2_Q50_T7.pt"))
        torch.save(stage1(), os.path.join(OUT, "stage1_S128_B1_Q20.pt"))
        import scipy
        torch.save({"scipy": scipy.__version__, "cases": matcher_cases()}, os.path.join(OUT, "lsap_scipy_cases.pt"))
    for name in CASES:
        if only and name not in only:
            continue
        import time
        t0 = time.time()
        torch.save(run_case(name), os.path.join(OUT, name + ".pt"))
        print(f"{name}: {time.time() - t0:.1f} s", flush=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
