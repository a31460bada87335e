#!/usr/bin/env python
"""bench.py — images/sec of one Counting-DETR train step (512x512, 300 queries) on N B200s.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload c3|c2|c4] [--no-graph]

A "step" = zero_grad -> model forward -> criterion (device Hungarian matching + losses) -> backward
(-> gradient all-reduce over NCCL when N > 1) on one batch of synthetic input.  Default
workload is BASELINE.json's stage-2 config C3 (B=16/GPU, S=512, Q=300, T=50 targets/image): it is the
configuration the north star quotes the attention target on and it exercises the whole path
(exemplar injection, RCDA encoder/decoder, matcher, Laplace loss).

One JSON line on stdout (rank 0): value = device-timed whole-job images/s with inputs resident in HBM
(CUDA-graph replay of the step unless --no-graph); e2e = the same step driven through the public
build_model()/criterion API from pinned HOST buffers with the H2D copies and the loss D2H read inside
the timed region; roofline = algorithmic GEMM FLOPs / measured GEMM time of one instrumented step;
cpu_baseline = the oracle port (CPU PyTorch restatement of the reference) on the host cores.
`--impl reference` prints the CPU arm alone (the reference itself cannot travel to the GPU box).
"""
import argparse
import itertools
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (stage, B per GPU, S, Q, T)
    "c3": (2, 16, 512, 300, 50),
    "c2": (1, 8, 512, 300, 300),
    "c4": (2, 8, 800, 500, 50),
    "tiny": (2, 2, 128, 50, 7),
}
METRIC = "images/sec train step (512x512, 300 queries)"


def workload_desc(name):
    st, B, S, Q, T = WORKLOADS[name]
    return (f"{name.upper()}: stage-{st} train step, B={B}/GPU, {S}x{S} synthetic RGB, Q={Q} learned queries"
            + (f", 3 exemplar boxes, T={T} target boxes/image" if st == 2 else f", {T} point/wh targets/image"))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(name, steps, warmup, batch=None):
    """The oracle port (oracle/model.py + oracle/criterion.py: CPU PyTorch fp32 restatement of the reference,
    pinned against the live reference import and its goldens) timed on the host cores: fwd + loss + bwd, at the
    workload's own per-GPU batch size unless `batch` says otherwise."""
    import torch
    from oracle import cases as OCS, criterion as OC, model as OM, weights as OW
    st, B, S, Q, T = WORKLOADS[name]
    batch = batch or B
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = OM.Config(stage=st, num_query_position=Q)
    from counting_detr_b200 import synthetic as SY
    sd = OCS.oracle_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0)
    inp = OW.make_inputs(batch, S, T=T, stage=st, Q=Q)

    def step():
        for v in sd.values():
            v.grad = None
        if st == 2:
            out, _ = OM.forward(sd, cfg, inp["image"], rects=inp["rects"])
            ld, _ = OC.set_criterion(out, inp["targets"])
            loss = sum(ld[k] * w for k, w in OC.STAGE2_WEIGHT_DICT.items())
        else:
            out = OM.forward(sd, cfg, inp["image"])
            ld = OC.bounding_box_criterion(out, {"points": inp["points"], "whs": inp["whs"]})
            loss = sum(ld[k] * w for k, w in OC.STAGE1_WEIGHT_DICT.items())
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": batch / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of B={batch} (the GPU workload's per-GPU batch, same shapes), {warmup} warm-up, "
                      f"fwd+criterion+bwd, torch CPU fp32 on {cores} threads, {dt * 1e3:.0f} ms/step"}, dt


def run_reference(args):
    """Reference arm: the reference's CPU PyTorch path (= the oracle port: /root/reference itself cannot travel to the
    GPU box and is not pip-installable) on the same workload, batch size, steps and warm-up as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_arm(args.workload, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload), "global_batch": WORKLOADS[args.workload][1],
                       "note": "reference CPU path = oracle port (pinned to the live reference and its goldens); one "
                               "process on the host cores whatever --gpus says"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ attention roofline
def attention_macs(name, a):
    """Algorithmic MACs of one attention-core call (RCDA core = L*W*E + L*H*E + L*H*W*E +
    L*min(H,W)*E per sample forward, backward = 2x split over the query / value / key kernels; MHA core = 2*L^2*E)."""
    if "mha" in name:
        B, Lq, E = a[0], a[1], a[2]
        return B * 2 * Lq * Lq * E * (2 if "bwd" in name else 1)
    B, Lq, H, W, E = a[:5]
    logits, contract = Lq * (W + H) * E, Lq * H * W * E + Lq * min(H, W) * E
    if name in ("cdetr_rcda_fwd", "cdetr_rcda_fwd_tc"):
        return B * (logits + contract)
    if name == "cdetr_rcda_bwd_q_tc":      # G = dO V^T, dA_r/dA_c, dq = dS K
        return B * (contract + logits)
    if name == "cdetr_rcda_bwd_v_tc":      # dV = sum_q A_c A_r dO
        return B * contract
    if name == "cdetr_rcda_bwd_k":         # dK = dS^T q
        return B * logits
    if name == "cdetr_rcda_bwd_kv":
        return B * (contract + logits)
    if name == "cdetr_rcda_bwd":
        return B * 2 * (logits + contract)
    return 0


def attention_roofline(atrace, peak, ms_step):
    """North-star figure: the encoder-decoder attention cores (RCDA fwd/bwd + decoder self-attention) against the
    tensor-pipe roofline.  Each call timed alone with CUDA events (same instrumented step as the GEMM trace)."""
    fam = {}
    for name, a, e0, e1 in atrace:
        f = fam.setdefault(name.replace("cdetr_", ""), [0, 0.0, 0.0])
        f[0] += 1
        f[1] += e0.elapsed_time(e1)
        f[2] += 2.0 * attention_macs(name, a)
    tot_ms = sum(f[1] for f in fam.values())
    tot_fl = sum(f[2] for f in fam.values())
    ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    return {"bound": "tensor", "kernel": "attention cores (rcda_fwd_tc, rcda_bwd_{q,v}_tc, rcda_bwd_k, mha_fwd/bwd)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "algorithmic_gflop_per_step": tot_fl / 1e9, "ms_per_step": tot_ms, "share_of_step": tot_ms / ms_step,
            "per_kernel": {k: {"launches": f[0], "ms": round(f[1], 4), "tflops": round(f[2] / (f[1] * 1e-3) / 1e12, 2) if f[1] > 0 else None}
                           for k, f in sorted(fam.items())},
            "note": "algorithmic fp32-equivalent FLOPs; tensor-core kernels issue 3 bf16 MMAs per product (split operands)"}


# ------------------------------------------------------------------------------------------ matcher metric
def matcher_bench(dev):
    """BASELINE.json's second metric: matcher us/image (cost matrix + assignment, device-resident indices), with scipy
    on the same host's cores beside it and an indices-equal flag (scipy on the cost matrix the oracle computes)."""
    import numpy as np
    import torch
    from scipy.optimize import linear_sum_assignment as lsa
    from counting_detr_b200.models import HungarianMatcher
    from counting_detr_b200 import synthetic as SY
    from oracle import criterion as OC
    out = {}
    for tag, B, Q, T, reps in (("16x300x50", 16, 300, 50, 20), ("1x1000x1000", 1, 1000, 1000, 3), ("148x1000x1000", 148, 1000, 1000, 2)):
        m = HungarianMatcher(2.0, 5.0, 2.0)
        logits = SY.uniform("m_logits", (B, Q, 2), -2.0, 2.0, 5)
        boxes = torch.cat([SY.uniform("m_c", (B, Q, 2), 0.0, 1.0, 5), SY.uniform("m_s", (B, Q, 2), 0.01, 0.21, 5)], -1)
        tb = [torch.cat([SY.uniform(f"m_tc{b}", (T, 2), 0.0, 1.0, 5), SY.uniform(f"m_ts{b}", (T, 2), 0.01, 0.21, 5)], -1) for b in range(B)]
        lg, bx = logits.to(dev), boxes.to(dev)
        tg = [{"boxes": t.to(dev), "labels": torch.zeros(T, dtype=torch.int64, device=dev)} for t in tb]
        oq, ot, on, _ = m.match_device(lg, bx, tg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            oq, ot, on, _ = m.match_device(lg, bx, tg)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps / B
        oq, ot = oq.cpu(), ot.cpu()
        nchk = min(B, 3)
        t_sc, equal = 0.0, True
        for b in range(nchk):
            c = OC.match_cost(logits[b], boxes[b], tb[b]).numpy()
            t0 = time.perf_counter()
            i, j = lsa(c)
            t_sc += time.perf_counter() - t0
            equal &= bool(np.array_equal(oq[b].numpy(), i) and np.array_equal(ot[b].numpy(), j))
        out[tag] = {"us_per_image": round(us, 2), "latency_ms_per_call": round(us * B / 1e3, 3),
                    "scipy_us_per_image_same_host": round(t_sc / nchk * 1e6, 1), "indices_equal_scipy": equal,
                    "images_checked": nchk}
    out["note"] = ("cost matrix + LSAP kernels, device-resident indices, CUDA events over back-to-back calls; scipy = "
                   "linear_sum_assignment alone (cost already on the host) on this box's CPU, 1 thread")
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from counting_detr_b200 import _lib as L
    from counting_detr_b200.data import DevicePrefetcher
    from counting_detr_b200.models import build_model
    from counting_detr_b200.optim import FusedAdamW
    from counting_detr_b200.step import CapturedStep
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.parallel import shard_seed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    st, B, S, Q, T = WORKLOADS[args.workload]
    steps, warm = args.steps, max(args.warmup, 3)

    def build():
        margs = SY.default_args(st, num_query_position=Q, device=str(dev))
        model, crit, _ = build_model(margs)
        model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
        model.to(dev).train(); crit.train()
        return model, crit

    model, crit = build()
    # two pinned host batches (alternating: the static device inputs really change between steps)
    host = []
    for i in range(2):
        inp = SY.make_inputs(B, S, T=T, seed=shard_seed(i, rank), stage=st, Q=Q)
        hb = {"image": inp["image"].pin_memory()}
        if st == 2:
            hb["rects"] = inp["rects"].pin_memory()
            hb["targets"] = [{"boxes": t["boxes"].pin_memory(), "labels": t["labels"]} for t in inp["targets"]]
        else:
            hb["points"] = inp["points"].pin_memory()
            hb["targets"] = {"points": hb["points"], "whs": inp["whs"].pin_memory()}
        host.append(hb)
    devb = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else
                 ([{kk: vv.to(dev) for kk, vv in t.items()} for t in v] if isinstance(v, list) else
                  {kk: vv.to(dev) for kk, vv in v.items()})) for k, v in hb.items()} for hb in host]

    def call(stepper, b):
        if st == 2:
            return stepper(b["image"], b["targets"], rects=b["rects"])
        return stepper(b["image"], b["targets"], points=b["points"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / n

    # ---- (1) the metric: fwd + criterion + bwd (+ gradient all-reduce), CUDA-graph replay, inputs resident in HBM
    step_fb = CapturedStep(model, crit, optimizer=None, group=group, use_graph=not args.no_graph)
    for i in range(warm):
        out = call(step_fb, devb[i & 1])
    torch.cuda.synchronize()
    loss_val = float(out[1])
    use_graph = step_fb._graph is not None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step = timed(lambda i: call(step_fb, devb[i & 1]), steps)
    clocks = sampler.stop() if rank == 0 else None
    value = B * world / ms_step * 1e3
    launches_per_step = step_fb.launches_per_step
    if launches_per_step is None:
        L.COUNTER["launches"] = 0
        call(step_fb, devb[0])
        launches_per_step = L.COUNTER["launches"]

    # ---- (2) e2e: the same step through the public API from pinned HOST batches: counting_detr_b200.data.DevicePrefetcher
    # (H2D of batch i+1 on a copy stream under the compute of batch i) -> CapturedStep -> loss.item() EVERY step (the
    # reference's loop reads the loss each iteration, A2/engine.py:44); all of it inside the timed region
    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    class E2ELoop:
        """The reference's iteration over a data loader, through the public API.  One DevicePrefetcher over an endless
        stream of pinned host batches (a real epoch builds it once for hundreds of iterations: its construction - pinned
        staging buffers, copy stream - is outside the timed region, every step's H2D copy and result read are inside).
        lagged=False: loss.item() right after every step (A2/engine.py:44).  lagged=True: the loss of step i leaves through
        an async copy into pinned memory and is read while step i+1 runs (still one read per step, the last one inside the
        timed region): the host never leaves the device idle."""

        def __init__(self, stepper, lagged):
            self.stepper, self.lagged, self.k = stepper, lagged, 0
            self.pf = DevicePrefetcher((host[i & 1] for i in itertools.count()), dev, defer=not lagged)
            self.it = iter(self.pf)

        def run(self, n):
            last = None
            for _ in range(n):
                b = next(self.it)
                _, total = call(self.stepper, b)
                if not self.lagged:
                    self.pf.kick()       # next batch's H2D is issued after this step's launch, before the blocking read
                    last = total.item()
                    continue
                slot = self.k & 1
                loss_pin[slot].copy_(total, non_blocking=True)
                loss_ev[slot].record()
                if self.k >= 1:
                    loss_ev[slot ^ 1].synchronize()
                    last = float(loss_pin[slot ^ 1])
                self.k += 1
            if self.lagged and n >= 1:
                loss_ev[(self.k - 1) & 1].synchronize()
                last = float(loss_pin[(self.k - 1) & 1])
            return last

        def timed(self, n):
            self.run(3)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0 = self.pf.h2d_bytes
            e0.record()
            self.run(n)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / n
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            h2d = (self.pf.h2d_bytes - h0) // n
            self.it.close()
            return ms, h2d

    ms_e2e, h2d_step = E2ELoop(step_fb, lagged=False).timed(steps)
    ms_e2e_lag, _ = E2ELoop(step_fb, lagged=True).timed(steps)
    e2e = {"value": B * world / ms_e2e * 1e3, "unit": "images/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": 4,
           "path": "DevicePrefetcher(pinned host batches) -> CapturedStep(model, criterion) -> loss.item() every step",
           "value_async_loss_read": B * world / ms_e2e_lag * 1e3,
           "async_note": "same path, but the loss of step i is read (from pinned memory) while step i+1 runs"}

    # ---- (3) the whole iteration of the reference's loop: + clip_grad_norm_(0.1) + AdamW (fused) + weight re-pack
    full = None
    if not args.no_full:
        model2, crit2 = build()
        groups = [{"params": [p for n, p in model2.named_parameters() if "backbone" not in n and p.requires_grad], "lr": 1e-4},
                  {"params": [p for n, p in model2.named_parameters() if "backbone" in n and p.requires_grad], "lr": 1e-5}]
        opt = FusedAdamW(groups, lr=1e-4, weight_decay=1e-4)
        step_full = CapturedStep(model2, crit2, optimizer=opt, max_norm=0.1, group=group, use_graph=not args.no_graph)
        for i in range(warm):
            call(step_full, devb[i & 1])
        ms_full = timed(lambda i: call(step_full, devb[i & 1]), steps)
        eng2 = model2.engine()
        ms_opt = timed(lambda i: opt.step(max_norm=0.1), 10)
        ms_pack = timed(lambda i: eng2.pack_weights(), 10)
        ms_full_e2e, _ = E2ELoop(step_full, lagged=False).timed(steps)
        full = {"ms_per_step": ms_full, "value": B * world / ms_full * 1e3, "e2e_value": B * world / ms_full_e2e * 1e3,
                "optimizer_plus_repack_ms_in_graph": ms_full - ms_step,
                "optimizer_ms_eager": ms_opt, "repack_ms_eager": ms_pack, "launches_per_step": step_full.launches_per_step,
                "what": "fwd + criterion + bwd (+ all-reduce) + fused clip_grad_norm_(0.1) + AdamW (2 LR groups) + re-pack "
                        "of the updated weights into split-bf16 operands, one CUDA graph; *_eager = those two parts "
                        "launched eagerly on their own (host-bound: ~180 launches / a 265-tensor Python loop)"}
        del step_full, model2, crit2, opt

    # ---- (4) eager launches through model()/criterion() (what the reference's unmodified engine.py drives)
    eager = None
    if use_graph and not args.no_full:
        def eager_step(i):
            b = devb[i & 1]
            if st == 2:
                o, _ = model(b["image"], None, b["rects"])
            else:
                o = model(b["image"], b["points"])
            ld = crit(o, b["targets"])
            loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
            loss.backward()
            if world > 1:
                model.allreduce_grads(group)
        eager_step(0)
        for i in range(4):        # lets the model capture its forward / backward launch sequences (automatic after 2 runs)
            eager_step(i)
        eager = {"ms_per_step": timed(eager_step, min(steps, 10)),
                 "what": "the reference's own loop body: model(...) -> criterion(...) -> losses.backward() called eagerly; the "
                         "model replays its forward / backward launch sequences as CUDA graphs after two runs of a signature"}
        model._auto_graph = False
        model._graphs.clear()
        eager_step(0)
        eager["ms_per_step_no_graphs"] = timed(eager_step, min(steps, 10))
        model._auto_graph = True
        eager["value"] = B * world / eager["ms_per_step"] * 1e3

    # ---- roofline: the tcgen05 GEMM family (all forward / dgrad / wgrad GEMMs of one step), instrumented step.  Every
    # rank runs it (the criterion's num_boxes all-reduce is a collective); rank 0 records the events.
    eng = model.engine()
    saved_streams = (eng.side_stream, eng.aux_streams)
    eng.side_stream, eng.aux_streams = None, []      # serialise: each GEMM timed alone on one stream
    b0 = devb[0]

    def plain_step():
        if st == 2:
            o, _ = model(b0["image"], None, b0["rects"])
        else:
            o = model(b0["image"], b0["points"])
        ld = crit(o, b0["targets"])
        sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict).backward()

    model._auto_graph = False            # every launch of the instrumented step is timed on its own: no graph replay
    plain_step()
    if rank == 0:
        L.GEMM_TRACE, L.CALL_TRACE = [], []
    plain_step()
    torch.cuda.synchronize()
    trace, L.GEMM_TRACE = L.GEMM_TRACE, None
    atrace, L.CALL_TRACE = L.CALL_TRACE, None
    eng.side_stream, eng.aux_streams = saved_streams
    barrier()

    line = None
    if rank == 0:
        flops = sum(2.0 * m * n * k for (m, n, k, _, _) in trace)
        gemm_ms = sum(a.elapsed_time(b) for (_, _, _, a, b) in trace)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        traffic, traffic_note = None, None
        if args.workload == "c3" and world == 1:
            try:
                tpath = next(p for p in (os.path.join(ROOT, "profiles", f) for f in ("r02_gemm_traffic_c3.json", "r01_gemm_traffic_c3.json"))
                             if os.path.exists(p))
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes"] / max(tj["launches"], 1)
                traffic_note = (f"dram__bytes_read+write summed over the {tj['launches']} GEMM launches of one step "
                                f"= {tj['dram_bytes'] / 1e9:.2f} GB per step (ncu, cold cache, {os.path.basename(tpath)}); value = mean bytes per launch")
            except Exception:
                pass
        roofline = {"bound": "tensor", "kernel": "gemm_split_kernel (tcgen05 split-bf16 GEMM family: conv/linear fwd, dgrad, wgrad)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_note": traffic_note,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                    "algorithmic_gflop_per_step": flops / 1e9, "gemm_ms_per_step": gemm_ms, "gemm_launches_per_step": len(trace),
                    "note": "algorithmic (fp32-equivalent) FLOPs; each is issued as 3 bf16 MMAs (hi*hi+hi*lo+lo*hi), so the "
                            "tensor pipe executes 3x this figure", "issued_bf16_frac": 3 * achieved / peak,
                    "gemm_serial_ms_over_step_ms": gemm_ms / ms_step,
                    "timing": "CUDA events around every GEMM launch of one extra step with the side/aux streams disabled "
                              "(in the timed step weight-gradient GEMMs overlap the dgrad chain on a second stream)"}
        roofline_attention = attention_roofline(atrace, peak, ms_step)
        matcher = matcher_bench(dev) if (st == 2 and not args.skip_matcher) else None
        if args.skip_cpu:
            cb = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": "skipped"}
        else:
            cb, _ = cpu_arm(args.workload, 2, 1)
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16 (3-pass split-bf16 operands, fp32 accumulate/norms/softmax/loss)",
                "data": "synthetic",
                "config": {"workload": workload_desc(args.workload), "global_batch": B * world,
                           "step": "zero_grad -> forward -> criterion (device Hungarian matching) -> backward"
                                   + (" -> NCCL all-reduce of the flat gradient buffer" if world > 1 else "")
                                   + " (SURVEY.md 8d; optimizer reported separately under full_iteration)",
                           "parallelism": f"dp{world}" + (" (NCCL grad all-reduce + the reference's num_boxes all-reduce)" if world > 1 else ""),
                           "launch": "counting_detr_b200.CapturedStep: CUDA graph replay of the whole step" if use_graph else "eager launches",
                           "l2": "per-step working set (~20 GB of activations, 50 MB image batch) >> 126 MB L2, no flush needed; "
                                 "two alternating input batches"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * steps,
                "launches_per_step": launches_per_step, "full_iteration": full, "eager_api": eager,
                "roofline": roofline, "roofline_attention": roofline_attention, "matcher": matcher,
                "cpu_baseline": cb, "loss": loss_val}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-matcher", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the full-iteration (optimizer) and eager-API legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
