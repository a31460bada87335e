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
def cpu_arm(name, steps, warmup, batch=1):
    """The oracle port (oracle/model.py + oracle/criterion.py: CPU PyTorch fp32 restatement of the reference,
    pinned against the live reference import and its goldens) timed on the host cores: fwd + loss + bwd."""
    import torch
    from oracle import criterion as OC, model as OM, weights as OW
    st, _, S, Q, T = WORKLOADS[name]
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = OM.Config(stage=st, num_query_position=Q)
    sd = OW.make_state_dict(cfg, 0)
    frozen = ("backbone.body.conv1", "backbone.body.layer1", "running_", ".bn", "downsample.1")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and not any(f in k for f in frozen)) for k, v in sd.items()}
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
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of B={batch} (same shapes per image as the GPU workload), {warmup} warm-up, "
                      f"fwd+criterion+bwd, torch CPU fp32, {dt * 1e3:.0f} ms/step"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 3)
    warm = min(max(args.warmup, 1), 1)
    cb, dt = cpu_arm(args.workload, steps, warm, batch=1)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload), "note": "reference CPU path = oracle port; "
                       "/root/reference is not present on the GPU box; B=1 sample per step"},
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


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from counting_detr_b200 import _lib as L
    from counting_detr_b200.models import build_model
    from counting_detr_b200 import synthetic as SY
    from counting_detr_b200.parallel import shard_seed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    st, B, S, Q, T = WORKLOADS[args.workload]
    margs = SY.default_args(st, num_query_position=Q, device=str(dev))
    model, crit, _ = build_model(margs)
    model.load_state_dict(SY.make_state_dict(SY.SynthCfg(stage=st, num_query_position=Q), 0), strict=True)
    model.to(dev).train(); crit.train()
    if world > 1:
        # keep NCCL out of the captured graph: gradients alias the flat buffer, one all-reduce after the replay;
        # T is constant in this benchmark so the criterion's 1-float num_boxes all-reduce is frozen
        model.alias_param_grads(True)
        if st == 2:
            crit.fixed_num_boxes = float(B * T)
    inp = SY.make_inputs(B, S, T=T, seed=shard_seed(0, rank), stage=st, Q=Q)
    img_h = inp["image"].pin_memory()
    img_d = img_h.to(dev)
    rects_h = inp.get("rects")
    if st == 2:
        tb_h = [t["boxes"].pin_memory() for t in inp["targets"]]
        targets = [{"boxes": b.to(dev), "labels": t["labels"].to(dev)} for b, t in zip(tb_h, inp["targets"])]
    else:
        pts_h, whs_h = inp["points"].pin_memory(), inp["whs"].pin_memory()
        pts_d = pts_h.to(dev)
        targets = {"points": pts_d, "whs": whs_h.to(dev)}

    def step():
        model.zero_grad(set_to_none=True)
        if st == 2:
            out, _ = model(img_d, None, rects_h)
        else:
            out = model(img_d, pts_d)
        ld = crit(out, targets)
        loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
        loss.backward()
        return loss.detach()     # keep no reference to the autograd graph (needed for graph capture)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager): allocates every buffer, packs weights
    for _ in range(max(args.warmup, 3)):
        loss = step()
    torch.cuda.synchronize()
    L.COUNTER["launches"] = 0
    step()
    launches_per_step = L.COUNTER["launches"]
    loss_val = float(loss)
    del loss
    # ---- optional whole-step CUDA graph
    graph, use_graph, g_loss = None, not args.no_graph, None
    if use_graph:
        try:
            s = torch.cuda.Stream(priority=-1)     # critical path at high priority; wgrad side stream is low
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                g_loss = step()
            graph.replay()
            torch.cuda.synchronize()
            if abs(float(g_loss) - loss_val) > 1e-3 * abs(loss_val):
                raise RuntimeError(f"graph replay loss {float(g_loss)} != eager loss {loss_val}")
        except Exception as e:  # fall back to eager launches of the same kernels
            if rank == 0:
                import traceback
                print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {str(e)[:300]}); timing eager launches", file=sys.stderr)
                if os.environ.get("BENCH_DEBUG"):
                    traceback.print_exc()
            graph, use_graph = None, False
            torch.cuda.synchronize()

    if world > 1:   # all ranks must take the same path (a lone eager rank would issue different collectives)
        flag = torch.tensor([1 if graph is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            graph, use_graph = None, False

    def run_one():
        if graph is not None:
            graph.replay()
        else:
            step()
        if world > 1:
            model.allreduce_grads(dist.group.WORLD)

    for _ in range(3):
        run_one()
    # ---- timed region: exactly K steps, device events, max over ranks
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_one()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t)
    ms_step = ms_total / args.steps
    value = B * world / ms_step * 1e3

    # ---- e2e: public API from pinned host buffers, H2D + loss D2H inside the timed region.  With a captured
    # step the host copies land in the static input tensors the graph reads, then the graph is replayed.
    # Input pipeline of the captured path (what a training loop with a prefetching loader does): the pinned host batch of
    # step i+1 is copied to a device staging buffer on a copy stream while step i computes; step i+1 starts with a
    # device-to-device move into the tensors the graph reads.  Every step's H2D copy and loss read are inside the
    # timed region.
    copy_stream = torch.cuda.Stream()
    img_stage = torch.empty_like(img_d)
    h2d_done = torch.cuda.Event()

    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"n": 0}

    def flush_loss():   # read the last step's loss (end of a run of e2e steps)
        if graph is not None and e2e_state["n"] >= 1:
            slot = (e2e_state["n"] - 1) & 1
            loss_ev[slot].synchronize()
            e2e_state["n"] = 0
            return float(loss_pin[slot])
        return None

    stage_free = torch.cuda.Event()

    def prefetch_inputs():
        copy_stream.wait_event(stage_free)   # the d2d move of the running step has consumed the staging buffer
        with torch.cuda.stream(copy_stream):
            img_stage.copy_(img_h, non_blocking=True)
            h2d_done.record(copy_stream)

    def e2e_step():
        if graph is not None:
            torch.cuda.current_stream().wait_event(h2d_done)
            img_d.copy_(img_stage, non_blocking=True)
            stage_free.record()
            if st == 2:
                for t, b in zip(targets, tb_h):
                    t["boxes"].copy_(b, non_blocking=True)
            else:
                pts_d.copy_(pts_h, non_blocking=True)
                targets["whs"].copy_(whs_h, non_blocking=True)
            graph.replay()
            if world > 1:
                model.allreduce_grads(dist.group.WORLD)
            prefetch_inputs()          # next step's image batch travels while this step computes
            # the loss leaves through an asynchronous copy into pinned memory; the host reads the PREVIOUS step's
            # value while this step runs (one device->host read per step, none of them skipped: see flush_loss)
            slot = e2e_state["n"] & 1
            loss_pin[slot].copy_(g_loss, non_blocking=True)
            loss_ev[slot].record()
            e2e_state["n"] += 1
            if e2e_state["n"] >= 2:
                loss_ev[slot ^ 1].synchronize()
                return float(loss_pin[slot ^ 1])
            return None
        img = img_h.to(dev, non_blocking=True)
        model.zero_grad(set_to_none=True)
        if st == 2:
            tg = [{"boxes": b.to(dev, non_blocking=True), "labels": t["labels"]} for b, t in zip(tb_h, targets)]
            out, _ = model(img, None, rects_h)
            ld = crit(out, tg)
        else:
            p = pts_h.to(dev, non_blocking=True)
            out = model(img, p)
            ld = crit(out, {"points": p, "whs": whs_h.to(dev, non_blocking=True)})
        loss = sum(ld[k] * crit.weight_dict[k] for k in ld if k in crit.weight_dict)
        loss.backward()
        if world > 1:
            model.allreduce_grads(dist.group.WORLD)
        return loss.item()

    if graph is not None:
        stage_free.record()
        prefetch_inputs()
    for _ in range(2):
        e2e_step()
    flush_loss()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    flush_loss()                                           # the K-th loss is read inside the timed region too
    torch.cuda.current_stream().wait_stream(copy_stream)   # the K-th prefetch copy also ends inside the timed region
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    h2d = img_h.numel() * 4 + (sum(b.numel() for b in tb_h) * 4 if st == 2 else (pts_h.numel() + whs_h.numel()) * 4)
    e2e = {"value": B * world / (ms_e2e / args.steps) * 1e3, "unit": "images/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 4}

    line = None
    if rank == 0:
        # ---- roofline: the tcgen05 GEMM family (all forward / dgrad / wgrad GEMMs of one step), instrumented step
        eng = model.engine()
        saved_streams = (eng.side_stream, eng.aux_streams)
        eng.side_stream, eng.aux_streams = None, []      # serialise: each GEMM timed alone on one stream
        L.GEMM_TRACE, L.CALL_TRACE = [], []
        step()
        torch.cuda.synchronize()
        trace, L.GEMM_TRACE = L.GEMM_TRACE, None
        atrace, L.CALL_TRACE = L.CALL_TRACE, None
        eng.side_stream, eng.aux_streams = saved_streams
        flops = sum(2.0 * m * n * k for (m, n, k, _, _) in trace)
        gemm_ms = sum(a.elapsed_time(b) for (_, _, _, a, b) in trace)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        # DRAM traffic of the family from the committed ncu capture of the same workload (None for other workloads)
        traffic, traffic_note = None, None
        if args.workload == "c3" and world == 1:
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic_c3.json")))
                traffic = tj["dram_bytes"] / max(tj["launches"], 1)
                traffic_note = (f"dram__bytes_read+write summed over the {tj['launches']} GEMM launches of one step "
                                f"= {tj['dram_bytes'] / 1e9:.2f} GB per step (ncu, cold cache); value = mean bytes per launch")
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
        cb, _ = cpu_arm(args.workload, 2, 1, batch=1) if not args.skip_cpu else ({"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": "skipped"}, 0)
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16 (3-pass split-bf16 operands, fp32 accumulate/norms/softmax/loss)",
                "data": "synthetic",
                "config": {"workload": workload_desc(args.workload), "global_batch": B * world,
                           "parallelism": f"dp{world}" + (" (NCCL grad all-reduce)" if world > 1 else ""),
                           "launch": "CUDA graph replay of the whole step" if use_graph else "eager launches",
                           "l2": "per-step working set (~20 GB of activations, 50 MB image batch) >> 126 MB L2, no flush needed"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "launches_per_step": launches_per_step, "roofline": roofline, "roofline_attention": roofline_attention,
                "cpu_baseline": cb,
                "loss": loss_val}
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
