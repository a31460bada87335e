"""One training iteration of the reference's loop as a replayable CUDA graph.

    step = CapturedStep(model, criterion, optimizer=FusedAdamW(...), max_norm=0.1)
    for batch in loader:
        losses, total = step(batch["image"], targets, rects=batch["ex_rects"])     # device scalars, no host sync
        ...
        total.item()           # only when the caller wants the value (the reference reads it every step)

The body is exactly the call order of the reference's train_one_epoch (A2/engine.py:24-63, A1/engine.py:48-72):
model forward -> criterion (device Hungarian matching + losses) -> weighted sum over criterion.weight_dict ->
zero_grad -> backward -> clip_grad_norm_(max_norm) -> optimizer.step(); it runs through the public
model()/criterion() objects of counting_detr_b200.models, once eagerly (warm-up: allocates every buffer at a fixed
address) and once under torch.cuda.graph; afterwards every call copies the batch into the static input tensors
(pinned-host or device sources, non-blocking) and replays ~1000 kernel launches as one graph launch.

What makes the step capturable: no kernel of the path reads device data on the host -- exemplar
centres, padding masks, num_boxes and the matcher all stay on the device -- and every intermediate lives in the
engine's fixed-address buffers.  Data parallel (world > 1): NCCL stays outside the graph; the 1-float num_boxes
all-reduce of the reference (A2/models/anchor_detr.py:321-325) is issued before the replay, the flat-gradient
all-reduce after it, then the fused clip + AdamW.
"""
import torch

from . import _lib as L


class CapturedStep:
    def __init__(self, model, criterion, optimizer=None, max_norm=0.0, group=None, use_graph=True, warmup=2):
        """optimizer: a counting_detr_b200.optim.FusedAdamW (device-resident step counter / hyper-parameters) or None
        (forward + loss + backward only).  group: torch.distributed process group for data parallelism."""
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.max_norm = float(max_norm or 0.0)
        self.group = group
        self.use_graph = use_graph
        self.warmup = max(int(warmup), 1)
        self._key = None
        self._graph = self._graph_tail = None
        self._static = None
        self._out = None
        self._stream = None
        self._version = None
        self._evictions = 0
        self.launches_per_step = None
        model.alias_param_grads(True)      # p.grad = views of the flat gradient buffer: fixed addresses, no copies
        if optimizer is not None and not hasattr(optimizer, "note_replayed_step"):
            raise TypeError("CapturedStep needs counting_detr_b200.optim.FusedAdamW (device-resident optimizer state)")

    # ------------------------------------------------------------------ static inputs
    @staticmethod
    def _sig(samples, targets, rects, points):
        def shp(t):
            return None if t is None else tuple(t.shape)
        if isinstance(targets, dict):
            tsig = tuple((k, shp(v)) for k, v in sorted(targets.items()))
        else:
            tsig = tuple(shp(t["boxes"]) for t in targets)
        if hasattr(samples, "decompose"):
            im, mk = samples.decompose()
            ssig = (shp(im), shp(mk))
        else:
            ssig = (shp(samples), None)
        return (ssig, tsig, shp(rects) if isinstance(rects, torch.Tensor) else None,
                shp(points) if isinstance(points, torch.Tensor) else None)

    def _make_static(self, samples, targets, rects, points, dev):
        st = {}
        if hasattr(samples, "decompose"):
            im, mk = samples.decompose()
        else:
            im, mk = samples, None
        st["image"] = torch.empty(im.shape, dtype=torch.float32, device=dev)
        st["mask"] = torch.empty(mk.shape, dtype=torch.bool, device=dev) if mk is not None else None
        st["rects"] = torch.empty(tuple(rects.shape), dtype=torch.float32, device=dev) if rects is not None else None
        if isinstance(targets, dict):
            st["targets"] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in targets.items()}
        else:
            st["targets"] = [{"boxes": torch.empty(t["boxes"].shape, dtype=torch.float32, device=dev),
                              "labels": torch.zeros(t["boxes"].shape[0], dtype=torch.int64, device=dev)} for t in targets]
        st["points"] = None
        if isinstance(points, torch.Tensor):
            st["points"] = torch.empty(points.shape, dtype=torch.float32, device=dev)
        elif points is not None:      # ndarray (the reference's `sampled_points`): fixed for the lifetime of a capture
            st["points"] = torch.as_tensor(points, dtype=torch.float32).to(dev)
        return st

    def _fill(self, samples, targets, rects, points):
        st = self._static
        if hasattr(samples, "decompose"):
            im, mk = samples.decompose()
            st["mask"].copy_(mk, non_blocking=True)
        else:
            im = samples
        st["image"].copy_(im, non_blocking=True)
        if st["rects"] is not None:
            st["rects"].copy_(torch.as_tensor(rects), non_blocking=True)
        if isinstance(targets, dict):
            for k, v in targets.items():
                st["targets"][k].copy_(v, non_blocking=True)
        else:
            src = [t["boxes"] for t in targets]
            if all(x.is_cuda and x.dtype == torch.float32 for x in src):
                torch._foreach_copy_([s["boxes"] for s in st["targets"]], src)      # one multi-tensor launch
            else:
                for s, t in zip(st["targets"], targets):
                    s["boxes"].copy_(t["boxes"], non_blocking=True)
        if isinstance(points, torch.Tensor):
            st["points"].copy_(points, non_blocking=True)

    # ------------------------------------------------------------------ the step body (public API calls only)
    def _body(self):
        st, model, crit = self._static, self.model, self.criterion
        samples = st["image"] if st["mask"] is None else _Nested(st["image"], st["mask"])
        if model.stage == 2:
            out, _ = model(samples, st["points"], st["rects"])
        else:
            out = model(samples, st["points"])
        ld = crit(out, st["targets"])
        wd = crit.weight_dict
        total = sum(ld[k] * wd[k] for k in ld if k in wd)
        total.backward()                        # gradients land in the engine's flat buffer (zeroed by the forward)
        return {k: v.detach() for k, v in ld.items()}, total.detach()

    def _tail(self):
        """clip_grad_norm_ + AdamW (one fused multi-tensor pass) and the re-pack of the weights it wrote into the
        split-bf16 operand layout the next forward reads (125 pack + 53 FrozenBN-fold launches)."""
        if self.optimizer is None:
            return
        self.optimizer.step(max_norm=self.max_norm if self.max_norm > 0 else None)
        self.model.engine().pack_weights()
        self.model._param_version = self.model._current_version()

    def _prep(self):
        """hoisted 1-float num_boxes all-reduce (A2/models/anchor_detr.py:321-325): issued OUTSIDE any capture, the
        criterion's forward then finds the device-resident value prepared for exactly these target tensors"""
        crit = self.criterion
        if hasattr(crit, "prepare_num_boxes"):
            crit.prepare_num_boxes(self._static["targets"], next(self.model.parameters()).device, self.group)

    def _eager_step(self):
        self._prep()
        out = self._body()
        if self.group is not None:
            self.model.allreduce_grads(self.group)
        self._tail()
        return out

    def _capture(self, dev):
        self._graph = self._graph_tail = self._out = None
        if not self.use_graph:
            return
        # warm-up iterations are real steps (they allocate every buffer and build the optimizer's pointer table); with
        # an optimizer they would train on the first batch several times, so parameters and optimizer state are rolled
        # back afterwards: the first replay IS the first training step
        opt = self.optimizer
        ps = [p for p in self.model.parameters() if p.requires_grad] if opt is not None else []
        snap_p = [p.detach().clone() for p in ps]
        snap_o = opt.snapshot() if opt is not None else None
        self._stream = torch.cuda.Stream(device=dev, priority=-1)   # critical path high priority, wgrad side stream low
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(self.warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(self._stream)
        torch.cuda.synchronize(dev)
        c0 = L.COUNTER["launches"]
        self._prep()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            self._out = self._body()
            if self.group is None:
                self._tail()              # single GPU: optimizer + re-pack ride in the same graph
        self._graph = g
        if self.group is not None and self.optimizer is not None:
            g2 = torch.cuda.CUDAGraph()   # data parallel: [fwd..bwd] -> NCCL all-reduce (eager) -> [optimizer + re-pack]
            with torch.cuda.graph(g2, stream=self._stream):
                self._tail()
            self._graph_tail = g2
        self.launches_per_step = L.COUNTER["launches"] - c0
        if opt is not None:
            with torch.no_grad():
                for p, q in zip(ps, snap_p):
                    p.copy_(q)
            opt.restore(snap_o)
            self.model.engine().pack_weights()
            self.model._param_version = self.model._current_version()

    # ------------------------------------------------------------------ call
    def __call__(self, samples, targets, rects=None, points=None):
        model, crit = self.model, self.criterion
        dev = next(model.parameters()).device
        key = self._sig(samples, targets, rects, points)
        eng = model.engine()
        if key != self._key or eng.evictions != self._evictions:     # new shapes, or the engine freed this capture's buffers
            self._static = self._make_static(samples, targets, rects, points, dev)
            self._fill(samples, targets, rects, points)
            self._capture(dev)
            self._key = key
            self._evictions = eng.evictions
            self._version = model._current_version()
        self._fill(samples, targets, rects, points)
        if self.optimizer is not None:
            self.optimizer.sync_hyper()          # LR scheduler changes reach the device hyper-parameter array
        if self._graph is None:
            return self._eager_step()
        self._prep()
        ver = model._current_version()
        if ver != self._version:                 # someone else changed the weights (load_state_dict, ...): re-pack
            model.engine().pack_weights()
            model._param_version = self._version = ver
        self._graph.replay()
        if self.group is not None:
            model.allreduce_grads(self.group)
            if self._graph_tail is not None:
                self._graph_tail.replay()
        if self.optimizer is not None:
            self.optimizer.note_replayed_step()  # host mirrors of what the replayed kernels did on the device
        return self._out


class _Nested:
    """Minimal NestedTensor (A2/util/misc.py:311-333): .decompose() -> (tensors, mask)."""

    def __init__(self, tensors, mask):
        self.tensors, self.mask = tensors, mask

    def decompose(self):
        return self.tensors, self.mask
