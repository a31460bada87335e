"""Optimizer tail of the reference's train loop on the sm_100a library.

Drop-in replacements for the two calls that follow `losses.backward()` in the reference
(A2/engine.py:53-57, A2/main.py:157-189):

    grad_total_norm = clip_grad_norm_(model.parameters(), max_norm)      # torch.nn.utils.clip_grad_norm_
    optimizer = FusedAdamW(param_dicts, lr=args.lr, weight_decay=args.weight_decay)   # torch.optim.AdamW
    optimizer.step()

Both run as multi-tensor kernels over a (tensor, chunk) block table (cdetr_mt_* in include/cdetr.h): one
sum-of-squares pass + one in-place scale for the clip, one pass for AdamW.  `optimizer.step(max_norm=...)` fuses
the two: the clip coefficient is applied to the gradients on the fly inside the AdamW pass (the gradients are not
rewritten), which is the HBM floor for fp32 AdamW: 16 bytes read + 12 bytes written per parameter.
Nothing synchronises with the host: the step counter and the hyper-parameters live on the device, so the whole
train step including the optimizer can be captured in a CUDA graph; an LR scheduler only changes
`param_groups[i]["lr"]`, which is mirrored into the device hyper-parameter array when it changes.
There is no CPU fallback: tensors must live on a CUDA device.
"""
import ctypes as C

import torch

from . import _lib as L

CHUNK = 4096 * 4      # elements per block


class _MtTensor(C.Structure):
    """cdetr_mt_tensor_t"""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64),
                ("group", C.c_int32), ("pad_", C.c_int32)]


class _Table:
    """Device copy of the tensor table + block list for a fixed set of (p, g, m, v) pointers."""

    def __init__(self, entries, dev):
        # entries: list of (p_ptr, g_ptr, m_ptr, v_ptr, n, group)
        arr = (_MtTensor * len(entries))()
        blocks = []
        for i, (p, g, m, v, n, grp) in enumerate(entries):
            arr[i] = _MtTensor(p, g, m, v, n, grp, 0)
            for c in range((n + CHUNK - 1) // CHUNK):
                blocks += [i, c]
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(dev)
        self.blocks = torch.tensor(blocks, dtype=torch.int32).to(dev)
        self.nblocks = len(blocks) // 2
        self.partial = torch.empty(max(self.nblocks, 1), device=dev)
        self.key = tuple(e[:4] for e in entries)


_clip_cache = {}


@torch.no_grad()
def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ (L2 only) as two multi-tensor kernels; returns the total norm (0-dim tensor)."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the L2 norm the reference uses is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.tensor(0.0)
    dev = ps[0].grad.device
    if dev.type != "cuda":
        raise L.CdetrError("clip_grad_norm_: gradients must live on a CUDA device (no CPU fallback)")
    key = tuple(p.grad.data_ptr() for p in ps)
    tb = _clip_cache.get(dev)
    if tb is None or tb.key_g != key:
        for p in ps:
            assert p.grad.dtype == torch.float32 and p.grad.is_contiguous()
        tb = _Table([(0, p.grad.data_ptr(), 0, 0, p.grad.numel(), 0) for p in ps], dev)
        tb.key_g = key
        tb.norm_out = torch.zeros(3, device=dev)
        _clip_cache[dev] = tb
    L.call("cdetr_mt_grad_norm", tb.table, tb.blocks, tb.nblocks, CHUNK, float(max_norm), tb.partial, tb.norm_out)
    L.call("cdetr_mt_clip_scale", tb.table, tb.blocks, tb.nblocks, CHUNK, tb.norm_out)
    return tb.norm_out[1]


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (amsgrad=False, maximize=False) with the update of every tensor in ONE kernel launch.

    Same constructor and `param_groups` / `state_dict()` layout as torch.optim.AdamW (state: step, exp_avg,
    exp_avg_sq per parameter), so LR schedulers and the reference's checkpoint code work unchanged."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        b = {tuple(g["betas"]) for g in self.param_groups} | set()
        e = {g["eps"] for g in self.param_groups}
        if len(b) > 1 or len(e) > 1:
            raise NotImplementedError("betas and eps must be the same in every param group")
        self._tb = None
        self._hyper_host = None
        self._hyper_last = None
        self._step_dev = None
        self._norm_out = None
        self._ps = None
        self._pending_steps = 0

    def _ensure(self):
        entries, ps = [], []
        for gi, grp in enumerate(self.param_groups):
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda":
                    raise L.CdetrError("FusedAdamW: parameters must live on a CUDA device (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32)      # host copy, torch.optim.AdamW layout
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                assert p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                entries.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                p.numel(), gi))
                ps.append(p)
        if not entries:
            return None, ps
        dev = ps[0].device
        key = tuple(e[:4] for e in entries)
        if self._tb is None or self._tb.key != key:
            self._tb = _Table(entries, dev)
            if self._step_dev is None:
                first = self.state[ps[0]]["step"]
                self._step_dev = torch.full((1,), int(first), dtype=torch.int32, device=dev)
                self._norm_out = torch.zeros(3, device=dev)
                self._hyper_host = torch.zeros(len(self.param_groups), 4).pin_memory()
                self._hyper_dev = torch.zeros(len(self.param_groups), 4, device=dev)
        self._ps = ps
        self.sync_hyper()
        return self._tb, ps

    def sync_hyper(self):
        """Mirror param_groups[i]["lr"/"weight_decay"] into the device hyper-parameter array when they changed (an LR
        scheduler step); a captured optimizer step reads them from there."""
        if self._hyper_host is None:
            return
        hyper = [(float(g["lr"]), float(g["weight_decay"])) for g in self.param_groups]
        if hyper != self._hyper_last:
            for i, (lr, wd) in enumerate(hyper):
                self._hyper_host[i, 0] = lr
                self._hyper_host[i, 1] = wd
            self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
            self._hyper_last = hyper

    def note_replayed_step(self, n=1):
        """A CUDA-graph replay ran the captured step() kernels n more times: the device step counter advanced by
        itself, the host mirrors (state[p]["step"], torch.optim layout) are brought up to date lazily."""
        self._pending_steps += n

    def _flush_pending(self):
        if self._pending_steps:
            for p in self._ps or []:
                self.state[p]["step"] += self._pending_steps
            self._pending_steps = 0

    def snapshot(self):
        """Copy of the optimizer state (device step counter, moments, host step mirrors); see restore()."""
        self._flush_pending()
        return {"step_dev": None if self._step_dev is None else self._step_dev.clone(),
                "state": {p: (st["step"].clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone())
                          for p, st in self.state.items() if len(st)}}

    def restore(self, snap):
        """Undo every step() since snapshot() (CapturedStep rolls its warm-up iterations back with this).  Tensors are
        restored in place: the pointer table of a captured step stays valid."""
        self._flush_pending()
        for p, st in self.state.items():
            if not len(st):
                continue
            old = snap["state"].get(p)
            if old is None:
                st["step"].zero_(); st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
            else:
                st["step"].copy_(old[0]); st["exp_avg"].copy_(old[1]); st["exp_avg_sq"].copy_(old[2])
        if self._step_dev is not None:
            if snap["step_dev"] is None:
                self._step_dev.zero_()
            else:
                self._step_dev.copy_(snap["step_dev"])

    def state_dict(self):
        self._flush_pending()
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None, max_norm=None):
        """AdamW update.  With max_norm > 0 the gradient-norm clip (torch.nn.utils.clip_grad_norm_ semantics over
        the tensors this optimizer owns) is fused in and the total norm is returned (0-dim device tensor)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._flush_pending()
        tb, ps = self._ensure()
        if tb is None:
            return loss
        g0 = self.param_groups[0]
        norm = None
        if max_norm is not None and max_norm > 0:
            L.call("cdetr_mt_grad_norm", tb.table, tb.blocks, tb.nblocks, CHUNK, float(max_norm), tb.partial, self._norm_out)
            norm = self._norm_out
        L.call("cdetr_mt_adamw", tb.table, tb.blocks, tb.nblocks, CHUNK, self._hyper_dev, g0["betas"][0], g0["betas"][1],
               g0["eps"], self._step_dev, norm)
        for p in ps:                       # host mirror of the step count (state_dict compatibility); no sync
            self.state[p]["step"] += 1
        # the kernel wrote the parameters through raw pointers: bump their autograd version counters so that
        # in-place-modification checks and the model's "weights changed -> re-pack" test see the update
        torch.autograd.graph.increment_version(ps)
        if max_norm is not None and max_norm > 0:
            return self._norm_out[1]
        return loss
