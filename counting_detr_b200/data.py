"""Host -> device input staging for the training loop (first slice).

The reference moves every batch with blocking `.to(device)` calls at the top of the iteration
(A2/engine.py:25-31, A1/engine.py:49-54), so the H2D copy of batch i+1 never overlaps the compute of batch i.
`DevicePrefetcher` wraps any iterable of batches (dicts / lists / tuples of tensors, arbitrarily nested; non-tensor
leaves pass through): it pins each host tensor once into reusable pinned staging buffers, issues the copy of the NEXT
batch on a dedicated copy stream while the caller works on the current one, and hands out device tensors that are
safe to use on the current stream (event-ordered; double-buffered so that a batch is never overwritten while a step
still reads it).  The host only ever waits for a copy issued `depth` batches earlier (before re-using its pinned staging buffer).
"""
import ctypes as C

import torch

from . import _lib as L

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # A2/data/fsc147.py:23


def normalize_u8(images_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """transforms.ToTensor() + transforms.Normalize(mean, std) (A2/data/fsc147.py:22-24,82) on the device: uint8
    [B,H,W,3] (or [H,W,3]) CUDA tensor -> fp32 [B,3,H,W], bit-identical to torchvision's result.  Ship the decoded,
    resized uint8 pixels to the GPU (DevicePrefetcher) and normalise there: a quarter of the H2D bytes."""
    x = images_u8 if images_u8.dim() == 4 else images_u8[None]
    assert x.dtype == torch.uint8 and x.shape[-1] == 3 and x.is_cuda, "uint8 [B,H,W,3] CUDA tensor expected"
    x = x.contiguous()
    B, H, W, _ = x.shape
    if out is None:
        out = torch.empty(B, 3, H, W, device=x.device)
    m = (C.c_float * 3)(*[float(v) for v in mean]); s = (C.c_float * 3)(*[float(v) for v in std])
    L.call("cdetr_normalize_u8", x, B, H, W, m, s, out)
    return out


class DevicePrefetcher:
    """see the module docstring.  Small host tensors of a batch (boxes, labels, rects, sizes ...: everything below
    `pack_below` bytes) are packed into ONE pinned staging buffer and cross PCIe as one copy; the device tensors handed
    out are 256-byte-aligned views of one device buffer.  Large tensors (the image batch) are copied on their own,
    straight from the caller's pinned memory when it is pinned."""

    def __init__(self, iterable, device, depth=2, pack_below=1 << 16, defer=False):
        self.it = iterable
        self.dev = torch.device(device)
        self.depth = max(int(depth), 2)
        self.pack_below = int(pack_below)
        self.defer = bool(defer)      # True: the caller issues the next batch itself with kick() (see there)
        self._it, self._i, self._pending = None, 0, None
        self.stream = torch.cuda.Stream(device=self.dev)
        self._slots = [dict(pin={}, dev={}, ready=torch.cuda.Event(), free=None, pack_pin=None, pack_dev=None)
                       for _ in range(self.depth)]
        self.h2d_bytes = 0
        self.h2d_copies = 0

    # ---- pass 1: collect the small leaves; pass 2: rebuild the structure with device tensors
    def _collect(self, obj, small):
        if isinstance(obj, torch.Tensor):
            if obj.device.type != "cuda" and obj.numel() * obj.element_size() < self.pack_below:
                small.append(obj)
        elif isinstance(obj, dict):
            for v in obj.values():
                self._collect(v, small)
        elif isinstance(obj, (list, tuple)):
            for v in obj:
                self._collect(v, small)

    def _stage_small(self, small, slot):
        offs, total = [], 0
        for t in small:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        if total == 0:
            return {}
        if slot["pack_pin"] is None or slot["pack_pin"].numel() < total:
            slot["pack_pin"] = torch.empty(total, dtype=torch.uint8).pin_memory()
            slot["pack_dev"] = torch.empty(total, dtype=torch.uint8, device=self.dev)
        pin, dev = slot["pack_pin"], slot["pack_dev"]
        out = {}
        for t, o in zip(small, offs):
            nb = t.numel() * t.element_size()
            pin[o:o + nb].view(t.dtype).view(t.shape).copy_(t.contiguous())          # host memcpy, a few hundred bytes
            out[id(t)] = dev[o:o + nb].view(t.dtype).view(t.shape)
        dev[:total].copy_(pin[:total], non_blocking=True)
        self.h2d_bytes += sum(t.numel() * t.element_size() for t in small)
        self.h2d_copies += 1
        return out

    def _stage(self, obj, slot, path, packed):
        if isinstance(obj, torch.Tensor):
            if obj.device.type == "cuda":
                return obj
            if id(obj) in packed:
                return packed[id(obj)]
            key = (path, tuple(obj.shape), obj.dtype)
            d = slot["dev"].get(key)
            if d is None:
                d = slot["dev"][key] = torch.empty(obj.shape, dtype=obj.dtype, device=self.dev)
            src = obj
            if not obj.is_pinned():
                p = slot["pin"].get(key)
                if p is None:
                    p = slot["pin"][key] = torch.empty(obj.shape, dtype=obj.dtype).pin_memory()
                p.copy_(obj)
                src = p
            d.copy_(src, non_blocking=True)
            self.h2d_bytes += obj.numel() * obj.element_size()
            self.h2d_copies += 1
            return d
        if isinstance(obj, dict):
            return {k: self._stage(v, slot, path + (k,), packed) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._stage(v, slot, path + (i,), packed) for i, v in enumerate(obj))
        return obj

    def _issue(self, batch, i):
        slot = self._slots[i % self.depth]
        if i >= self.depth:
            slot["ready"].synchronize()              # this slot's previous H2D has left its pinned staging buffers
        if slot["free"] is not None:
            self.stream.wait_event(slot["free"])     # the step that used this slot's tensors has been enqueued & done
        small = []
        self._collect(batch, small)
        with torch.cuda.stream(self.stream):
            packed = self._stage_small(small, slot)
            out = self._stage(batch, slot, (), packed)
            slot["ready"].record(self.stream)
        return out, slot

    def kick(self):
        """Issue the next batch's copies NOW.  A loop that reads a result synchronously every step (the reference reads
        the loss) should call this right after launching its step: the copies then overlap the step instead of sitting,
        together with their host-side set-up, between the read and the next launch.  Without kick() the next batch is
        issued just before the current one is handed out (always overlapped, but on the host's critical path)."""
        if self._pending is None and self._it is not None:
            try:
                self._pending = self._issue(next(self._it), self._i)
                self._i += 1
            except StopIteration:
                self._it = None

    def __iter__(self):
        self._it = iter(self.it)
        self._i = 0
        self._pending = None
        self.kick()
        first = True
        while self._pending is not None:
            cur, slot = self._pending
            self._pending = None
            if not self.defer or first:
                self.kick()                          # next batch travels while the caller computes on this one
            first = False
            torch.cuda.current_stream(self.dev).wait_event(slot["ready"])
            yield cur
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))   # everything the caller enqueued on this batch
            slot["free"] = ev
            if self._pending is None:
                self.kick()                          # deferred mode and the caller did not kick: issue now
