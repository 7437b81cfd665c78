"""Host -> device input pipeline for the training loop.

The reference copies every batch synchronously on the compute stream (`v.cuda()` ... train.py:183-187): 154 MB of
fp32 image per 256-sample batch, about 2.5 ms of PCIe time in front of every step.  `DevicePrefetcher` wraps any
iterable of reference-style batches `(v, q, a_indices, a_values, a_length, index, q_len)` (host tensors, ideally
pinned as main.py:122-132 requests with pin_memory=True), copies batch i+1 on a side stream into one of two
device buffers while batch i computes, and yields device batches in the same tuple order, so `run_batch` and the
unchanged loop of train.py:69-81 consume them directly (their `.cuda()` / `.to(device)` calls become no-ops).
Every batch is still copied exactly once; only the wait moves off the critical path.

The device buffers are a FIXED set of `depth` slots that live as long as the prefetcher (re-iterating it, epoch after
epoch, reuses them), so `GraphedTrainStep` sees at most `depth` distinct input buffer sets and replays one captured CUDA
graph per slot.

Image dtype: the reference stores pre-processed images as float16 (preprocessing/preprocess_images.py:40) and its Dataset
widens every sample to float32 on the host (preprocessing/data_preprocessing.py:174, `.astype('float32')`).  `VqaNet.forward`
reads float16 images natively (bit-identical results: the widening is exact), so a loader that drops that `.astype` hands
over half the bytes.  `image_dtype=torch.float16` makes the prefetcher do the narrowing itself for loaders that still
yield float32 -- only valid when the values are float16-representable, which they are for the reference's h5 files by
construction; it is checked on the first batch.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable[Sequence], device: Optional[torch.device] = None, depth: int = 2,
                 image_dtype: Optional[torch.dtype] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("DevicePrefetcher needs a CUDA device (dl_vqa_b200 has no CPU path)")
        self.batches = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = max(2, int(depth))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0
        self.image_dtype = image_dtype
        self._checked_image = False
        self._slots = [dict() for _ in range(self.depth)]      # persistent device buffers: stable addresses across epochs
        self._pinned = [None] * self.depth                     # staging for narrowed images (image_dtype)

    def _narrow_image(self, v, k):
        """float32 host image batch -> pinned `image_dtype` staging buffer of slot k (exactness checked once)."""
        if self.image_dtype is None or v.dtype == self.image_dtype or not v.is_floating_point():
            return v
        buf = self._pinned[k]
        if buf is None or buf.shape != v.shape:
            buf = self._pinned[k] = torch.empty(v.shape, dtype=self.image_dtype).pin_memory()
        buf.copy_(v)
        if not self._checked_image:
            if not torch.equal(buf.to(v.dtype), v):
                raise ValueError(f"DevicePrefetcher(image_dtype={self.image_dtype}): the images are not exactly representable")
            self._checked_image = True
        return buf

    def _stage(self, host_batch, slot, k=0):
        """Issue the copies of one batch on the copy stream; returns (device tensors, ready event)."""
        compute = torch.cuda.current_stream(self.device)
        if self.image_dtype is not None and torch.is_tensor(host_batch[0]):
            if self._pinned[k] is not None:
                slot["ev_host"].synchronize()                   # the copy that last read this staging buffer has finished
            host_batch = (self._narrow_image(host_batch[0], k),) + tuple(host_batch[1:])
        # device buffers belong to the compute stream's allocator pool (they are consumed there)
        for i, t in enumerate(host_batch):
            if torch.is_tensor(t):
                buf = slot.get(i)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    slot[i] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        # the buffers of this slot may still be read by the step that consumed them `depth` batches ago
        self.copy_stream.wait_stream(compute)
        with torch.cuda.stream(self.copy_stream):
            out = []
            for i, t in enumerate(host_batch):
                if not torch.is_tensor(t):
                    out.append(t)
                    continue
                slot[i].copy_(t, non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
                out.append(slot[i])
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            slot["ev_host"] = ev
        return tuple(out), ev

    def __iter__(self) -> Iterator[tuple]:
        slots = self._slots
        it = iter(self.batches)
        pending = []
        n = 0
        for host_batch in it:
            pending.append(self._stage(host_batch, slots[n % self.depth], n % self.depth))
            n += 1
            if len(pending) == self.depth:
                dev_batch, ev = pending.pop(0)
                torch.cuda.current_stream(self.device).wait_event(ev)
                yield dev_batch
        while pending:
            dev_batch, ev = pending.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield dev_batch
