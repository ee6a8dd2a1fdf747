"""Input pipeline: what the notebook does on the host for every image -
`img = (img - img.min()) / (img.max() - img.min())` in float64 (README.md:87), a float64 DataLoader batch, then
`imgs.type(Tensor)` (README.md:785) - moved to the device and overlapped with the training step.

The host hands over RAW voxel values in their storage dtype (uint8 / uint16 / int16 / float32 / float64); the batch
is staged in pinned memory, copied on a dedicated copy stream, and one kernel (`vg_normalize_images`) does the
per-image min-max normalisation in float64 arithmetic and writes the fp32 batch the step consumes.  Two slots
alternate, so the copy + normalisation of batch k+1 run under the kernels of step k:

    pipe = InputPipeline(device, (B, 1, 96, 96), torch.uint8)
    pipe.submit(raw[0])
    for k in range(steps):
        x = pipe.get()               # current stream waits for slot k's event (no host sync)
        pipe.submit(raw[k + 1])      # H2D + normalise of the next batch, overlapped with this step
        trainer.step(x)
"""
from __future__ import annotations

import torch

from . import _lib

_RAW = {torch.uint8: 0, torch.uint16: 1, torch.int16: 2, torch.float32: 3, torch.float64: 4}


def normalize_images(raw: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """(N, 1, H, W) or (N, H, W) raw device tensor -> per-image min-max normalised fp32 (N, 1, H, W)."""
    assert raw.is_cuda and raw.is_contiguous() and raw.dtype in _RAW, "contiguous CUDA tensor of a supported raw dtype"
    _lib.ensure_device(raw.device)
    n = raw.shape[0]
    pixels = raw.numel() // max(n, 1)
    if out is None:
        shape = raw.shape if raw.dim() == 4 else (n, 1) + tuple(raw.shape[1:])
        out = torch.empty(shape, dtype=torch.float32, device=raw.device)
    _lib.call("vg_normalize_images", raw.data_ptr(), _RAW[raw.dtype], n, pixels, out.data_ptr(), None, _lib.stream_ptr())
    return out


class InputPipeline:
    """Double-buffered pinned-host -> device staging with the normalisation kernel on a copy stream."""

    def __init__(self, device: torch.device, batch_shape, raw_dtype=torch.float32, normalize: bool = True, slots: int = 2):
        assert raw_dtype in _RAW
        assert normalize or raw_dtype == torch.float32, "without normalisation the batch must already be fp32 in [0, 1]"
        self.device, self.shape, self.raw_dtype, self.normalize = device, tuple(batch_shape), raw_dtype, normalize
        _lib.ensure_device(device)
        self.stream = torch.cuda.Stream(device=device)
        self.host = [torch.empty(self.shape, dtype=raw_dtype).pin_memory() for _ in range(slots)]
        self.raw = [torch.empty(self.shape, dtype=raw_dtype, device=device) for _ in range(slots)]
        self.out = [torch.empty(self.shape, dtype=torch.float32, device=device) for _ in range(slots)] if normalize else self.raw
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [None] * slots
        self.head = self.tail = 0
        self.h2d_bytes = self.host[0].numel() * self.host[0].element_size()

    def submit(self, batch: torch.Tensor):
        """Stage `batch` (host tensor of raw values; pinned tensors are copied from directly) into the next slot."""
        i = self.head % len(self.host)
        self.head += 1
        if self.consumed[i] is not None:
            self.stream.wait_event(self.consumed[i])        # the step that read this slot has finished with it
        src = batch
        if not batch.is_pinned():
            self.host[i].copy_(batch)                         # pageable -> pinned staging (host memcpy)
            src = self.host[i]
        with torch.cuda.stream(self.stream):
            self.raw[i].copy_(src, non_blocking=True)
            if self.normalize:
                n = self.shape[0]
                _lib.call("vg_normalize_images", self.raw[i].data_ptr(), _RAW[self.raw_dtype], n, self.raw[i].numel() // n,
                          self.out[i].data_ptr(), None, self.stream.cuda_stream)
            self.ready[i].record(self.stream)

    def get(self) -> torch.Tensor:
        """The oldest submitted batch as an fp32 device tensor; the CURRENT stream waits for its copy + normalisation."""
        assert self.tail < self.head, "get() without a matching submit()"
        i = self.tail % len(self.host)
        self.tail += 1
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[i])
        return self.out[i]

    def release(self, slot_tensor: torch.Tensor = None):
        """Mark the most recently returned batch as consumed by the work enqueued so far on the current stream."""
        i = (self.tail - 1) % len(self.host)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[i] = ev
