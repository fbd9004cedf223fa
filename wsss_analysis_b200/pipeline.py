"""Host-buffer batch pipeline: keeps two (or more) batches in flight from ONE host thread so that the
PCIe copies of one batch overlap the kernels of another.

The reference drives the CRF from plain loops over host NumPy arrays
(/root/reference/03c_hsn/utilities.py:424, 03a_sec-dsrg/model.py:665).  A batch whose inputs and
outputs live in host memory costs two PCIe transfers (unaries in, marginals out) around the GPU work;
run back to back they leave the GPU idle a third of the time.  Here every batch gets its own handle on
a long-lived per-slot stream in async-host mode (`DCRF_OPT_ASYNC_HOST`): its H2D copy, lattice build, iterations
and D2H copy are only enqueued, and the host moves on to the next batch while the previous one is
still computing / downloading.  (Driving the handles from several host threads instead makes the
threads collide inside the CUDA driver's allocator locks -- measured)
All calls go through the C ABI with host pointers (`on_device = 0`); host buffers should be
page-locked (`pinned_empty`).
"""
import ctypes as C

import numpy as np

from . import _lib
from .densecrf import DenseCRFBatch


class BatchPipeline(object):
    def __init__(self, n_slots=3, device=None):
        self.n_slots = int(n_slots)
        self.device = device
        self._slots = [None] * self.n_slots   # (handle, result views, ticket)
        self._next = 0
        self._done = {}
        # one long-lived stream (and therefore one device-memory pool) per slot
        self._lib = _lib.load()
        self._streams = []
        for _ in range(self.n_slots):
            s = C.c_void_p()
            _lib.check(self._lib.dcrf_stream_create(-1 if device is None else int(device), C.byref(s)))
            self._streams.append(s.value)

    def _finish(self, slot):
        if self._slots[slot] is None:
            return
        crf, res, ticket = self._slots[slot]
        crf.synchronize()
        crf.close()
        self._slots[slot] = None
        self._done[ticket] = res

    def submit(self, sizes, n_labels, unary, rgb, cfg, out=None, labels=False):
        """Enqueue one batch; returns a ticket for result().  unary / rgb / out: concatenated host
        arrays (or per-image lists); they must stay alive and untouched until result(ticket)."""
        ticket = self._next
        slot = ticket % self.n_slots
        self._next += 1
        self._finish(slot)
        crf = DenseCRFBatch(sizes, n_labels, device=self.device, stream=self._streams[slot])
        crf.set_async_host(True)
        crf.setUnaryEnergy(unary)
        crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
        crf.addPairwiseBilateral(sxy=cfg["bi_sxy"], srgb=cfg["bi_srgb"], rgbim=rgb, compat=cfg["bi_compat"])
        if hasattr(out, "data_ptr"):   # torch CUDA tensor: device-resident output (inputs may be too)
            res = crf.map_device(cfg["iterations"], out=out) if labels else crf.inference_device(cfg["iterations"], out=out)
        else:
            res = crf.map(cfg["iterations"], out=out) if labels else crf.inference(cfg["iterations"], out=out)
        self._slots[slot] = (crf, res, ticket)
        return ticket

    def result(self, ticket):
        """Block until batch `ticket` is complete; -> list of per-image (L, N_b) marginals or label maps."""
        if ticket not in self._done:
            self._finish(ticket % self.n_slots)
        return self._done.pop(ticket)

    def map(self, batches):
        """batches: iterable of dicts(sizes, n_labels, unary, rgb, cfg[, out, labels]); results in order."""
        tickets = [self.submit(**b) for b in batches]
        return [self.result(t) for t in tickets]

    def close(self):
        for s in range(self.n_slots):
            self._finish(s)
        for st in self._streams:
            self._lib.dcrf_stream_destroy(st)
        self._streams = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def pinned_empty(n, dtype=np.float32):
    """Host array in page-locked memory (via torch), for use as pipeline input / output."""
    import torch

    t = torch.empty(int(n), dtype={np.float32: torch.float32, np.uint8: torch.uint8, np.int32: torch.int32}[dtype])
    return t.pin_memory().numpy()  # the ndarray keeps the pinned tensor alive through its base
