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
    """n_slots handles in flight on n_slots long-lived streams.  `chunk_images`: batches with more
    images are cut into sub-batches of at most that many images, each a handle of its own, so the
    upload of one sub-batch overlaps the kernels of the previous one and the first kernels start (and
    the last download ends) after a fraction of the batch's PCIe time -- lower latency for a single
    large call; in a long sweep the steady-state rate is the same."""

    def __init__(self, n_slots=3, device=None, chunk_images=None):
        self.n_slots = int(n_slots)
        self.device = device
        self.chunk_images = None if not chunk_images else int(chunk_images)
        self._slots = [None] * self.n_slots   # (handle, result views, internal id)
        self._next = 0          # internal (per-handle) ids
        self._next_ticket = 0   # public tickets
        self._parts = {}        # ticket -> internal ids
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
        crf, res, iid = self._slots[slot]
        crf.synchronize()
        crf.close()
        self._slots[slot] = None
        self._done[iid] = res

    def _submit_one(self, sizes, n_labels, unary, rgb, cfg, out, labels):
        iid = self._next
        slot = iid % self.n_slots
        self._next += 1
        self._finish(slot)
        crf = DenseCRFBatch(sizes, n_labels, device=self.device, stream=self._streams[slot])
        crf.set_async_host(True)
        # unaries first: in async-host mode their upload runs on a separate stream and the lattice
        # builds enqueued next overlap it (the small image upload goes ahead on the handle's stream)
        crf.setUnaryEnergy(unary)
        crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
        crf.addPairwiseBilateral(sxy=cfg["bi_sxy"], srgb=cfg["bi_srgb"], rgbim=rgb, compat=cfg["bi_compat"])
        if hasattr(out, "data_ptr"):   # torch CUDA tensor: device-resident output (inputs may be too)
            res = crf.map_device(cfg["iterations"], out=out) if labels else crf.inference_device(cfg["iterations"], out=out)
        else:
            res = crf.map(cfg["iterations"], out=out) if labels else crf.inference(cfg["iterations"], out=out)
        self._slots[slot] = (crf, res, iid)
        return iid

    @staticmethod
    def _cut(x, lo, hi, starts, per_pixel):
        """Images [lo, hi) of a per-image list or of a flat concatenation (`per_pixel` elements per pixel)."""
        if x is None:
            return None
        if isinstance(x, (list, tuple)):
            return x[lo:hi]
        flat = x.reshape(-1) if hasattr(x, "reshape") else x
        return flat[starts[lo] * per_pixel:starts[hi] * per_pixel]

    def submit(self, sizes, n_labels, unary, rgb, cfg, out=None, labels=False):
        """Enqueue one batch; returns a ticket for result().  unary / rgb / out: concatenated host
        arrays (or per-image lists); they must stay alive and untouched until result(ticket)."""
        ticket = self._next_ticket
        self._next_ticket += 1
        B = len(sizes)
        chunk = self.chunk_images
        if not chunk or B <= chunk:
            self._parts[ticket] = [self._submit_one(sizes, n_labels, unary, rgb, cfg, out, labels)]
            return ticket
        starts = np.concatenate([[0], np.cumsum([int(w) * int(h) for (w, h) in sizes])]).astype(np.int64)
        parts = []
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            parts.append(self._submit_one(
                sizes[lo:hi], n_labels, self._cut(unary, lo, hi, starts, n_labels), self._cut(rgb, lo, hi, starts, 3),
                cfg, self._cut(out, lo, hi, starts, 1 if labels else n_labels), labels))
        self._parts[ticket] = parts
        return ticket

    def result(self, ticket):
        """Block until batch `ticket` is complete; -> list of per-image (L, N_b) marginals or label maps."""
        parts = []
        for iid in self._parts.pop(ticket):
            if iid not in self._done:
                self._finish(iid % self.n_slots)
            parts.append(self._done.pop(iid))
        if len(parts) == 1:
            return parts[0]
        if all(isinstance(r, list) for r in parts):   # host results: per-image lists, in image order
            return [x for r in parts for x in r]
        return parts                                   # device results: one tensor per sub-batch

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
