"""Diagnostic: per-phase device timeline of the host-buffer pipeline (3 slots, one host thread).

Every batch runs on its slot's stream: H2D + layout change (setUnaryEnergy), lattice builds
(addPairwise*), iterations (run), layout change + D2H (marginals).  CUDA events on the slot stream
mark the phase boundaries; all times are printed relative to the first event, so the overlap of the
copies of one batch with the kernels of another is visible without a profiler."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

import bench
from wsss_analysis_b200 import _lib
from wsss_analysis_b200.densecrf import DenseCRFBatch
from wsss_analysis_b200.pipeline import pinned_empty

B = 32
N_SLOTS = int(os.environ.get("SLOTS", "3"))
STEPS = 9
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
imgs, unaries = bench.make_inputs(B)
sizes = [(bench.W_IMG, bench.H_IMG)] * B
nU = B * bench.L_LAB * bench.W_IMG * bench.H_IMG
hU = [pinned_empty(nU, np.float32) for _ in range(N_SLOTS)]
hI = [pinned_empty(B * bench.W_IMG * bench.H_IMG * 3, np.uint8) for _ in range(N_SLOTS)]
hQ = [pinned_empty(nU, np.float32) for _ in range(N_SLOTS)]
for s in range(N_SLOTS):
    hU[s][:] = np.concatenate([u.ravel() for u in unaries])
    hI[s][:] = np.concatenate([im.ravel() for im in imgs])
lib = _lib.load()
streams = []
for _ in range(N_SLOTS):
    p = C.c_void_p()
    _lib.check(lib.dcrf_stream_create(0, C.byref(p)))
    streams.append(p.value)
ext = [torch.cuda.ExternalStream(s, device=dev) for s in streams]


def ev(slot):
    e = torch.cuda.Event(enable_timing=True)
    e.record(ext[slot])
    return e


def run(steps, record):
    slots = [None] * N_SLOTS
    marks = []
    for t in range(steps):
        s = t % N_SLOTS
        if slots[s] is not None:
            slots[s].synchronize()
            slots[s].close()
        crf = DenseCRFBatch(sizes, bench.L_LAB, device=0, stream=streams[s])
        crf.set_async_host(True)
        m = [ev(s)]
        crf.setUnaryEnergy(hU[s]); m.append(ev(s))
        crf.addPairwiseGaussian(sxy=3, compat=3)
        crf.addPairwiseBilateral(sxy=80, srgb=13, rgbim=hI[s], compat=10); m.append(ev(s))
        crf.run(10); m.append(ev(s))
        crf.marginals(out=hQ[s]); m.append(ev(s))
        slots[s] = crf
        if record:
            marks.append(m)
    for c in slots:
        if c is not None:
            c.synchronize()
            c.close()
    return marks


run(2 * N_SLOTS, False)
torch.cuda.synchronize()
marks = run(STEPS, True)
torch.cuda.synchronize()
base = marks[0][0]
print("slots=%d   batch: start | H2D+layout end | build end | iterations end | D2H end   (ms since first event)" % N_SLOTS)
for t, m in enumerate(marks):
    ts = [base.elapsed_time(e) for e in m]
    print("batch %d slot %d: %7.2f | %7.2f | %7.2f | %7.2f | %7.2f   (h2d %.2f build %.2f iter %.2f d2h %.2f)" % (
        t, t % N_SLOTS, ts[0], ts[1], ts[2], ts[3], ts[4], ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3]))
tot = base.elapsed_time(marks[-1][4])
print("steady state: %.2f ms per batch (last D2H end - first start) / %d" % (tot / STEPS, STEPS))
