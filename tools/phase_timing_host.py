"""Diagnostic: host-timed phases of one batched step with pinned HOST buffers (the e2e leg)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200 import densecrf as G
B = 32
imgs, unaries = bench.make_inputs(B)
sizes = [(bench.W_IMG, bench.H_IMG)] * B
U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).pin_memory()
I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).pin_memory()
Q = torch.empty(B * bench.L_LAB * bench.W_IMG * bench.H_IMG, dtype=torch.float32).pin_memory()
Un, In, Qn = U.numpy(), I.numpy(), Q.numpy()
for rep in range(4):
    t = [time.perf_counter()]
    def lap():
        torch.cuda.synchronize(); t.append(time.perf_counter())
    crf = G.DenseCRFBatch(sizes, bench.L_LAB, device=0); lap()
    crf.setUnaryEnergy(Un); lap()
    crf.addPairwiseGaussian(sxy=3, compat=3); lap()
    crf.addPairwiseBilateral(sxy=80, srgb=13, rgbim=In, compat=10); lap()
    crf.inference(10, out=Qn); lap()
    crf.close(); lap()
    d = np.diff(t) * 1e3
    print("host rep%d: create %.2f unary %.2f gauss %.2f bilat %.2f infer %.2f close %.2f | total %.2f ms" % (rep, *d, d.sum()))
