"""Diagnostic for the small configurations: host wall time of each API call of one VOC image (device
inputs), with and without a device synchronisation after the call, against the GPU-busy time of the
same calls -- tells whether the build is bound by the host's enqueue cost or by the device."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200 import densecrf as G
name = sys.argv[1] if len(sys.argv) > 1 else "voc1"
cfg = bench.CONFIGS[name]
dev = torch.device("cuda", 0)
imgs, unaries = bench.make_inputs(cfg)
U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).to(dev)
I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).to(dev)
Q = torch.empty(bench.npix(cfg) * cfg["L"], dtype=torch.float32, device=dev)
for sync in (True, False):
    acc = np.zeros(6)
    reps = 200
    for rep in range(reps + 20):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        def lap():
            if sync: torch.cuda.synchronize()
            t.append(time.perf_counter())
        crf = G.DenseCRFBatch(cfg["sizes"], cfg["L"], device=0); lap()
        crf.setUnaryEnergy(U); lap()
        crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"]); lap()
        crf.addPairwiseBilateral(sxy=cfg["b_sxy"], srgb=cfg["b_srgb"], rgbim=I, compat=cfg["b_compat"]); lap()
        crf.inference_device(cfg["iters"], out=Q); lap()
        torch.cuda.synchronize(); t.append(time.perf_counter())
        crf.close()
        if rep >= 20: acc += np.diff(t) * 1e3
    acc /= reps
    print("%s sync_after_each_call=%s: create %.3f unary %.3f gauss %.3f bilat %.3f infer(enqueue%s) %.3f drain %.3f | total %.3f ms" % (
        name, sync, *acc[:4], "+run" if sync else "", acc[4], acc[5], acc.sum()), flush=True)
