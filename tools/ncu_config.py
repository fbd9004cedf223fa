"""Steps of one bench configuration with device-resident inputs, for ncu launch lists / captures.
usage: ncu_config.py <config> [steps] [arith]"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200 import densecrf as G

name = sys.argv[1] if len(sys.argv) > 1 else "voc32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = bench.CONFIGS[name]
dev = torch.device("cuda", 0)
imgs, unaries = bench.make_inputs(cfg)
U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).to(dev)
I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).to(dev)
Q = torch.empty(bench.npix(cfg) * cfg["L"], dtype=torch.float32, device=dev)
torch.cuda.synchronize()
import time
for it in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    crf = G.DenseCRFBatch(cfg["sizes"], cfg["L"], device=0)
    if len(sys.argv) > 3:
        crf.set_arithmetic(sys.argv[3])
    crf.setUnaryEnergy(U)
    t1 = time.perf_counter()
    crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    crf.addPairwiseBilateral(sxy=cfg["b_sxy"], srgb=cfg["b_srgb"], rgbim=I, compat=cfg["b_compat"])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    crf.inference_device(cfg["iters"], out=Q)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    crf.close()
    print("step %d: create+unary %.2f gauss %.2f bilat %.2f infer %.2f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
