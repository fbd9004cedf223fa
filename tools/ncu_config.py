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
for it in range(steps):
    crf = G.DenseCRFBatch(cfg["sizes"], cfg["L"], device=0)
    if len(sys.argv) > 3:
        crf.set_arithmetic(sys.argv[3])
    crf.setUnaryEnergy(U)
    crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
    crf.addPairwiseBilateral(sxy=cfg["b_sxy"], srgb=cfg["b_srgb"], rgbim=I, compat=cfg["b_compat"])
    crf.inference_device(cfg["iters"], out=Q)
    crf.close()
    print("step", it, "done", float(Q[:100].sum()), flush=True)
