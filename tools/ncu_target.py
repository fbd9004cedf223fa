"""One batched step (B images, device-resident) for ncu captures.  usage: ncu_target.py [B] [iters]"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200 import densecrf as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
imgs, unaries = bench.make_inputs(B)
sizes = [(bench.W_IMG, bench.H_IMG)] * B
U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).to(dev)
I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).to(dev)
Q = torch.empty(B * bench.L_LAB * bench.W_IMG * bench.H_IMG, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
crf = G.DenseCRFBatch(sizes, bench.L_LAB, device=0)
crf.setUnaryEnergy(U)
crf.addPairwiseGaussian(sxy=3, compat=3)
crf.addPairwiseBilateral(sxy=80, srgb=13, rgbim=I, compat=10)
crf.inference_device(iters, out=Q)
print("done", float(Q[:100].sum()))
