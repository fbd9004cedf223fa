"""Diagnostic: host-timed phases of one batched step (device-resident inputs)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200 import densecrf as G

def main(B, profile):
    dev = torch.device("cuda", 0)
    imgs, unaries = bench.make_inputs(B)
    sizes = [(bench.W_IMG, bench.H_IMG)] * B
    U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).to(dev)
    I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).to(dev)
    Q = torch.empty(B * bench.L_LAB * bench.W_IMG * bench.H_IMG, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev); torch.cuda.synchronize(); torch.cuda.set_stream(stream)
    for rep in range(4):
        t = [time.perf_counter()]
        def lap():
            torch.cuda.synchronize(); t.append(time.perf_counter())
        crf = G.DenseCRFBatch(sizes, bench.L_LAB, device=0, stream=stream); lap()
        if profile: crf.profile_enable(True)
        crf.setUnaryEnergy(U); lap()
        crf.addPairwiseGaussian(sxy=3, compat=3); lap()
        crf.addPairwiseBilateral(sxy=80, srgb=13, rgbim=I, compat=10); lap()
        crf.inference_device(10, out=Q); lap()
        crf.close(); lap()
        d = np.diff(t) * 1e3
        print("B=%d prof=%d rep%d: create %.2f unary %.2f gauss %.2f bilat %.2f infer %.2f close %.2f | total %.2f ms" % (
            B, profile, rep, *d, d.sum()))

if __name__ == "__main__":
    for B in (1, 8, 32):
        main(B, False)
    main(32, True)
