"""Diagnostic (not a test): GPU vs oracle diffs + rough timings on one VOC-shaped image."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from wsss_analysis_b200 import densecrf as G, synthetic as S
from oracle import oracle as O


def run(W, H, L, n_iter, gs, bs, srgb, img_kind="natural", seed=0):
    img = getattr(S, img_kind + "_image")(H, W, seed)
    U = S.random_unary(L, W * H, seed)
    t0 = time.time()
    o = O.DenseCRF2D(W, H, L); o.setUnaryEnergy(U)
    o.addPairwiseGaussian(sxy=gs, compat=3); o.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=img, compat=10)
    Qo = o.inference(n_iter); t_cpu = time.time() - t0
    t0 = time.time()
    g = G.DenseCRF2D(W, H, L); g.setUnaryEnergy(U)
    g.addPairwiseGaussian(sxy=gs, compat=3); g.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=img, compat=10)
    Qg = g.inference(n_iter); t_gpu = time.time() - t0
    for k in range(2):
        eo, eg = o.lattice(k), g.lattice_export(k)
        ok = dict(M=eo.M == eg["M"])
        if ok["M"]:
            ok["keys"] = np.array_equal(eo.keys, eg["keys"])
            ok["offsets"] = np.array_equal(eo.offsets, eg["offsets"])
            ok["bary"] = np.array_equal(eo.bary.view(np.uint32), eg["bary"].view(np.uint32))
            ok["neigh"] = np.array_equal(eo.neighbours, eg["neighbours"])
            ok["norm_maxrel"] = float(np.abs(o.norm(k) / eg["norm"] - 1).max())
        print("  lattice", k, "M", eo.M, eg["M"], ok)
    d = np.abs(Qo - Qg)
    agree = (Qo.argmax(0) == Qg.argmax(0)).mean()
    print("%dx%dx%d it=%d %s: max|dQ|=%.3g argmax agree=%.5f cpu %.2fs gpu(first call) %.3fs" % (
        W, H, L, n_iter, img_kind, d.max(), agree, t_cpu, t_gpu))
    # steady-state GPU timing of the full object lifecycle
    for rep in range(3):
        t0 = time.time()
        g = G.DenseCRF2D(W, H, L); g.setUnaryEnergy(U)
        g.addPairwiseGaussian(sxy=gs, compat=3); g.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=img, compat=10)
        t1 = time.time()
        Qg2 = g.inference(n_iter); t2 = time.time()
    print("  steady: setup %.2f ms, inference(%d)+D2H %.2f ms; rerun identical: %s" % (
        (t1 - t0) * 1e3, n_iter, (t2 - t1) * 1e3, np.array_equal(Qg, Qg2)))


if __name__ == "__main__":
    run(64, 48, 5, 3, 3, 20, 13)
    run(41, 41, 21, 5, 3 / 12, 80 / 12, 13)
    run(500, 375, 21, 10, 3, 80, 13)
    run(500, 375, 21, 10, 3, 80, 13, "iid")
    run(321, 321, 2, 5, 1.5, 40, 13, "histo")
    run(320, 240, 29, 5, 1, 10, 40, "histo")
