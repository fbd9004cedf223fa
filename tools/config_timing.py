"""Timing of the other BASELINE.json configs (device time of the full object lifecycle, host buffers)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from wsss_analysis_b200 import densecrf as G, synthetic as S

def run(name, sizes, L, n_iter, gs, gc, bs, srgb, bc, kind, reps=5):
    imgs = [getattr(S, kind + "_image")(h, w, i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, i) for i, (w, h) in enumerate(sizes)]
    U = np.concatenate([u.ravel() for u in Us]); I = np.concatenate([im.ravel() for im in imgs])
    npix = sum(w * h for w, h in sizes)
    ts = []
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = G.DenseCRFBatch(sizes, L)
        d.setUnaryEnergy(U); d.addPairwiseGaussian(sxy=gs, compat=gc); d.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=I, compat=bc)
        t1 = time.perf_counter()
        lab = d.map(n_iter)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        Mg, Mb = d.lattice_info(0)[1], d.lattice_info(1)[1]
        d.close(); ts.append((t1 - t0, t2 - t1))
    s, i = np.median([t[0] for t in ts]) * 1e3, np.median([t[1] for t in ts]) * 1e3
    print("%-34s N=%9d L=%2d it=%2d  M_g/N=%.3f M_b/N=%.3f  setup %.2f ms  iterate+labels %.2f ms  -> %.0f Mpix*iter/s (e2e, host in / labels out)" % (
        name, npix, L, n_iter, Mg / npix, Mb / npix, s, i, npix * n_iter / ((s + i) * 1e-3) / 1e6))

if __name__ == "__main__":
    run("SEC train 32 x 41x41 (config 2)", [(41, 41)] * 32, 21, 5, 3 / 12, 3, 80 / 12, 13, 10, "natural")
    run("SEC train 16 x 41x41", [(41, 41)] * 16, 21, 5, 3 / 12, 3, 80 / 12, 13, 10, "natural")
    run("HSN ADP-morph 1 x 1088^2 (config 3)", [(1088, 1088)], 29, 5, 1, 20, 10, 40, 50, "histo")
    run("HSN ADP-func 4 x 1088^2", [(1088, 1088)] * 4, 5, 5, 3, 40, 10, 4, 25, "histo")
    run("HSN as run 16 x 321^2", [(321, 321)] * 16, 21, 10, 1.5, 3, 40, 13, 10, "histo")
    run("DeepGlobe 1 x 2448^2 (config 4)", [(2448, 2448)], 6, 10, 3, 3, 80, 13, 10, "natural")
    run("DeepGlobe as run 8 x 612^2", [(612, 612)] * 8, 6, 10, 3, 3, 50, 5, 10, "natural")
    run("VOC 1 x 500x375 (config 0)", [(500, 375)], 21, 10, 3, 3, 80, 13, 10, "natural")
