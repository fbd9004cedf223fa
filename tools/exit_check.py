"""Exit-path check: small handles (concurrent lattice builds) in the situations a process can end in."""
import sys, os, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from wsss_analysis_b200 import densecrf as G, synthetic as S
mode = sys.argv[1]
W, H, L = 64, 48, 5
img = S.natural_image(H, W, 1); U = S.random_unary(L, W * H, 1)
def make():
    g = G.DenseCRF2D(W, H, L); g.setUnaryEnergy(U)
    g.addPairwiseGaussian(sxy=3, compat=3); g.addPairwiseBilateral(sxy=40, srgb=13, rgbim=img, compat=10)
    return g
if mode == "closed":
    g = make(); g.inference(2); g.close()
elif mode == "unclosed":
    g = make(); g.inference(2)
elif mode == "pending_closed":
    g = make(); g.close()
elif mode == "pending_unclosed":
    g = make()
elif mode == "many":
    hs = [make() for _ in range(5)]
    for h in hs: h.inference(1)
    for h in hs: h.close()
elif mode == "pipeline":
    from wsss_analysis_b200.pipeline import BatchPipeline
    cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 40, "bi_srgb": 13, "bi_compat": 10, "iterations": 2}
    with BatchPipeline(n_slots=3) as pipe:
        t = [pipe.submit([(W, H)] * 2, L, [U, U], [img, img], cfg) for _ in range(5)]
        for x in t: pipe.result(x)
print("done", mode, flush=True)
