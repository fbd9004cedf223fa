import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from oracle import oracle as O
from wsss_analysis_b200 import densecrf as G, synthetic as S
W = H = 612; L = 6
img = S.natural_image(H, W, 4); U = S.random_unary(L, W * H, 4)
o, g, gx = O.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L)
gx.set_exact_arithmetic(True)
for m in (o, g, gx):
    m.setUnaryEnergy(U); m.addPairwiseGaussian(sxy=3, compat=3); m.addPairwiseBilateral(sxy=50, srgb=5, rgbim=img, compat=10)
for n in (1, 2, 5, 10):
    Qo, Qf, Qx = o.inference(n), g.inference(n), gx.inference(n)
    for name, Q in (("fast", Qf), ("exact", Qx)):
        d = np.abs(Qo - Q).max(0)
        print("iters %2d %5s: max %.3g  p99.99 %.3g  p99.9 %.3g  #>1e-4 %d  #>1e-5 %d  argmax agree %.6f" % (
            n, name, d.max(), np.percentile(d, 99.99), np.percentile(d, 99.9), (d > 1e-4).sum(), (d > 1e-5).sum(),
            (Qo.argmax(0) == Q.argmax(0)).mean()))
