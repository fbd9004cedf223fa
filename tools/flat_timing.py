import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from wsss_analysis_b200 import densecrf as G, synthetic as S
W, H, L = 500, 375, 21
for name, img in (("flat", np.full((H, W, 3), 200, np.uint8)), ("two-tone", np.where((np.arange(W)[None, :, None] < 250), 60, 200).astype(np.uint8) * np.ones((H, 1, 3), np.uint8)), ("natural", S.natural_image(H, W, 0))):
    img = np.ascontiguousarray(img)
    U = S.random_unary(L, W * H, 0)
    for rep in range(3):
        d = G.DenseCRF2D(W, H, L); d.setUnaryEnergy(U); d.addPairwiseGaussian(sxy=3, compat=3); d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
        torch.cuda.synchronize(); t0 = time.perf_counter(); Q = d.inference(10); t1 = time.perf_counter()
    print(name, "M_b", d.lattice_info(1)[1], "inference(10) %.2f ms" % ((t1 - t0) * 1e3))
