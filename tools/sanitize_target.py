"""Small workload touching every kernel family (uniform + mixed batches, flat image long rows, L<=2,
generic-G, exact mode, GPU unaries, resize, confusion) for compute-sanitizer runs."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from wsss_analysis_b200 import densecrf as G, synthetic as S, evaluation as E, wsss

def batch(sizes, L, imgs=None, exact=False, arith=None, persistent=None):
    imgs = imgs or [S.natural_image(h, w, i) for i, (w, h) in enumerate(sizes)]
    d = G.DenseCRFBatch(sizes, L)
    if exact: d.set_exact_arithmetic(True)
    if arith: d.set_arithmetic(arith)
    if persistent is not None: d.set_persistent(persistent)
    d.setUnaryEnergy([S.random_unary(L, w * h, i) for i, (w, h) in enumerate(sizes)])
    d.addPairwiseGaussian(sxy=3, compat=3); d.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
    q = d.inference(2); lab = d.map(2); d.close(); return q, lab

batch([(48, 36)] * 3, 21)                                   # uniform: replicated Gaussian lattice
batch([(48, 36), (30, 50)], 6)                              # mixed sizes
batch([(48, 36), (30, 50), (48, 36), (30, 50), (21, 17)], 6)  # repeated sizes: Gaussian lattice once per distinct size
batch([(64, 64)], 21, [np.full((64, 64, 3), 90, np.uint8)]) # flat: long-row tail kernel
batch([(40, 30)], 2); batch([(40, 30)], 1); batch([(20, 20)], 37); batch([(40, 30)], 5, exact=True)
for L_ in (13, 16, 20, 24, 28, 32):                          # cooperative splat / slice at G = 4, 5, 6, 7, 8
    batch([(30, 20), (17, 25)], L_)
# round 2: reference-association kernels (every lane-group width, long-row tails), the persistent
# cooperative kernel, uint8 labels, the wrappers on CUDA tensors
for L_ in (3, 6, 13, 21, 29, 37):
    batch([(30, 20), (17, 25)], L_, arith="reference")
batch([(64, 64)], 21, [np.full((64, 64, 3), 90, np.uint8)], arith="reference")
batch([(64, 64)], 21, [np.full((64, 64, 3), 90, np.uint8)], arith="strict")
for mode in ("fma", "reference"):
    for L_ in (6, 21, 29):
        batch([(30, 20), (17, 25)], L_, arith=mode, persistent=True)
    batch([(64, 64)], 21, [np.full((64, 64, 3), 90, np.uint8)], arith=mode, persistent=True)
d_ = G.DenseCRFBatch([(30, 20)], 5); d_.setUnaryEnergy([S.random_unary(5, 600, 0)]); d_.addPairwiseGaussian(sxy=3, compat=3)
d_.addPairwiseBilateral(sxy=40, srgb=13, rgbim=[S.natural_image(20, 30, 0)], compat=10); d_.map(2, dtype=np.uint8); d_.map_device(2, dtype=torch.uint8); d_.close()
dev_ = torch.device("cuda", 0)
t_img = torch.from_numpy(np.stack([S.natural_image(24, 32, i) for i in range(2)])).to(dev_)
wsss.dcrf_process(torch.from_numpy(np.stack([S.blob_probs(6, 24, 32, seed=i, n_active=3) for i in range(2)])).to(dev_), t_img, (1.5, 3, 40, 13, 10, 3.0))
wsss.sec_crf_layer(torch.randn((2, 24, 32, 5), device=dev_), t_img.float(), {"g_sxy": 0.25, "g_compat": 3, "bi_sxy": 80 / 12, "bi_srgb": 13, "bi_compat": 10, "iterations": 2}, 5)
wsss.crf_inference_label_batch(t_img.float(), torch.randint(0, 4, (2, 24, 32), device=dev_), n_labels=4)
x_ = np.linspace(-104, 0, 1000).astype(np.float32); y_ = np.empty_like(x_)
from wsss_analysis_b200 import _lib
_lib.check(_lib.load().dcrf_expf_ref(x_.ctypes.data, y_.ctypes.data, x_.size, -1))
from wsss_analysis_b200.pipeline import BatchPipeline, pinned_empty   # async-host upload stream, sub-batches
sz = [(32, 24)] * 5; n_ = 5 * 32 * 24
U_ = pinned_empty(n_ * 4); U_[:] = np.concatenate([S.random_unary(4, 32 * 24, i).ravel() for i in range(5)])
I_ = pinned_empty(n_ * 3, np.uint8); I_[:] = np.concatenate([S.natural_image(24, 32, i).ravel() for i in range(5)])
with BatchPipeline(n_slots=2, chunk_images=2) as pipe:
    pipe.result(pipe.submit(sz, 4, U_, I_, {"g_sxy": 3, "g_compat": 3, "bi_sxy": 40, "bi_srgb": 13, "bi_compat": 10, "iterations": 2}, out=pinned_empty(n_ * 4)))
probs = np.stack([S.blob_probs(6, 24, 32, seed=i, n_active=3) for i in range(2)])
wsss.dcrf_process(probs, np.stack([S.histo_image(24, 32, i) for i in range(2)]), (1.5, 3, 40, 13, 10, 3.0))
wsss.crf_inference_label(S.natural_image(24, 32, 0).astype(np.float32), S.gt_map(24, 32, 4, 0, ignore=0), "voc12", n_labels=4)
wsss.sec_crf_layer(np.random.default_rng(0).standard_normal((2, 41, 41, 5)).astype(np.float32), np.stack([S.natural_image(41, 41, i) for i in range(2)]).astype(np.float32),
                   {"g_sxy": 0.25, "g_compat": 3, "bi_sxy": 80 / 12, "bi_srgb": 13, "bi_compat": 10, "iterations": 2}, 5)
acc = E.ConfusionAccumulator(6); lab = S.gt_map(40, 40, 6, 0, ignore=255)
acc.update(lab, E.resize_nearest(S.gt_map(10, 10, 6, 1, ignore=0), (40, 40))); E.resize_bilinear(np.ones((5, 7, 3), np.float32), (20, 11))
print("sanitize target ok", acc.result().sum())
