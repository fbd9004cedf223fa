"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle
on the same seeded inputs.  PARITY UNPINNED: the oracle is this repo's restatement of pydensecrf
(see oracle/densecrf_oracle.c); tolerances are BASELINE.json's: lattice integers bit-exact,
Q max-abs <= 1e-4, argmax agreement >= 99.9 %."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q_TOL = 1e-4
AGREE = 0.999


@pytest.fixture(scope="module")
def mods():
    from oracle import oracle as O
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S

    return O, G, S


def _pair(mods, W, H, L, gs, bs, srgb, kind="natural", seed=0, gc=3, bc=10, **kw):
    O, G, S = mods
    img = getattr(S, kind + "_image")(H, W, seed)
    U = S.random_unary(L, W * H, seed)
    o, g = O.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L)
    for m in (o, g):
        m.setUnaryEnergy(U)
        m.addPairwiseGaussian(sxy=gs, compat=gc, **kw)
        m.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=img, compat=bc, **kw)
    return o, g


def _check_lattice(o, g, k):
    eo, eg = o.lattice(k), g.lattice_export(k)
    assert eo.M == eg["M"]
    assert np.array_equal(eo.keys, eg["keys"])
    assert np.array_equal(eo.offsets, eg["offsets"])
    assert np.array_equal(eo.bary.view(np.uint32), eg["bary"].view(np.uint32))
    assert np.array_equal(eo.neighbours, eg["neighbours"])
    np.testing.assert_allclose(eg["norm"], o.norm(k), rtol=2e-6, atol=0)


CASES = [
    # W, H, L, gauss sxy, bilateral sxy, srgb, image kind
    (64, 48, 5, 3, 20, 13, "natural"),
    (41, 41, 21, 3 / 12, 80 / 12, 13, "natural"),      # SEC train config (SEC.py:19)
    (97, 61, 21, 3, 80, 13, "iid"),
    (128, 96, 6, 3, 80, 13, "histo"),
    (75, 50, 29, 1, 10, 40, "histo"),                    # ADP-morph test config (SEC.py:24-25)
    (33, 17, 1, 3, 50, 5, "natural"),                    # L = 1
    (50, 40, 2, 3, 50, 5, "natural"),                    # L = 2 (value_size <= 2 association)
    (60, 45, 3, (3, 5), (60, 40), (5, 9, 13), "natural"),  # per-axis sigmas
]


@pytest.mark.parametrize("case", CASES)
def test_lattice_bit_exact(mods, case):
    o, g = _pair(mods, *case)
    _check_lattice(o, g, 0)
    _check_lattice(o, g, 1)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("n_iter", [0, 1, 5])
def test_inference_parity(mods, case, n_iter):
    o, g = _pair(mods, *case)
    Qo, Qg = o.inference(n_iter), g.inference(n_iter)
    assert Qg.shape == Qo.shape and Qg.dtype == np.float32
    assert np.abs(Qo - Qg).max() <= Q_TOL
    assert (Qo.argmax(0) == Qg.argmax(0)).mean() >= AGREE
    np.testing.assert_allclose(Qg.sum(0), 1.0, atol=1e-5)


def test_voc_full_size_parity(mods):
    """BASELINE config 1: 500x375, 21 labels, 10 iterations, sxy 3 / 80, srgb 13, compat 3 / 10."""
    o, g = _pair(mods, 500, 375, 21, 3, 80, 13)
    _check_lattice(o, g, 0)
    _check_lattice(o, g, 1)
    Qo, Qg = o.inference(10), g.inference(10)
    assert np.abs(Qo - Qg).max() <= Q_TOL
    assert (Qo.argmax(0) == Qg.argmax(0)).mean() >= AGREE
    lab = g.map(10)
    assert np.array_equal(lab, Qg.argmax(0).astype(np.int32))


@pytest.mark.parametrize("vs", [1, 2, 3, 21])
def test_lattice_filter_parity(mods, vs):
    O, G, S = mods
    W, H = 80, 60
    img = S.natural_image(H, W, 3)
    o, g = O.DenseCRF2D(W, H, 4), G.DenseCRF2D(W, H, 4)
    for m in (o, g):
        m.addPairwiseBilateral(sxy=30, srgb=10, rgbim=img, compat=1)
    rng = np.random.default_rng(5)
    v = rng.random((vs, W * H)).astype(np.float32)
    lat = O.Lattice(np.stack([np.tile(np.arange(W), H) / np.float32(30), np.repeat(np.arange(H), W) / np.float32(30),
                              img[..., 0].ravel() / np.float32(10), img[..., 1].ravel() / np.float32(10),
                              img[..., 2].ravel() / np.float32(10)]).astype(np.float32))
    ref = lat.compute(v)
    out = g.lattice_filter(0, v)
    # splat rows are summed in the oracle's order, blur and slice use the same association:
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("ntype", [0, 1, 2, 3])
def test_normalization_types(mods, ntype):
    o, g = _pair(mods, 48, 36, 4, 3, 30, 13, normalization=ntype)
    Qo, Qg = o.inference(3), g.inference(3)
    assert np.abs(Qo - Qg).max() <= Q_TOL


def test_compat_kinds(mods):
    O, G, S = mods
    W, H, L = 40, 30, 5
    rng = np.random.default_rng(11)
    img = S.natural_image(H, W, 2)
    U = S.random_unary(L, W * H, 2)
    diag = -rng.uniform(0.5, 3, L).astype(np.float32)
    mat = -rng.uniform(0, 2, (L, L)).astype(np.float32)
    o, g = O.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L)
    for m in (o, g):
        m.setUnaryEnergy(U)
        m.addPairwiseGaussian(sxy=3, compat=diag)
        m.addPairwiseBilateral(sxy=30, srgb=13, rgbim=img, compat=mat)
    Qo, Qg = o.inference(4), g.inference(4)
    assert np.abs(Qo - Qg).max() <= Q_TOL


def test_add_pairwise_energy_and_densecrf_nd(mods):
    O, G, S = mods
    from wsss_analysis_b200 import utils

    W, H, L = 36, 28, 4
    img = S.natural_image(H, W, 4)
    U = S.random_unary(L, W * H, 4)
    fg = utils.create_pairwise_gaussian((3, 3), (H, W))
    fb = utils.create_pairwise_bilateral((20, 20), (13, 13, 13), img, chdim=2)
    o, g = O.DenseCRF(W * H, L), G.DenseCRF(W * H, L)
    for m in (o, g):
        m.setUnaryEnergy(U)
        m.addPairwiseEnergy(np.ascontiguousarray(fg), compat=3)
        m.addPairwiseEnergy(np.ascontiguousarray(fb), compat=10)
    assert np.abs(o.inference(3) - g.inference(3)).max() <= Q_TOL
    eo, eg = o.lattice(1), g.lattice_export(1)
    assert eo.M == eg["M"] and np.array_equal(eo.offsets, eg["offsets"])


def test_stepping_api_and_kl(mods):
    o, g = _pair(mods, 40, 30, 5, 3, 30, 13)
    Qo, _, _ = o.startInference()
    Qg, t1, t2 = g.startInference()
    assert np.abs(Qo - Qg).max() <= 1e-6
    for _ in range(3):
        o.stepInference(Qo)
        g.stepInference(Qg, t1, t2)
    assert np.abs(Qo - Qg).max() <= Q_TOL
    klo, klg = o.klDivergence(Qo), g.klDivergence(Qg)
    assert abs(klo - klg) <= 1e-4 * max(1.0, abs(klo))
    assert np.abs(g.inference(3) - Qg).max() <= 1e-6


def test_run_to_run_determinism(mods):
    _, g1 = _pair(mods, 120, 90, 21, 3, 80, 13)
    _, g2 = _pair(mods, 120, 90, 21, 3, 80, 13)
    a, b, c = g1.inference(5), g2.inference(5), g1.inference(5)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


def test_batch_matches_single_images(mods):
    O, G, S = mods
    sizes = [(64, 48), (41, 41), (50, 70), (64, 48)]
    L = 6
    imgs = [S.natural_image(h, w, 10 + i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, 10 + i) for i, (w, h) in enumerate(sizes)]
    gb = G.DenseCRFBatch(sizes, L)
    gb.setUnaryEnergy(Us)
    gb.addPairwiseGaussian(sxy=3, compat=3)
    gb.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
    Qb = gb.inference(5)
    _, _, per = gb.lattice_info(1)
    labs = gb.map(5)
    for i, (w, h) in enumerate(sizes):
        o = O.DenseCRF2D(w, h, L)
        o.setUnaryEnergy(Us[i])
        o.addPairwiseGaussian(sxy=3, compat=3)
        o.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs[i], compat=10)
        Qo = o.inference(5)
        assert per[i] == o.lattice(1).M
        eg = gb.lattice_export(1, image=i)
        eo = o.lattice(1)
        assert np.array_equal(eo.offsets, eg["offsets"]) and np.array_equal(eo.neighbours, eg["neighbours"])
        assert np.abs(Qo - Qb[i]).max() <= Q_TOL
        assert labs[i].shape == (h, w)
        assert (labs[i].ravel() == Qb[i].argmax(0)).all()


def test_torch_device_handoff(mods):
    import torch

    O, G, S = mods
    W, H, L = 64, 48, 5
    img = S.natural_image(H, W, 1)
    U = S.random_unary(L, W * H, 1)
    g = G.DenseCRF2D(W, H, L)
    g.setUnaryEnergy(torch.from_numpy(U).cuda())
    g.addPairwiseGaussian(sxy=3, compat=3)
    g.addPairwiseBilateral(sxy=20, srgb=13, rgbim=torch.from_numpy(img).cuda(), compat=10)
    Qd = g.inference_device(4)
    assert Qd.is_cuda and Qd.shape == (L, W * H)
    h = G.DenseCRF2D(W, H, L)
    h.setUnaryEnergy(U)
    h.addPairwiseGaussian(sxy=3, compat=3)
    h.addPairwiseBilateral(sxy=20, srgb=13, rgbim=img, compat=10)
    assert np.array_equal(Qd.cpu().numpy(), h.inference(4))


def test_error_behaviour(mods):
    O, G, S = mods
    g = G.DenseCRF2D(20, 10, 3)
    with pytest.raises(ValueError, match="Bad shape for unary energy"):
        g.setUnaryEnergy(np.zeros((3, 199), np.float32))
    with pytest.raises(ValueError):
        g.setUnaryEnergy(np.zeros((3, 200), np.float64))
    with pytest.raises(ValueError):
        g.setUnaryEnergy(np.zeros((200, 3), np.float32).T)
    with pytest.raises(ValueError, match="Bad shape for pairwise bilateral"):
        g.addPairwiseBilateral(sxy=3, srgb=3, rgbim=np.zeros((20, 10, 3), np.uint8), compat=1)
    with pytest.raises(ValueError):
        g.addPairwiseBilateral(sxy=3, srgb=3, rgbim=np.zeros((10, 20, 3), np.float32), compat=1)
    # the never-used L = 0 constructor of dcrf_process must not crash (03c_hsn/utilities.py:427-428)
    G.DenseCRF2D(20, 10, 0)
    # float-valued iteration count (03c_hsn/demo.py:159)
    g.setUnaryEnergy(np.zeros((3, 200), np.float32))
    assert g.inference(np.float64(2.0)).shape == (3, 200)


def test_confusion_bit_exact(mods):
    import torch

    O, G, S = mods
    from wsss_analysis_b200 import evaluation as E

    rng = np.random.default_rng(0)
    C_ = 21
    gt = rng.integers(0, C_, 100000).astype(np.int32)
    gt[rng.random(gt.size) < 0.05] = 255
    gt[rng.random(gt.size) < 0.01] = -1
    pred = rng.integers(0, C_, gt.size).astype(np.int32)
    acc = E.ConfusionAccumulator(C_)
    acc.update(torch.from_numpy(gt[:60000]).cuda(), torch.from_numpy(pred[:60000]).cuda())
    acc.update(gt[60000:], pred[60000:])
    conf = acc.result()
    ref = O.confusion(gt, pred, C_)
    assert np.array_equal(conf, ref)
    m = gt.astype(np.int64)
    valid = (m >= 0) & (m < C_)
    ref2 = np.bincount(C_ * m[valid] + pred[valid], minlength=C_ * C_).reshape(C_, C_)
    assert np.array_equal(conf[:C_], ref2)


def test_uniform_batch_shares_the_gaussian_lattice(mods):
    """All images of one size: the position-only lattice is built once and replicated; every image
    must still export exactly the oracle's lattice, and Q must match image by image."""
    O, G, S = mods
    sizes = [(60, 44)] * 5
    L = 7
    imgs = [S.natural_image(h, w, 30 + i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, 30 + i) for i, (w, h) in enumerate(sizes)]
    gb = G.DenseCRFBatch(sizes, L)
    gb.setUnaryEnergy(Us)
    gb.addPairwiseGaussian(sxy=3, compat=3)
    gb.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
    Qb = gb.inference(5)
    for i, (w, h) in enumerate(sizes):
        o = O.DenseCRF2D(w, h, L)
        o.setUnaryEnergy(Us[i])
        o.addPairwiseGaussian(sxy=3, compat=3)
        o.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs[i], compat=10)
        for k in range(2):
            eo, eg = o.lattice(k), gb.lattice_export(k, image=i)
            assert eo.M == eg["M"]
            assert np.array_equal(eo.keys, eg["keys"]) and np.array_equal(eo.offsets, eg["offsets"])
            assert np.array_equal(eo.neighbours, eg["neighbours"])
            assert np.array_equal(eo.bary.view(np.uint32), eg["bary"].view(np.uint32))
            np.testing.assert_allclose(eg["norm"], o.norm(k), rtol=2e-6, atol=0)
        assert np.abs(o.inference(5) - Qb[i]).max() <= Q_TOL


@pytest.mark.parametrize("arith", ["fma", "strict"])
def test_mixed_batch_builds_the_gaussian_lattice_once_per_size(mods, arith):
    """A batch with repeated sizes among distinct ones (PASCAL VOC val: 500x375, 500x333, 375x500, ...):
    the position-only lattice is built once per DISTINCT size and replicated with per-image id shifts.
    Every image must export exactly the oracle's lattice (keys, offsets, barycentric bits, neighbours,
    norm) and Q must match image by image -- bit for bit in the strict reference arithmetic."""
    O, G, S = mods
    sizes = [(60, 44), (37, 52), (60, 44), (41, 41), (37, 52), (60, 44), (64, 30)]
    L = 7
    imgs = [S.natural_image(h, w, 50 + i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, 50 + i) for i, (w, h) in enumerate(sizes)]
    gb = G.DenseCRFBatch(sizes, L)
    gb.set_arithmetic(arith)
    gb.setUnaryEnergy(Us)
    gb.addPairwiseGaussian(sxy=3, compat=3)
    gb.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
    Qb = gb.inference(5)
    for i, (w, h) in enumerate(sizes):
        o = O.DenseCRF2D(w, h, L)
        o.setUnaryEnergy(Us[i])
        o.addPairwiseGaussian(sxy=3, compat=3)
        o.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs[i], compat=10)
        for k in range(2):
            eo, eg = o.lattice(k), gb.lattice_export(k, image=i)
            assert eo.M == eg["M"]
            assert np.array_equal(eo.keys, eg["keys"]) and np.array_equal(eo.offsets, eg["offsets"])
            assert np.array_equal(eo.neighbours, eg["neighbours"])
            assert np.array_equal(eo.bary.view(np.uint32), eg["bary"].view(np.uint32))
            np.testing.assert_allclose(eg["norm"], o.norm(k), rtol=2e-6, atol=0)
        Qo = o.inference(5)
        if arith == "strict":
            assert np.array_equal(Qo.view(np.uint32), Qb[i].view(np.uint32))
        else:
            assert np.abs(Qo - Qb[i]).max() <= Q_TOL


@pytest.mark.parametrize("terms", ["gauss", "bilat", "gauss+bilat+bilat", "bilat+gauss", "energy3d"])
def test_other_term_combinations(mods, terms):
    """Anything but (Gaussian, bilateral) takes the generic fast slice: one term, three terms,
    reversed order, a 3-D position-only kernel through addPairwiseEnergy."""
    O, G, S = mods
    from wsss_analysis_b200 import utils

    W, H, L = 44, 33, 7
    img = S.natural_image(H, W, 21)
    U = S.random_unary(L, W * H, 21)
    o, g = O.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L)
    for m in (o, g):
        m.setUnaryEnergy(U)
        for t in terms.split("+"):
            if t == "gauss":
                m.addPairwiseGaussian(sxy=2, compat=4)
            elif t == "bilat":
                m.addPairwiseBilateral(sxy=30, srgb=11, rgbim=img, compat=7)
            else:
                f = utils.create_pairwise_bilateral((5, 5), (20,), img[..., 0], chdim=-1)
                m.addPairwiseEnergy(np.ascontiguousarray(f), compat=5)
    assert g.num_pairwise() == len(terms.split("+"))
    Qo, Qg = o.inference(4), g.inference(4)
    assert np.abs(Qo - Qg).max() <= Q_TOL
    assert (Qo.argmax(0) == Qg.argmax(0)).mean() >= AGREE


def test_c_abi_error_codes_and_messages(mods):
    """Direct ctypes calls: bad arguments come back as DCRF_EINVAL / DCRF_ESTATE with a message,
    never as a crash (include/dcrf_b200.h error contract)."""
    import ctypes as C

    from wsss_analysis_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dcrf_create(0, 10, 3, -1, None, C.byref(h)) == _lib.DCRF_EINVAL
    assert b"width/height" in lib.dcrf_last_error()
    assert lib.dcrf_create(10, 10, 129, -1, None, C.byref(h)) == _lib.DCRF_EINVAL
    assert lib.dcrf_create(10, 10, 3, 9999, None, C.byref(h)) == _lib.DCRF_EINVAL
    assert lib.dcrf_create(10, 10, 3, -1, None, C.byref(h)) == _lib.DCRF_OK
    potts = np.array([1.0], np.float32)
    assert lib.dcrf_set_unary(h, None, 0) == _lib.DCRF_EINVAL
    assert lib.dcrf_add_pairwise_gaussian(h, 3.0, 3.0, 7, potts.ctypes.data, 1, 3) == _lib.DCRF_EINVAL
    assert lib.dcrf_add_pairwise_gaussian(h, 3.0, 3.0, 0, potts.ctypes.data, 1, 9) == _lib.DCRF_EINVAL
    assert lib.dcrf_add_pairwise_gaussian(h, 3.0, 3.0, 0, None, 1, 3) == _lib.DCRF_EINVAL
    assert lib.dcrf_step_inference(h) == _lib.DCRF_ESTATE          # before startInference
    q = np.zeros((3, 100), np.float32)
    assert lib.dcrf_get_q(h, q.ctypes.data, 0) == _lib.DCRF_ESTATE
    assert lib.dcrf_inference(h, -1, q.ctypes.data, 0) == _lib.DCRF_EINVAL
    assert lib.dcrf_lattice_info(h, 0, None, None, None) == _lib.DCRF_EINVAL  # no pairwise term yet
    feats = np.zeros((8, 100), np.float32)
    assert lib.dcrf_add_pairwise_energy(h, feats.ctypes.data, 8, 0, 0, potts.ctypes.data, 1, 3) == _lib.DCRF_EINVAL
    for _ in range(4):
        assert lib.dcrf_add_pairwise_gaussian(h, 3.0, 3.0, 0, potts.ctypes.data, 1, 3) == _lib.DCRF_OK
    assert lib.dcrf_add_pairwise_gaussian(h, 3.0, 3.0, 0, potts.ctypes.data, 1, 3) == _lib.DCRF_EINVAL  # max 4 terms
    assert lib.dcrf_inference(h, 1, q.ctypes.data, 0) == _lib.DCRF_OK
    np.testing.assert_allclose(q.sum(0), 1.0, atol=1e-5)
    lib.dcrf_destroy(h)
    lib.dcrf_destroy(None)  # no-op
    assert lib.dcrf_launch_count() > 0


def test_lsd_sort_path_gives_the_same_rows():
    """The CSR rows normally come from one radix pass + the bucket kernel; images with more than 2^23
    lattice vertices fall back to the multi-pass LSD sort + finalising kernel.  DCRF_SORT_LSD=1 forces that
    path (read once per process, hence the subprocesses): marginals of a mixed batch must be bit-identical."""
    import hashlib
    import os
    import subprocess
    import sys

    code = (
        "import hashlib, numpy as np\n"
        "from wsss_analysis_b200 import densecrf as G, synthetic as S\n"
        "sizes = [(70, 50), (33, 61), (70, 50)]\n"
        "imgs = [S.natural_image(h, w, 7 + i) for i, (w, h) in enumerate(sizes)]\n"
        "imgs[1][:] = 120\n"   # a flat image: very long rows
        "Us = [S.random_unary(6, w * h, 7 + i) for i, (w, h) in enumerate(sizes)]\n"
        "d = G.DenseCRFBatch(sizes, 6); d.set_arithmetic('reference'); d.setUnaryEnergy(Us)\n"
        "d.addPairwiseGaussian(sxy=3, compat=3); d.addPairwiseBilateral(sxy=30, srgb=13, rgbim=imgs, compat=10)\n"
        "Q = d.inference(4)\n"
        "print(hashlib.sha256(b''.join(np.ascontiguousarray(q).tobytes() for q in Q)).hexdigest())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for lsd in ("0", "1"):
        env = dict(os.environ, DCRF_SORT_LSD=lsd, PYTHONPATH=root)
        out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append(out.stdout.strip().splitlines()[-1])
    assert len(digests[0]) == 64 and digests[0] == digests[1]
