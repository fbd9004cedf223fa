"""GPU: seeded random configurations (sizes, label counts, kernel widths, compatibilities, image
statistics) against the oracle: lattice integers bit-exact, Q within 1e-4, labels identical where the
oracle's top-2 margin is decided."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed, lo=2, hi=90):
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
    L = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 16, 17, 21, 24, 25, 29, 32, 33, 40]))
    kind = str(rng.choice(["natural", "iid", "histo", "flat"]))
    # every kernel width in the tree, incl. HSN VOC-M7's 3/12/4 and 80/12/4 (demo.py:161)
    g_sxy = float(rng.choice([3 / 12 / 4, 0.25, 1, 1.5, 3, 5]))
    b_sxy = float(rng.choice([80 / 12 / 4, 80 / 12, 10, 40, 50, 80]))
    b_srgb = float(rng.choice([4, 5, 13, 40]))
    n_iter = int(rng.integers(0, 6))
    return W, H, L, kind, g_sxy, float(rng.uniform(1, 20)), b_sxy, b_srgb, float(rng.uniform(1, 30)), n_iter


def _run(seed, case):
    from oracle import oracle as O
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S

    W, H, L, kind, g_sxy, g_c, b_sxy, b_srgb, b_c, n_iter = case
    img = np.full((H, W, 3), 77, np.uint8) if kind == "flat" else getattr(S, kind + "_image")(H, W, seed)
    U = S.random_unary(L, W * H, seed, sharp=1.5)
    o, g, gs = O.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L), G.DenseCRF2D(W, H, L)
    gs.set_arithmetic("strict")
    for m in (o, g, gs):
        m.setUnaryEnergy(U)
        m.addPairwiseGaussian(sxy=g_sxy, compat=g_c)
        m.addPairwiseBilateral(sxy=b_sxy, srgb=b_srgb, rgbim=img, compat=b_c)
    for k in range(2):
        eo, eg = o.lattice(k), g.lattice_export(k)
        assert eo.M == eg["M"], (seed, k)
        assert np.array_equal(eo.keys, eg["keys"])
        assert np.array_equal(eo.offsets, eg["offsets"])
        assert np.array_equal(eo.neighbours, eg["neighbours"])
        assert np.array_equal(eo.bary.view(np.uint32), eg["bary"].view(np.uint32))
    Qo, Qg, Qs = o.inference(n_iter), g.inference(n_iter), gs.inference(n_iter)
    assert np.abs(Qo - Qg).max() <= 1e-4, (seed, W, H, L, kind, n_iter, float(np.abs(Qo - Qg).max()))
    # "strict" arithmetic: every operation in the oracle's order => identical bits
    assert np.array_equal(Qo.view(np.uint32), Qs.view(np.uint32)), (seed, W, H, L, kind, n_iter)
    if L > 1:
        srt = np.sort(Qo, axis=0)
        decided = (srt[-1] - srt[-2]) > 1e-4
        assert (Qo.argmax(0) == g.map(n_iter))[decided].all()


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration(seed):
    _run(seed, _case(seed))


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_configuration_larger_images(seed):
    """Same draw at 150..340 pixels per side."""
    _run(seed, _case(seed, 150, 340))
