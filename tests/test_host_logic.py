"""CPU: host-side mirrors of the reference interface (NumPy glue, metric conventions, sharding)."""
import numpy as np
import pytest

from oracle import oracle as O
from wsss_analysis_b200 import evaluation as E
from wsss_analysis_b200 import synthetic as S
from wsss_analysis_b200 import utils, wsss


def test_unary_from_softmax_matches_restatement():
    rng = np.random.default_rng(0)
    sm = rng.random((5, 7, 9))
    sm /= sm.sum(0)
    sm[0, 0, 0] = 0.0  # exercises the clip
    for kw in ({}, {"scale": 0.7}, {"clip": None, "scale": 0.5}):
        a, b = utils.unary_from_softmax(sm, **kw), O.unary_from_softmax(sm, **kw)
        assert a.dtype == np.float32 and a.shape == (5, 63) and np.array_equal(a, b)
    assert utils.unary_from_softmax(sm)[0, 0] == np.float32(-np.log(1e-5))


def test_unary_from_labels_matches_restatement():
    rng = np.random.default_rng(1)
    lab = rng.integers(0, 4, (6, 5))
    for zu in (True, False):
        a, b = utils.unary_from_labels(lab, 4, 0.7, zero_unsure=zu), O.unary_from_labels(lab, 4, 0.7, zero_unsure=zu)
        assert a.dtype == np.float32 and np.array_equal(a, b)
    U = utils.unary_from_labels(lab, 4, 0.7, zero_unsure=False)
    p = lab.ravel()
    assert np.allclose(U[p, np.arange(p.size)], -np.log(0.7))
    assert np.allclose(np.delete(U[:, 0], p[0]), -np.log(0.3 / 3))


def test_create_pairwise_features_match_2d_kernels():
    H, W = 5, 7
    img = S.natural_image(H, W, 0)
    fg = utils.create_pairwise_gaussian((3, 2), (H, W))
    assert fg.shape == (2, H * W)
    ys, xs = np.mgrid[0:H, 0:W]
    assert np.allclose(fg[0], ys.ravel() / 3) and np.allclose(fg[1], xs.ravel() / 2)
    fb = utils.create_pairwise_bilateral((3, 2), (13, 13, 13), img, chdim=2)
    assert fb.shape == (5, H * W) and np.allclose(fb[2], img[..., 0].ravel() / 13)


def test_active_class_grouping_follows_dcrf_process():
    probs = S.blob_probs(6, 8, 8, seed=0, n_active=3)
    act = wsss._active_classes(probs)
    assert len(act) == 3 and (probs[act].sum((1, 2)) > 0).all()
    assert wsss._group_by([3, 2, 3, 0, 2]) == {3: [0, 2], 2: [1, 4], 0: [3]}


def test_unary_from_featmap_is_softmax_neglog():
    rng = np.random.default_rng(2)
    f = rng.standard_normal((4, 5, 3)).astype(np.float32)
    U = wsss._unary_from_featmap(f, use_log=True)
    assert U.shape == (3, 20) and U.flags.c_contiguous and U.dtype == np.float32
    p = np.exp(f) / np.exp(f).sum(2, keepdims=True)
    assert np.allclose(U, -np.log(p).reshape(20, 3).T, atol=1e-5)


def _reference_iou_irn(confusion):
    # literal restatement of 03b_irn/step/eval_sem_seg.py:43-50
    gtj = confusion.sum(axis=1)
    resj = confusion.sum(axis=0)
    gtjresj = np.diag(confusion)
    denominator = gtj + resj - gtjresj
    iou = gtjresj / denominator
    return iou, np.nanmean(iou)


def _reference_iou_sec(gt, pred, C_):
    # literal restatement of 03a_sec-dsrg/model.py:698-719,736 (VOC branch)
    intersect, union = np.zeros(C_), np.zeros(C_)
    for k in range(C_):
        gt_mask, pred_mask = gt == k, pred == k
        intersect[k] += np.sum(gt_mask & pred_mask)
        union[k] += np.sum(gt_mask | pred_mask)
    return intersect / (union + 1e-7), np.mean(intersect / (union + 1e-7))


def test_miou_conventions():
    rng = np.random.default_rng(3)
    C_ = 7
    gt = rng.integers(0, C_, 4000).astype(np.int32)
    gt[::13] = 255
    pred = rng.integers(0, C_ - 1, gt.size).astype(np.int32)  # class C-1 never predicted
    conf = O.confusion(gt, pred, C_)
    iou, miou = E.iou_irn(conf)
    with np.errstate(divide="ignore", invalid="ignore"):
        riou, rmiou = _reference_iou_irn(conf[:C_])
    assert np.array_equal(np.nan_to_num(iou, nan=-1), np.nan_to_num(riou, nan=-1)) and miou == rmiou
    iou2, miou2 = E.iou_sec(conf)
    riou2, rmiou2 = _reference_iou_sec(gt, pred, C_)
    assert np.array_equal(iou2, riou2) and miou2 == rmiou2


def test_shard_balanced_partitions_and_balances():
    rng = np.random.default_rng(3)
    counts = rng.choice([500 * 375, 375 * 500, 500 * 333, 500 * 500, 334 * 500, 2448 * 2448], size=203)
    for world in (1, 2, 3, 8):
        shards = [E.shard_balanced(counts, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(len(counts)))          # a partition
        assert all(s == sorted(s) for s in shards)
        loads = [int(sum(counts[i] for i in s)) for s in shards]
        assert max(loads) - min(loads) <= int(counts.max())                 # LPT bound
    assert E.shard_balanced([5, 5, 5, 5], 0, 2) == [0, 2] and E.shard_balanced([5, 5, 5, 5], 1, 2) == [1, 3]
    assert E.shard_balanced([], 0, 4) == []


def test_shard_indices_stride():
    assert E.shard_indices(10, 1, 4) == [1, 5, 9]
    allidx = sorted(sum((E.shard_indices(1449, r, 8) for r in range(8)), []))
    assert allidx == list(range(1449))


def test_synthetic_inputs_are_seeded():
    assert np.array_equal(S.natural_image(20, 30, 5), S.natural_image(20, 30, 5))
    assert not np.array_equal(S.natural_image(20, 30, 5), S.natural_image(20, 30, 6))
    U = S.random_unary(4, 50, 0)
    assert U.dtype == np.float32 and np.allclose(np.exp(-U).sum(0), 1, atol=1e-5)
    gt = S.gt_map(30, 40, 5, 0)
    assert gt.dtype == np.int32 and set(np.unique(gt)) <= set(range(5)) | {255}


def test_bench_algorithmic_bytes_formula():
    import bench

    N, L, Mg, Mb = 1000, 21, 130, 640
    per_iter = (bench.algorithmic_bytes("splat", 2, N, L, Mg) + bench.algorithmic_bytes("splat", 5, N, L, Mb)
                + 3 * bench.algorithmic_bytes("blur", 2, N, L, Mg) + 6 * bench.algorithmic_bytes("blur", 5, N, L, Mb)
                + bench.algorithmic_bytes("slice", None, N, L, None, [(2, Mg), (5, Mb)]))
    # SURVEY.md section 8d formula + one extra Q read (two splat launches) + the CSR row starts
    survey = 12 * L * N + sum(16 * (d + 1) * N + 8 * L * M + (d + 1) * (8 * L * M + 8 * M) for d, M in ((2, Mg), (5, Mb)))
    extra = 4 * L * N + 4 * (Mg + Mb)
    assert per_iter == survey + extra
    assert bench.iteration_bytes(N, L, [(2, Mg), (5, Mb)]) == survey


def test_bench_configs_cover_baseline_and_both_arms_share_the_config():
    """Every BASELINE.json configuration has a bench entry with the reference's CRF parameters
    (SURVEY.md Appendix B), and the CPU reference arm describes its workload with the same `config`
    object as the CUDA arm (the driver compares them)."""
    import bench

    C = bench.CONFIGS
    assert list(C)[0] == bench.HEADLINE == "voc32"
    assert {"voc32", "voc1", "sec41x32", "hsn321x16", "adp1088_morph", "adp1088_func", "dg612x8", "dg2448"} <= set(C)
    v = C["voc32"]
    assert (v["sizes"][0], v["L"], v["iters"], v["g_sxy"], v["g_compat"], v["b_sxy"], v["b_srgb"], v["b_compat"]) == \
        ((500, 375), 21, 10, 3.0, 3.0, 80.0, 13.0, 10.0)                      # 03a_sec-dsrg/SEC.py:20
    s_ = C["sec41x32"]
    assert (s_["sizes"][0], s_["iters"], s_["g_sxy"], s_["b_sxy"]) == ((41, 41), 5, 3.0 / 12, 80.0 / 12)   # SEC.py:19
    assert (C["adp1088_morph"]["L"], C["adp1088_func"]["L"], C["dg2448"]["L"]) == (29, 5, 6)
    assert (C["dg612x8"]["b_sxy"], C["dg612x8"]["b_srgb"]) == (50.0, 5.0)    # IRN crf_inference_label defaults
    assert C["dg2448"]["sizes"] == [(2448, 2448)]
    assert bench.SWEEP_IMAGES == 1449
    for name, cfg in C.items():
        d = bench.config_dict(name, cfg)
        assert d["name"] == name and d["images_per_gpu_per_step"] == len(cfg["sizes"]) and "impl" not in d
        assert bench.npix(cfg) * 6 < 2 ** 31                                 # one handle per step


def test_wrapper_batches_are_chunked_below_the_int32_entry_limit(monkeypatch):
    assert wsss._MAX_BATCH_PIXELS * 6 < 2 ** 31
    monkeypatch.setattr(wsss, "_MAX_BATCH_PIXELS", 25)
    assert wsss._chunks(range(5), [10] * 5) == [[0, 1], [2, 3], [4]]
    assert wsss._chunks([0, 2, 4], [30, 1, 2, 3, 40]) == [[0], [2], [4]]  # an over-size image runs alone
    assert wsss._chunks([], []) == []
    assert wsss._chunks(range(6), [1] * 6, max_pixels=4) == [[0, 1, 2, 3], [4, 5]]      # memory budget
    monkeypatch.setattr(wsss, "_MAX_BATCH_IMAGES", 2)                                     # launch-grid bound
    assert wsss._chunks(range(5), [1] * 5, max_pixels=100) == [[0, 1], [2, 3], [4]]
    assert wsss._bytes_per_pixel(21) > 21 * 4 * 4 and wsss._bytes_per_pixel(29) > wsss._bytes_per_pixel(21)


def test_wrapper_chunks_are_halved_when_the_device_runs_out_of_memory():
    """a chunk whose handle raises MemoryError (DCRF_ENOMEM) is split and retried; every
    index is still processed exactly once, in order; a single image that does not fit re-raises."""
    seen = []

    def fn(idx):
        if len(idx) > 2:
            raise MemoryError("pretend the lattices were denser than estimated")
        seen.append(list(idx))

    wsss._run_chunked(range(7), [1] * 7, 21, None, fn, budget=100)
    assert seen == [[0], [1, 2], [3, 4], [5, 6]] and sum(seen, []) == list(range(7))

    def always(idx):
        raise MemoryError("too large")

    try:
        wsss._run_chunked(range(2), [1, 1], 21, None, always, budget=100)
        raise AssertionError("expected MemoryError")
    except MemoryError:
        pass


def test_pipeline_sub_batch_views():
    """BatchPipeline._cut: images [lo, hi) of a flat concatenation or of a per-image list (the views
    handed to the sub-batch handles; no GPU involved)."""
    from wsss_analysis_b200.pipeline import BatchPipeline

    sizes = [(4, 3), (2, 5), (6, 1)]
    starts = np.concatenate([[0], np.cumsum([w * h for w, h in sizes])])
    L = 3
    flat = np.arange(int(starts[-1]) * L, dtype=np.float32)
    got = BatchPipeline._cut(flat, 1, 3, starts, L)
    assert got.base is flat or got.base is flat.base or np.shares_memory(got, flat)
    assert got[0] == 12 * L and got.size == (10 + 6) * L
    lst = ["a", "b", "c"]
    assert BatchPipeline._cut(lst, 0, 2, starts, L) == ["a", "b"]
    assert BatchPipeline._cut(None, 0, 2, starts, L) is None
    rgb = np.zeros(int(starts[-1]) * 3, np.uint8)
    assert BatchPipeline._cut(rgb, 2, 3, starts, 3).size == 6 * 3
