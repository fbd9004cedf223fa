"""GPU: the CUDA path against the committed golden fixtures, and the reference call-site wrappers
(wsss.py) against literal restatements of the reference functions driven by the CPU oracle."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("resize_"))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_matches_golden(path):
    from wsss_analysis_b200 import densecrf as G

    z = np.load(path)
    W, H, L, n, gs, gc, bs, srgb, bc = z["params"]
    W, H, L, n = int(W), int(H), int(L), int(n)
    d = G.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(z["U"])
    d.addPairwiseGaussian(sxy=gs, compat=gc)
    d.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=z["img"], compat=bc)
    for k, pre in ((0, "g_"), (1, "b_")):
        e = d.lattice_export(k)
        assert e["M"] == int(z[pre + "M"])
        assert np.array_equal(e["keys"], z[pre + "keys"])
        assert np.array_equal(e["offsets"], z[pre + "offsets"])
        assert np.array_equal(e["bary"].view(np.uint32), z[pre + "bary"].view(np.uint32))
        assert np.array_equal(e["neighbours"], z[pre + "neigh"])
    Q = d.inference(n)
    assert np.abs(Q - z["Q"]).max() <= 1e-4
    assert (Q.argmax(0) == z["labels"]).mean() >= 0.999


def _ref_dcrf_process(probs, images, config):
    """03c_hsn/utilities.py:399-445 with the oracle standing in for pydensecrf."""
    from oracle import oracle as O

    gauss_sxy, gauss_compat, bilat_sxy, bilat_srgb, bilat_compat, n_infer = config
    n, C_ = probs.shape[0], probs.shape[1]
    size = images.shape[1:3]
    crf = np.zeros((n, C_, size[0], size[1]))
    for i in range(n):
        pass_class_inds = np.where(np.sum(np.sum(probs[i], axis=1), axis=1) > 0)
        d = O.DenseCRF2D(size[1], size[0], len(pass_class_inds[0]))
        if len(pass_class_inds[0]) > 0:
            U = np.ascontiguousarray(O.unary_from_softmax(probs[i, pass_class_inds[0]]))
            d.setUnaryEnergy(U)
            d.addPairwiseGaussian(sxy=gauss_sxy, compat=gauss_compat)
            d.addPairwiseBilateral(sxy=bilat_sxy, srgb=bilat_srgb, rgbim=np.uint8(images[i]), compat=bilat_compat)
            Q = d.inference(n_infer)
            crf[i, pass_class_inds] = np.array(Q).reshape((len(pass_class_inds[0]), size[0], size[1]))
    return np.argmax(crf, axis=1), crf


def test_dcrf_process_drop_in():
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    B, C_, H, W = 5, 8, 48, 56
    n_active = [3, 5, 3, 0, 8]
    probs = np.stack([S.blob_probs(C_, H, W, seed=i, n_active=n_active[i]) if n_active[i] else np.zeros((C_, H, W))
                      for i in range(B)])
    images = np.stack([S.histo_image(H, W, i) for i in range(B)]).astype(np.float64)
    config = np.array([3 / 2, 3, 80 / 2, 13, 10, 10.0])  # 03c_hsn/demo.py:159 (n_infer arrives as a float)
    out = wsss.dcrf_process(probs, images, config)
    ref, ref_crf = _ref_dcrf_process(probs, images, config)
    assert out.shape == (B, H, W) and out.dtype == np.int64
    # identical labels wherever the oracle's top-2 margin is not at float-noise level
    srt = np.sort(ref_crf, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-4
    assert (out == ref)[decided].all()
    assert (out == ref).mean() >= 0.999


def test_crf_inference_and_sec_layer():
    from oracle import oracle as O
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    cfg = {"g_sxy": 3 / 12, "g_compat": 3, "bi_sxy": 80 / 12, "bi_srgb": 13, "bi_compat": 10, "iterations": 5}
    B, h, w, C_ = 4, 41, 41, 21
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((B, h, w, C_)).astype(np.float32) * 2
    image = np.stack([S.natural_image(h, w, i) for i in range(B)]).astype(np.float32)
    got = wsss.sec_crf_layer(feat, image, cfg, C_, min_prob=1e-4)
    assert got.shape == (B, h, w, C_) and got.dtype == np.float32
    # literal SEC.py:270-280 with the oracle
    ret = np.zeros(feat.shape, np.float32)
    for i in range(B):
        f = np.exp(feat[i] - feat[i].max(2, keepdims=True))
        f /= f.sum(2, keepdims=True)
        U = np.copy(np.swapaxes((-np.log(f)).reshape(-1, C_), 0, 1), order="C")
        d = O.DenseCRF2D(w, h, C_)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
        d.addPairwiseBilateral(sxy=cfg["bi_sxy"], srgb=cfg["bi_srgb"], rgbim=image[i].astype(np.uint8), compat=cfg["bi_compat"])
        ret[i] = np.transpose(d.inference(5).reshape(C_, h, w), (1, 2, 0))
    one = wsss.crf_inference(image[0].astype(np.uint8), cfg, C_, feat[0], use_log=True)
    assert np.abs(one - ret[0]).max() <= 1e-4
    ret[ret < 1e-4] = 1e-4
    ret /= ret.sum(3, keepdims=True)
    ref = np.log(ret)
    assert np.abs(got - ref).max() <= 2e-3  # log of values clamped at 1e-4: 1e-4 abs on Q -> <= ~1e-3 here
    assert np.abs(np.exp(got) - np.exp(ref)).max() <= 1e-4


def test_marginals_hwc_and_sec_epilogue_match_numpy():
    """dcrf_get_q_hwc: the (H, W, C) layout is a pure re-arrangement of inference()'s output (bit for
    bit); with min_prob the clamp + renormalisation equal the NumPy lines of SEC.py:277-278 bit for bit
    (the kernel follows NumPy's float32 summation order), and the log agrees within float rounding."""
    import torch

    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S

    for L in (21, 5, 8, 2, 37):
        sizes = [(41, 41), (23, 17), (1, 9)]
        imgs = [S.natural_image(h, w, 3 + i) for i, (w, h) in enumerate(sizes)]
        Us = [S.random_unary(L, w * h, 7 + i) * 3 for i, (w, h) in enumerate(sizes)]   # peaky: many Q < 1e-4
        d = G.DenseCRFBatch(sizes, L)
        d.setUnaryEnergy(Us)
        d.addPairwiseGaussian(sxy=3 / 12, compat=3)
        d.addPairwiseBilateral(sxy=80 / 12, srgb=13, rgbim=imgs, compat=10)
        d.run(5)
        Q = d.marginals()
        hwc = d.marginals_hwc()
        clamped = d.marginals_hwc(min_prob=1e-4)
        logged = d.marginals_hwc(min_prob=1e-4, log=True)
        dev = d.marginals_hwc_device(min_prob=1e-4, log=True)
        torch.cuda.synchronize()
        assert np.array_equal(dev.cpu().numpy(), np.concatenate([x.ravel() for x in logged]))
        n_clamped = 0
        for q, a, c, lg, (w, h) in zip(Q, hwc, clamped, logged, sizes):
            ref = np.ascontiguousarray(np.transpose(q.reshape(L, h, w), (1, 2, 0)))
            assert a.shape == (h, w, L) and np.array_equal(a, ref)
            ret = ref.copy()[None]
            n_clamped += int((ret < 1e-4).sum())
            ret[ret < 1e-4] = 1e-4
            ret /= np.sum(ret, axis=3, keepdims=True)
            assert np.array_equal(c, ret[0])
            np.testing.assert_allclose(lg, np.log(ret[0]), rtol=0, atol=2e-6)
        if L >= 5:
            assert n_clamped > 0
        d.close()


def test_crf_inference_label_drop_in():
    from oracle import oracle as O
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    H, W, nl = 60, 80, 4
    img = S.natural_image(H, W, 7).astype(np.float32)  # IRN loaders hand float32 0-255 images
    labels = (S.gt_map(H, W, nl, 7, ignore=0, border=3)).astype(np.int64)
    got = wsss.crf_inference_label(img, labels, "voc12", n_labels=nl)
    d = O.DenseCRF2D(W, H, nl)
    d.setUnaryEnergy(O.unary_from_labels(labels, nl, gt_prob=0.7, zero_unsure=False))
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=50, srgb=5, rgbim=np.ascontiguousarray(img.astype(np.uint8)), compat=10)
    Q = d.inference(10).reshape(nl, H, W)
    ref = np.argmax(Q, axis=0)
    assert got.shape == (H, W)
    srt = np.sort(Q, axis=0)
    decided = (srt[-1] - srt[-2]) > 1e-4
    assert (got == ref)[decided].all() and (got == ref).mean() >= 0.999


def test_sharded_sweep_confusion_matches_oracle_pipeline():
    """Config-5-style sweep on one GPU, emulating 2 ranks by running both shards and summing: the
    int64 confusion must equal the single-shard run bit for bit, and equal a NumPy bincount over
    the oracle's label maps wherever the oracle's argmax is decided."""
    from oracle import oracle as O
    from wsss_analysis_b200 import evaluation as E
    from wsss_analysis_b200 import sweep

    C_, n = 6, 5

    def item(i):
        img, U, gt = sweep.synthetic_item(i, C_)
        return img[:90, :120].copy(), np.ascontiguousarray(U.reshape(C_, *gt.shape)[:, :90, :120].reshape(C_, -1)), gt[:90, :120].copy()

    items = [item(i) for i in range(n)]
    whole = sweep.run_sweep(items, C_, 0, 1, batch=3, all_reduce=False)
    parts = [sweep.run_sweep(items, C_, r, 2, batch=2, all_reduce=False) for r in range(2)]
    assert np.array_equal(whole["confusion"], parts[0]["confusion"] + parts[1]["confusion"])
    assert whole["images"] == n and parts[0]["images"] + parts[1]["images"] == n
    # oracle pipeline
    ref = np.zeros((C_ + 1, C_), np.int64)
    undecided = 0
    for img, U, gt in items:
        h, w = gt.shape
        d = O.DenseCRF2D(w, h, C_)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
        Q = d.inference(10)
        srt = np.sort(Q, axis=0)
        undecided += int(((srt[-1] - srt[-2]) <= 1e-4).sum())
        ref += O.confusion(gt, Q.argmax(0), C_)
    assert np.abs(whole["confusion"] - ref).sum() <= 2 * undecided
    assert abs(whole["miou_irn"] - E.iou_irn(ref)[1]) < 1e-3


def test_batch_pipeline_matches_blocking_calls():
    """pipeline.BatchPipeline (async-host handles on dedicated streams) returns the same bits as the
    blocking DenseCRFBatch calls, for marginals and for label maps."""
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200.pipeline import BatchPipeline, pinned_empty

    cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 40, "bi_srgb": 13, "bi_compat": 10, "iterations": 4}
    L = 5
    batches, ref = [], []
    for k in range(5):
        sizes = [(48 + 4 * k, 36), (40, 30 + k)]
        imgs = [S.natural_image(h, w, 10 * k + i) for i, (w, h) in enumerate(sizes)]
        Us = [S.random_unary(L, w * h, 10 * k + i) for i, (w, h) in enumerate(sizes)]
        n = sum(w * h for w, h in sizes)
        U = pinned_empty(n * L)
        U[:] = np.concatenate([u.ravel() for u in Us])
        I = pinned_empty(n * 3, np.uint8)
        I[:] = np.concatenate([im.ravel() for im in imgs])
        want_labels = k % 2 == 1
        out = pinned_empty(n, np.int32) if want_labels else pinned_empty(n * L)
        batches.append(dict(sizes=sizes, n_labels=L, unary=U, rgb=I, cfg=cfg, out=out, labels=want_labels))
        d = G.DenseCRFBatch(sizes, L)
        d.setUnaryEnergy(Us)
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
        ref.append(d.map(4) if want_labels else d.inference(4))
    with BatchPipeline(n_slots=2) as pipe:
        got = pipe.map(batches)
    for g_, r_ in zip(got, ref):
        assert len(g_) == len(r_)
        for a, b in zip(g_, r_):
            assert a.shape == b.shape and np.array_equal(a, b)


def test_batch_pipeline_chunked_sub_batches():
    """chunk_images cuts a batch into sub-batches (own handles, uploads overlapping kernels); every
    image is an independent CRF, so the results equal the un-chunked call bit for bit."""
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200.pipeline import BatchPipeline, pinned_empty

    cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 40, "bi_srgb": 13, "bi_compat": 10, "iterations": 3}
    L = 4
    sizes = [(40 + 3 * i, 30 + (i % 3)) for i in range(7)]
    imgs = [S.natural_image(h, w, i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, 100 + i) for i, (w, h) in enumerate(sizes)]
    n = sum(w * h for w, h in sizes)
    U = pinned_empty(n * L)
    U[:] = np.concatenate([u.ravel() for u in Us])
    I = pinned_empty(n * 3, np.uint8)
    I[:] = np.concatenate([im.ravel() for im in imgs])
    d = G.DenseCRFBatch(sizes, L)
    d.setUnaryEnergy(Us)
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=40, srgb=13, rgbim=imgs, compat=10)
    ref_q = d.inference(3)
    ref_l = [q.argmax(0).astype(np.int32) for q in ref_q]
    for chunk in (3, 2, 7, 100):
        outq, outl = pinned_empty(n * L), pinned_empty(n, np.int32)
        with BatchPipeline(n_slots=3, chunk_images=chunk) as pipe:
            t1 = pipe.submit(sizes, L, U, I, cfg, out=outq)
            t2 = pipe.submit(sizes, L, Us, imgs, cfg, out=outl, labels=True)   # per-image lists
            got_q, got_l = pipe.result(t1), pipe.result(t2)
        assert len(got_q) == len(got_l) == len(sizes)
        for a, b in zip(got_q, ref_q):
            assert a.shape == b.shape and np.array_equal(a, b)
        for a, b, (w, h) in zip(got_l, ref_l, sizes):
            assert a.shape == (h, w) and np.array_equal(a.ravel(), b)
        # the flat output buffer holds the images back to back exactly like the un-chunked call
        assert np.array_equal(outq, np.concatenate([q.ravel() for q in ref_q]))


def test_gpu_unary_construction_matches_numpy():
    """dcrf_set_unary_from_{probs,logits,labels} against pydensecrf.utils-style NumPy unaries: the
    unary is read back through inference(0) = softmax(-U)."""
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import utils, wsss

    sizes, L = [(37, 23), (16, 40)], 6
    rng = np.random.default_rng(0)

    def q0(setter):
        d = G.DenseCRFBatch(sizes, L)
        setter(d)
        return d.inference(0)

    def ref_q0(Us):
        return q0(lambda d: d.setUnaryEnergy(Us))

    probs = [S.blob_probs(L, h, w, seed=i) for i, (w, h) in enumerate(sizes)]
    probs[0][2, :3, :3] = 0.0  # exercises the clip
    for kw in ({}, {"scale": 0.6}, {"clip": None, "scale": 0.9}):
        got = q0(lambda d: d.setUnaryFromSoftmax(probs, **kw))
        want = ref_q0([utils.unary_from_softmax(p, **kw) for p in probs])
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-9)
    p32 = [p.astype(np.float32) for p in probs]
    for a, b in zip(q0(lambda d: d.setUnaryFromSoftmax(p32)), ref_q0([utils.unary_from_softmax(p) for p in p32])):
        np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-9)

    feats = [rng.standard_normal((h, w, L)).astype(np.float32) * 3 for (w, h) in sizes]
    for use_log in (True,):
        got = q0(lambda d: d.setUnaryFromLogits(feats, use_log))
        want = ref_q0([wsss._unary_from_featmap(f, use_log) for f in feats])
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=5e-6, atol=1e-9)

    labels = [rng.integers(0, L, (h, w)) for (w, h) in sizes]
    for zu in (False, True):
        got = q0(lambda d: d.setUnaryFromLabels(labels, 0.7, zero_unsure=zu))
        want = ref_q0([utils.unary_from_labels(lab, L, 0.7, zero_unsure=zu) for lab in labels])
        for a, b in zip(got, want):
            np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-9)
    with pytest.raises(ValueError, match="label out of range"):
        q0(lambda d: d.setUnaryFromLabels([lab + L for lab in labels], 0.7, zero_unsure=False))


def test_crf_inference_label_shares_lattices_between_label_sets():
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    imgs = [S.natural_image(40, 52, i).astype(np.float32) for i in range(3)]
    fg = [S.gt_map(40, 52, 4, i, ignore=0, border=2) for i in range(3)]
    bg = [S.gt_map(40, 52, 4, 10 + i, ignore=0, border=2) for i in range(3)]
    both = wsss.crf_inference_label_batch(imgs, fg, n_labels=4, extra_labels=[bg])
    sep = [wsss.crf_inference_label_batch(imgs, fg, n_labels=4), wsss.crf_inference_label_batch(imgs, bg, n_labels=4)]
    for k in range(2):
        for a, b in zip(both[k], sep[k]):
            assert np.array_equal(a, b)


def test_wrappers_accept_cuda_tensors_and_move_no_payload_over_pcie():
    """dcrf_process, crf_inference_batch, sec_crf_layer and
    crf_inference_label_batch given torch CUDA tensors return CUDA tensors with the same values as
    the host-array calls, and the library's own H2D / D2H byte counters (dcrf_copy_count) only see
    batch geometry and vertex counts -- no unaries, images, marginals or label maps."""
    import torch

    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    dev = torch.device("cuda", 0)
    B, C_, H, W = 3, 7, 40, 56
    rng = np.random.default_rng(5)
    images = np.stack([S.natural_image(H, W, 20 + i) for i in range(B)])
    probs = np.stack([S.blob_probs(C_, H, W, 30 + i, n_active=4) for i in range(B)])
    feat = rng.standard_normal((B, H, W, C_)).astype(np.float32) * 2
    labels = rng.integers(0, C_, (B, H, W)).astype(np.int64)
    cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 80, "bi_srgb": 13, "bi_compat": 10, "iterations": 5}
    hsn = (1.5, 3, 40, 13, 10, np.float64(5.0))
    # host-array results first (these do use PCIe)
    want_proc = wsss.dcrf_process(probs, images, hsn)
    want_inf = wsss.crf_inference_batch(list(images), cfg, C_, list(feat))
    want_sec = wsss.sec_crf_layer(feat, images.astype(np.float32), cfg, C_)
    want_lab = wsss.crf_inference_label_batch(list(images.astype(np.float32)), list(labels), n_labels=C_)
    t_probs, t_img = torch.from_numpy(probs).to(dev), torch.from_numpy(images).to(dev)
    t_feat, t_lab = torch.from_numpy(feat).to(dev), torch.from_numpy(labels).to(dev)
    t_imgf = t_img.to(torch.float32)
    torch.cuda.synchronize()
    h0, d0 = G.copy_count()
    got_proc = wsss.dcrf_process(t_probs, t_img, hsn)
    got_inf = wsss.crf_inference_batch(t_img, cfg, C_, t_feat)
    got_sec = wsss.sec_crf_layer(t_feat, t_imgf, cfg, C_)
    got_lab = wsss.crf_inference_label_batch(t_imgf, t_lab, n_labels=C_)
    torch.cuda.synchronize()
    h1, d1 = G.copy_count()
    payload = min(probs[0].nbytes, images[0].nbytes)
    assert h1 - h0 < 4096 and d1 - d0 < 4096 and payload > 4096, (h1 - h0, d1 - d0)
    for t in (got_proc, got_inf, got_sec, got_lab):
        assert t.is_cuda
    assert np.array_equal(got_proc.cpu().numpy(), want_proc)
    assert np.array_equal(got_inf.cpu().numpy(), np.stack(want_inf))
    assert np.array_equal(got_sec.cpu().numpy(), want_sec)
    assert np.array_equal(got_lab.cpu().numpy(), np.stack(want_lab))


def test_uint8_label_outputs_match_int32():
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S

    sizes = [(50, 30), (31, 44)]
    L = 21
    imgs = [S.natural_image(h, w, i) for i, (w, h) in enumerate(sizes)]
    Us = [S.random_unary(L, w * h, i) for i, (w, h) in enumerate(sizes)]
    d = G.DenseCRFBatch(sizes, L)
    d.setUnaryEnergy(Us)
    d.addPairwiseGaussian(sxy=3, compat=3)
    d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=imgs, compat=10)
    a, b = d.map(4), d.map(4, dtype=np.uint8)
    for x, y in zip(a, b):
        assert y.dtype == np.uint8 and x.dtype == np.int32 and np.array_equal(x, y)
    c = d.labels(dtype=np.uint8)
    assert all(np.array_equal(x, y) for x, y in zip(a, c))
    import torch

    t = d.map_device(4, dtype=torch.uint8)
    assert t.dtype == torch.uint8 and np.array_equal(t.cpu().numpy(), np.concatenate([x.ravel() for x in a]))


def test_handle_outlives_its_creating_thread():
    """a handle created in a worker thread and used / destroyed after that thread
    has exited must keep working -- it holds a reference on the thread's stream set."""
    import threading

    from oracle import oracle as O
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S

    W, H, L = 48, 32, 5
    img, U = S.natural_image(H, W, 1), S.random_unary(L, W * H, 1)
    box = {}

    def worker():
        g = G.DenseCRF2D(W, H, L)
        g.setUnaryEnergy(U)
        g.addPairwiseGaussian(sxy=3, compat=3)
        g.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
        box["g"] = g

    for _ in range(3):   # several generations of worker threads: stream handles get recycled
        th = threading.Thread(target=worker)
        th.start()
        th.join()
        o = O.DenseCRF2D(W, H, L)
        o.setUnaryEnergy(U)
        o.addPairwiseGaussian(sxy=3, compat=3)
        o.addPairwiseBilateral(sxy=80, srgb=13, rgbim=img, compat=10)
        assert np.abs(box["g"].inference(5) - o.inference(5)).max() <= 1e-4   # side streams created after the thread died
        box.pop("g").close()


def test_wrappers_take_an_arithmetic_mode():
    """the drop-in wrappers can select the arithmetic -- per call (`arithmetic=`) or for
    every call site at once (`wsss.ARITHMETIC`)."""
    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import synthetic as S
    from wsss_analysis_b200 import wsss

    cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 80, "bi_srgb": 13, "bi_compat": 10, "iterations": 5}
    H, W, C_ = 40, 52, 21
    rng = np.random.default_rng(3)
    feat = rng.standard_normal((H, W, C_)).astype(np.float32) * 2
    img = S.natural_image(H, W, 3)
    direct = {}
    for mode in ("fma", "strict"):
        d = G.DenseCRFBatch([(W, H)], C_)
        d.set_arithmetic(mode)
        d.setUnaryFromLogits([feat])
        d.addPairwiseGaussian(sxy=3, compat=3)
        d.addPairwiseBilateral(sxy=80, srgb=13, rgbim=[img], compat=10)
        d.run(5)
        direct[mode] = d.marginals_hwc()[0]
        d.close()
    assert not np.array_equal(direct["fma"], direct["strict"])
    assert np.array_equal(wsss.crf_inference(img, cfg, C_, feat, arithmetic="strict"), direct["strict"])
    assert np.array_equal(wsss.crf_inference(img, cfg, C_, feat), direct["fma"])         # auto -> fma at srgb = 13
    wsss.ARITHMETIC = "strict"
    try:
        assert np.array_equal(wsss.crf_inference(img, cfg, C_, feat), direct["strict"])
    finally:
        wsss.ARITHMETIC = None
