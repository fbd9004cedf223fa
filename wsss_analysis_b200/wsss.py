"""The reference's CRF call-site helpers, re-hosted on the batched B200 engine.

Each function keeps the name, argument meaning and return layout of the helper it replaces so the
wsss-analysis call sites switch over with an import change only:

  dcrf_process          /root/reference/03c_hsn/utilities.py:399-445          (in tree)
  crf_inference         03a_sec-dsrg `lib/crf.py` -- file missing from the tree; signature from its
                        call sites SEC.py:275, DSRG.py:328, model.py:689,693
  sec_crf_layer         the `crf` py_func closure of SEC.py:270-280 / DSRG.py:323-332
  crf_inference_label   03b_irn `misc/imutils.py` -- file missing; call sites
                        step/cam_to_ir_label.py:35,47,52,67

Where the reference loops over images serially, these run every image of the call in ONE batched
handle (all kernels cover the whole batch).  Host-side NumPy here is only shape / dtype glue and the
unary construction the reference also does in NumPy; the CRF itself has no CPU path.
"""
import numpy as np

from .densecrf import DenseCRFBatch, _is_torch
from .utils import unary_from_labels, unary_from_softmax  # noqa: F401  (host forms, kept for callers)

__all__ = ["dcrf_process", "crf_inference", "crf_inference_batch", "sec_crf_layer", "crf_inference_label",
           "crf_inference_label_batch", "IRN_CRF_CONFIG"]

# [EXT] defaults of jiwoon-ahn/irn `crf_inference_label` (last row)
IRN_CRF_CONFIG = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 50, "bi_srgb": 5, "bi_compat": 10, "iterations": 10}


# arithmetic of the handles the wrappers create: None = the library default ("auto", or whatever the
# DCRF_ARITHMETIC environment variable says); "fma" / "reference" / "strict" force a mode for every call
# site at once (the drop-in wrappers had no way to select the exact arithmetic)
ARITHMETIC = None


def _new_batch(sizes, n_labels, device, arithmetic=None):
    d = DenseCRFBatch(sizes, n_labels, device=device)
    mode = ARITHMETIC if arithmetic is None else arithmetic
    if mode is not None:
        d.set_arithmetic(mode)
    return d


def _on_gpu(x):
    """torch CUDA tensor?  Such inputs stay on the GPU: unaries, images, marginals and label maps are
    handed to / returned by the library as device pointers and the result is a
    CUDA tensor; only a few hundred bytes of batch geometry cross PCIe."""
    return _is_torch(x) and x.is_cuda


def _active_classes(probs_img):
    """`np.where(np.sum(np.sum(probs[i], axis=1), axis=1) > 0)` of utilities.py:425."""
    return np.where(np.sum(np.sum(probs_img, axis=1), axis=1) > 0)[0]


# a handle indexes lattice entries with int32: N * (d + 1) < 2^31 with d = 5; stay well below
_MAX_BATCH_PIXELS = (1 << 31) // 6 // 2
# grid.y of the per-image kernels
_MAX_BATCH_IMAGES = 65535


def _bytes_per_pixel(n_labels):
    """Device memory one pixel of a Gaussian + bilateral model needs while its handle is alive, for
    lattices of up to 4 vertices per pixel in total (VOC: 0.8, DeepGlobe with srgb = 5: 4.1): unary, Q,
    the two value ping-pong buffers per lattice, the entry / CSR tables and the build temporaries."""
    lp = (int(n_labels) + 3) // 4 * 4
    return 40 * lp + 830


def _pixel_budget(n_labels, device=None):
    """Pixels per handle: the int32 entry bound, or what 60 % of the free device memory holds."""
    import ctypes as C

    from . import _lib

    free = C.c_int64(0)
    _lib.check(_lib.load().dcrf_mem_info(-1 if device is None else int(device), C.byref(free), None))
    return max(1, min(_MAX_BATCH_PIXELS, int(0.6 * free.value) // _bytes_per_pixel(n_labels)))


def _chunks(indices, npix, max_pixels=None):
    """Split `indices` into consecutive runs whose total pixel count fits one handle (and whose
    image count fits a launch grid).  An image larger than the budget runs alone."""
    budget = _MAX_BATCH_PIXELS if max_pixels is None else max_pixels
    out, cur, tot = [], [], 0
    for i in indices:
        if cur and (tot + npix[i] > budget or len(cur) >= _MAX_BATCH_IMAGES):
            out.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += npix[i]
    if cur:
        out.append(cur)
    return out


def _run_chunked(indices, npix, n_labels, device, fn, budget=None):
    """fn(idx) for every memory-sized chunk of `indices`; a chunk that still runs out of device memory
    (lattices denser than the estimate) is halved and retried."""
    from .densecrf import trim_memory

    work = _chunks(indices, npix, _pixel_budget(n_labels, device) if budget is None else budget)
    while work:
        idx = work.pop(0)
        try:
            fn(idx)
        except MemoryError:
            if len(idx) == 1:
                raise
            try:
                trim_memory()
            except RuntimeError:
                pass
            work = [idx[:len(idx) // 2], idx[len(idx) // 2:]] + work


def _group_by(keys):
    """indices grouped by key, groups in order of first appearance, indices ascending."""
    groups = {}
    for i, k in enumerate(keys):
        groups.setdefault(k, []).append(i)
    return groups


def dcrf_process(probs, images, config, device=None, arithmetic=None):
    """Drop-in for `dcrf_process(probs, images, config)` (03c_hsn/utilities.py:399-445).

    probs  (B, C, H, W) class probabilities; images (B, H, W, 3) any dtype (cast with np.uint8 like
    the reference); config = (gauss_sxy, gauss_compat, bilat_sxy, bilat_srgb, bilat_compat, n_infer)
    with n_infer possibly float-valued.  Returns (B, H, W) int64 argmax over the C classes of the
    per-image CRF marginals scattered back to their class slots (inactive classes stay 0).

    Images are grouped by their number of active classes and each group runs as one batch."""
    gauss_sxy, gauss_compat, bilat_sxy, bilat_srgb, bilat_compat, n_infer = config
    if _on_gpu(probs):
        return _dcrf_process_device(probs, images, config, arithmetic)
    probs = np.asarray(probs)
    num_input_images, num_classes = probs.shape[0], probs.shape[1]
    size = images.shape[1:3]
    H, W = int(size[0]), int(size[1])
    # The reference scatters Q back into a zero (B, C, H, W) float64 array and takes np.argmax over C
    # (utilities.py:421,443-445).  Marginals of active classes are > 0 and sum to 1, inactive slots
    # are 0 and active class indices ascend, so that argmax equals active[argmax over the active
    # classes] (first maximum wins in both): only the int32 label map leaves the GPU.
    out = np.zeros((num_input_images, H, W), dtype=np.int64)
    active = [_active_classes(probs[i]) for i in range(num_input_images)]
    for n_act, members in _group_by([len(a) for a in active]).items():
        if n_act == 0:
            continue  # the reference builds DenseCRF2D(w, h, 0) and leaves crf[i] = 0 -> label 0

        def run(idx, n_act=n_act):
            d = _new_batch([(W, H)] * len(idx), n_act, device, arithmetic)
            try:
                d.setUnaryFromSoftmax([probs[i, active[i]] for i in idx])  # clip + -log on the GPU (utilities.py:431)
                d.addPairwiseGaussian(sxy=gauss_sxy, compat=gauss_compat)
                d.addPairwiseBilateral(sxy=bilat_sxy, srgb=bilat_srgb, rgbim=[np.uint8(images[i]) for i in idx],
                                       compat=bilat_compat)
                labels = d.map(n_infer, dtype=np.uint8 if n_act <= 256 else np.int32)
            finally:
                d.close()
            for j, i in enumerate(idx):
                out[i] = active[i][labels[j]]

        _run_chunked(members, [H * W] * num_input_images, n_act, device, run)
    return out


def _dcrf_process_device(probs, images, config, arithmetic=None):
    """dcrf_process for CUDA tensors: probs (B, C, H, W) float64 / float32, images (B, H, W, 3) any
    dtype; returns a (B, H, W) int64 CUDA tensor.  Only the (B, C) table of active classes is read on
    the host (it decides how the images are grouped into batches)."""
    import torch

    gauss_sxy, gauss_compat, bilat_sxy, bilat_srgb, bilat_compat, n_infer = config
    B, C_, H, W = (int(v) for v in probs.shape)
    dev = probs.device
    if probs.dtype not in (torch.float32, torch.float64):
        probs = probs.to(torch.float64)
    img8 = images.to(device=dev, dtype=torch.uint8).contiguous()          # np.uint8(images[i]) of utilities.py:439
    act_mask = (probs.sum(dim=(2, 3)) > 0).cpu().numpy()                   # utilities.py:425
    active = [np.flatnonzero(act_mask[i]) for i in range(B)]
    out = torch.zeros((B, H, W), dtype=torch.int64, device=dev)
    groups = []
    for n_act, members in _group_by([len(a) for a in active]).items():
        groups += [(n_act, c) for c in _chunks(members, [H * W] * B, _pixel_budget(n_act, dev.index))]
    for n_act, idx in groups:
        if n_act == 0:
            continue
        act_t = [torch.as_tensor(active[i], device=dev) for i in idx]
        d = _new_batch([(W, H)] * len(idx), n_act, dev.index, arithmetic)
        d.setUnaryFromSoftmax(torch.cat([probs[i].index_select(0, a).reshape(-1) for i, a in zip(idx, act_t)]))
        d.addPairwiseGaussian(sxy=gauss_sxy, compat=gauss_compat)
        d.addPairwiseBilateral(sxy=bilat_sxy, srgb=bilat_srgb, rgbim=img8[idx].contiguous(), compat=bilat_compat)
        labels = d.map_device(n_infer).view(len(idx), H, W).long()
        d.close()
        for j, (i, a) in enumerate(zip(idx, act_t)):
            out[i] = a[labels[j]]
    return out


def _unary_from_featmap(feat, use_log=True):
    """[EXT] unary of SEC's crf_inference(use_log=True): softmax over the class axis then -log,
    returned as C-contiguous (C, H*W) float32.  (Host form kept for callers / tests; the batched path
    computes it on the GPU.)  use_log=False is exercised by no call site of the reference and the
    wrapper's source is not in its tree, so that branch is not guessed."""
    if not use_log:
        raise NotImplementedError("crf_inference(use_log=False) is not exercised by the reference")
    feat = np.asarray(feat, dtype=np.float32)
    C_ = feat.shape[-1]
    feat = np.exp(feat - np.max(feat, axis=2, keepdims=True))
    feat /= np.sum(feat, axis=2, keepdims=True)
    unary = -np.log(feat)
    unary = np.reshape(unary, (-1, C_))
    unary = np.swapaxes(unary, 0, 1)
    return np.copy(unary, order="C").astype(np.float32, copy=False)


def crf_inference_batch(imgs, crf_config, num_classes, featmaps, use_log=True, device=None, min_prob=None,
                        log=False, arithmetic=None):
    """Batched `crf_inference`: imgs list of (H_b, W_b, 3) uint8, featmaps list of (H_b, W_b, C).
    Returns a list of (H_b, W_b, C) float32 marginals (written in that layout by the GPU).
    `min_prob` / `log`: the clamp + renormalise + log epilogue of the SEC / DSRG `crf` closure."""
    if _on_gpu(featmaps):
        # (B, H, W, C) float32 feature maps + (B, H, W, 3) images as CUDA tensors -> (B, H, W, C) CUDA tensor
        import torch

        B, H, W, C_ = (int(v) for v in featmaps.shape)
        dev = featmaps.device
        img8 = imgs.to(device=dev, dtype=torch.uint8).contiguous()
        res = torch.empty((B, H, W, C_), dtype=torch.float32, device=dev)
        for idx in _chunks(range(B), [H * W] * B, _pixel_budget(num_classes, dev.index)):
            d = _new_batch([(W, H)] * len(idx), num_classes, dev.index, arithmetic)
            d.setUnaryFromLogits(featmaps[idx[0]:idx[-1] + 1].to(torch.float32).contiguous(), use_log)
            d.addPairwiseGaussian(sxy=crf_config["g_sxy"], compat=crf_config["g_compat"])
            d.addPairwiseBilateral(sxy=crf_config["bi_sxy"], srgb=crf_config["bi_srgb"],
                                   rgbim=img8[idx[0]:idx[-1] + 1], compat=crf_config["bi_compat"])
            d.run(crf_config["iterations"])
            d.marginals_hwc_device(out=res[idx[0]:idx[-1] + 1].view(-1), min_prob=min_prob, log=log)
            d.close()
        return res
    all_sizes = [(int(im.shape[1]), int(im.shape[0])) for im in imgs]
    out = [None] * len(imgs)

    def run(idx):
        d = _new_batch([all_sizes[i] for i in idx], num_classes, device, arithmetic)
        try:
            d.setUnaryFromLogits([np.asarray(featmaps[i], dtype=np.float32) for i in idx], use_log)
            d.addPairwiseGaussian(sxy=crf_config["g_sxy"], compat=crf_config["g_compat"])
            d.addPairwiseBilateral(sxy=crf_config["bi_sxy"], srgb=crf_config["bi_srgb"],
                                   rgbim=[np.ascontiguousarray(imgs[i], dtype=np.uint8) for i in idx],
                                   compat=crf_config["bi_compat"])
            d.run(crf_config["iterations"])
            Q = d.marginals_hwc(min_prob=min_prob, log=log)
        finally:
            d.close()
        for i, q in zip(idx, Q):
            out[i] = q

    _run_chunked(range(len(imgs)), [w * h for w, h in all_sizes], num_classes, device, run)
    return out


def crf_inference(img, crf_config, num_classes, featmap, use_log=True, device=None, arithmetic=None):
    """Drop-in for SEC/DSRG's `crf_inference(img, crf_config, num_classes, featmap, use_log=True)`
    (call sites 03a_sec-dsrg/SEC.py:275, model.py:689-693): (H, W, 3) uint8 image + (H, W, C) feature
    map -> (H, W, C) float32 marginals."""
    return crf_inference_batch([img], crf_config, num_classes, [featmap], use_log, device, arithmetic=arithmetic)[0]


def sec_crf_layer(featemap, image, crf_config, num_classes, min_prob=1e-4, device=None, arithmetic=None):
    """The `crf` closure run through tf.py_func in SEC.py:270-280 / DSRG.py:323-332, whole batch in
    one handle: featemap (B, h, w, C) float32, image (B, h, w, 3) float -> uint8;
    returns log of the clamped (>= min_prob), renormalised marginals, (B, h, w, C) float32.
    CUDA tensors in -> CUDA tensor out (the training hook of SURVEY.md 8f rank 2: nothing but batch
    geometry crosses PCIe)."""
    if _on_gpu(featemap):
        return crf_inference_batch(image, crf_config, num_classes, featemap, use_log=True, min_prob=min_prob, log=True,
                                   arithmetic=arithmetic)
    featemap = np.asarray(featemap)
    batch_size = featemap.shape[0]
    image = np.asarray(image).astype(np.uint8)
    # clamp (>= min_prob), renormalisation over classes and log run on the GPU (dcrf_get_q_hwc); the
    # clamp, the sum (NumPy's float32 order) and the quotient are bit-identical to the NumPy lines
    out = crf_inference_batch([image[i] for i in range(batch_size)], crf_config, num_classes,
                              [featemap[i] for i in range(batch_size)], use_log=True, device=device,
                              min_prob=min_prob, log=True, arithmetic=arithmetic)
    ret = np.zeros(featemap.shape, dtype=np.float32)
    for i in range(batch_size):
        ret[i, :, :, :] = out[i]
    return ret


def crf_inference_label_batch(imgs, labels, n_labels=21, t=10, gt_prob=0.7, crf_config=None, device=None,
                              extra_labels=(), arithmetic=None):
    """Batched `crf_inference_label`: returns a list of (H_b, W_b) int label maps.

    `extra_labels`: further label sets for the SAME images (each a list like `labels`).  The
    lattices depend on the image only, so they are built once and every label set just replaces
    the unary -- the VOC branch of cam_to_ir_label.py runs the CRF twice per image (fg threshold
    :47, bg threshold :52).  With extra sets the return value is a list of result lists."""
    cfg = dict(IRN_CRF_CONFIG if crf_config is None else crf_config)
    if _on_gpu(imgs):
        # (B, H, W, 3) images and (B, H, W) label maps as CUDA tensors -> (B, H, W) int64 CUDA tensor(s)
        import torch

        B, H, W = (int(v) for v in imgs.shape[:3])
        dev = imgs.device
        img8 = imgs.to(torch.uint8).contiguous()
        sets = [labels] + list(extra_labels)
        res = [torch.empty((B, H, W), dtype=torch.int64, device=dev) for _ in sets]
        for idx in _chunks(range(B), [H * W] * B, _pixel_budget(n_labels, dev.index)):
            lo, hi = idx[0], idx[-1] + 1
            d = _new_batch([(W, H)] * len(idx), n_labels, dev.index, arithmetic)
            d.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
            d.addPairwiseBilateral(sxy=cfg["bi_sxy"], srgb=cfg["bi_srgb"], rgbim=img8[lo:hi], compat=cfg["bi_compat"])
            for k, ls in enumerate(sets):
                d.setUnaryFromLabels(ls[lo:hi].to(device=dev, dtype=torch.int32).contiguous(), gt_prob=gt_prob,
                                     zero_unsure=False)
                res[k][lo:hi] = d.map_device(t).view(len(idx), H, W)
            d.close()
        return res[0] if not extra_labels else res
    all_sizes = [(int(im.shape[1]), int(im.shape[0])) for im in imgs]
    label_sets = [labels] + list(extra_labels)
    outs = [[None] * len(imgs) for _ in label_sets]

    def run(idx):
        d = _new_batch([all_sizes[i] for i in idx], n_labels, device, arithmetic)
        try:
            d.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
            # the IRN loaders hand over float32 0-255 HWC images (voc12/dataloader.py:93,102-103)
            d.addPairwiseBilateral(sxy=cfg["bi_sxy"], srgb=cfg["bi_srgb"],
                                   rgbim=[np.ascontiguousarray(np.asarray(imgs[i]).astype(np.uint8)) for i in idx],
                                   compat=cfg["bi_compat"])
            for k, ls in enumerate(label_sets):
                d.setUnaryFromLabels([np.asarray(ls[i]) for i in idx], gt_prob=gt_prob, zero_unsure=False)
                for i, o in zip(idx, d.map(t, dtype=np.uint8 if n_labels <= 256 else np.int32)):
                    outs[k][i] = o.astype(np.int64)
        finally:
            d.close()

    _run_chunked(range(len(imgs)), [w * h for w, h in all_sizes], n_labels, device, run)
    return outs[0] if not extra_labels else outs


def crf_inference_label(img, labels, dataset=None, t=10, n_labels=21, gt_prob=0.7, device=None, arithmetic=None):
    """Drop-in for `imutils.crf_inference_label(img, labels, dataset, n_labels=...)`
    (03b_irn/step/cam_to_ir_label.py:35,47,52,67).  `dataset` is this fork's extra positional
    argument; its effect in the missing wrapper is unknown and it is ignored here."""
    del dataset
    return crf_inference_label_batch([img], [labels], n_labels=n_labels, t=t, gt_prob=gt_prob, device=device,
                                     arithmetic=arithmetic)[0]
