"""Sharded evaluation sweep: CRF every image of a list, accumulate the integer confusion matrix on the
GPU, sum it over ranks with ONE all-reduce, derive mIoU.

Mirrors the reference's only parallel use of the hot path -- images striped over workers with no
exchange (/root/reference/03b_irn/step/cam_to_ir_label.py:114-117, `split_dataset` + `mp.spawn`) --
followed by the evaluation of /root/reference/03b_irn/step/eval_sem_seg.py:28-50 (confusion over
the whole list, `iou = diag / (row + col - diag)`, `nanmean`) and of
/root/reference/03a_sec-dsrg/model.py:698-736.  One process per GPU; the only collective is the
int64 all-reduce of the (C+1, C) matrix (NCCL over NVLink on GPUs).

BASELINE.json config 5: 1449 VOC2012-val-shaped images (the length of
/root/reference/03b_irn/voc12/val.txt), 21 labels, 10 iterations.
"""
import time

import numpy as np

from . import synthetic
from .densecrf import DenseCRFBatch
from .evaluation import ConfusionAccumulator, iou_irn, iou_sec, shard_balanced, shard_indices

VOC_CRF = dict(g_sxy=3, g_compat=3, bi_sxy=80, bi_srgb=13, bi_compat=10, iterations=10)  # SEC.py:20


def synthetic_item(i, n_labels, seed=0):
    """Image i of the synthetic sweep: (image uint8 HxWx3, unary (L, N) f32, gt int32 HxW with 255 ignore)."""
    w, h = synthetic.voc_like_size(i, seed)
    img = synthetic.natural_image(h, w, seed * 100003 + i)
    gt = synthetic.gt_map(h, w, n_labels, seed * 100003 + i)
    # unary = noisy evidence for the GT label (ignored pixels get label 0 evidence)
    rng = np.random.default_rng(seed * 7 + i)
    z = rng.standard_normal((n_labels, h * w)).astype(np.float32) * 1.5
    lab = np.where(gt.ravel() == 255, 0, gt.ravel())
    z[lab, np.arange(h * w)] += 2.0
    z -= z.max(axis=0, keepdims=True)
    U = -(z - np.log(np.exp(z).sum(axis=0, keepdims=True)))
    return img, np.ascontiguousarray(U.astype(np.float32)), gt


def run_sweep(items, n_labels, rank=0, world_size=1, batch=16, crf=VOC_CRF, device=None, all_reduce=True,
              item_fn=None, pixel_counts=None):
    """items: number of images (synthetic) or a list of (img, unary, gt) tuples.
    Returns dict(confusion (C+1, C) int64, miou_irn, miou_sec, images (this rank), seconds)."""
    n_items = items if isinstance(items, int) else len(items)
    get = (item_fn or (lambda i: synthetic_item(i, n_labels))) if isinstance(items, int) else (lambda i: items[i])
    # striding (the reference's split_dataset convention) unless per-item pixel counts are given,
    # in which case the shards are balanced by pixel count
    mine = (shard_indices(n_items, rank, world_size) if pixel_counts is None
            else shard_balanced(pixel_counts, rank, world_size))
    acc = ConfusionAccumulator(n_labels, device=device)
    seconds = 0.0  # CRF + confusion only; fetching / synthesising the inputs is not part of the path
    for b0 in range(0, len(mine), batch):
        chunk = [get(i) for i in mine[b0:b0 + batch]]
        t0 = time.perf_counter()
        sizes = [(im.shape[1], im.shape[0]) for im, _, _ in chunk]
        d = DenseCRFBatch(sizes, n_labels, device=device)
        d.setUnaryEnergy([u for _, u, _ in chunk])
        d.addPairwiseGaussian(sxy=crf["g_sxy"], compat=crf["g_compat"])
        d.addPairwiseBilateral(sxy=crf["bi_sxy"], srgb=crf["bi_srgb"], rgbim=[im for im, _, _ in chunk],
                               compat=crf["bi_compat"])
        labels = d.map_device(crf["iterations"])          # int32, concatenated, stays on the GPU
        d.close()
        gt = np.concatenate([g.ravel() for _, _, g in chunk]).astype(np.int32)
        acc.update(gt, labels)
        acc.synchronize()
        seconds += time.perf_counter() - t0
    if all_reduce:
        acc.all_reduce()
    conf = acc.result()
    return dict(confusion=conf, miou_irn=iou_irn(conf)[1], miou_sec=iou_sec(conf)[1], images=len(mine),
                seconds=seconds, bad_predictions=acc.bad_predictions())


def run_sweep_device(n_items, n_labels, rank=0, world_size=1, batch=32, crf=VOC_CRF, device=0, all_reduce=True,
                     verify=True, seed=0, comm=None):
    """The same sweep with inputs synthesised on the GPU (`synthetic.torch_sweep_item`) and handed over
    as device tensors: unaries, images, label maps and the confusion matrix never touch the host.
    Image i goes to rank i mod world_size (`split_dataset` striding, cam_to_ir_label.py:114-117).
    verify: this rank's (C+1, C) matrix is also computed with np.bincount from the downloaded label
    maps and must match bit for bit (raises otherwise).
    Returns dict(confusion (all-reduced), miou_irn, miou_sec, images, pixels, seconds, verified)."""
    import torch

    dev = torch.device("cuda", int(device))
    mine = shard_indices(n_items, rank, world_size)
    acc = ConfusionAccumulator(n_labels, device=device)
    ref = np.zeros((n_labels + 1, n_labels), np.int64)
    seconds, pixels = 0.0, 0
    with torch.cuda.device(dev):
        for b0 in range(0, len(mine), batch):
            items = [synthetic.torch_sweep_item(i, n_labels, dev, seed) for i in mine[b0:b0 + batch]]
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            sizes = [(int(im.shape[1]), int(im.shape[0])) for im, _, _ in items]
            d = DenseCRFBatch(sizes, n_labels, device=device)
            d.setUnaryEnergy([u for _, u, _ in items])
            d.addPairwiseGaussian(sxy=crf["g_sxy"], compat=crf["g_compat"])
            d.addPairwiseBilateral(sxy=crf["bi_sxy"], srgb=crf["bi_srgb"], rgbim=[im for im, _, _ in items],
                                   compat=crf["bi_compat"])
            labels = d.map_device(crf["iterations"])          # int32, concatenated, stays on the GPU
            d.close()
            gt = torch.cat([g.reshape(-1) for _, _, g in items])
            acc.update(gt, labels)
            acc.synchronize()
            seconds += time.perf_counter() - t0
            pixels += int(gt.numel())
            if verify:
                g_h, p_h = gt.cpu().numpy().astype(np.int64), labels.cpu().numpy().astype(np.int64)
                row = np.where((g_h >= 0) & (g_h < n_labels), g_h, n_labels)
                ref += np.bincount(row * n_labels + p_h, minlength=(n_labels + 1) * n_labels).reshape(n_labels + 1, n_labels)
    if verify and not np.array_equal(acc.result(), ref):
        raise AssertionError("GPU confusion matrix differs from np.bincount on rank %d" % rank)
    if all_reduce:
        acc.all_reduce(comm=comm)   # comm: evaluation.CollectiveComm -> dcrf_confusion_allreduce (C ABI)
    conf = acc.result()
    return dict(confusion=conf, miou_irn=iou_irn(conf)[1], miou_sec=iou_sec(conf)[1], images=len(mine),
                pixels=pixels, seconds=seconds, verified=bool(verify), bad_predictions=acc.bad_predictions())
