"""Integer confusion-matrix reduction behind mIoU, on the GPU, summed across GPUs with one NCCL
all-reduce.

Replaces
  * chainercv `calc_semantic_segmentation_confusion` + the IoU arithmetic of
    /root/reference/03b_irn/step/eval_sem_seg.py:41-50 (`iou = diag / (row + col - diag)`,
    `miou = nanmean(iou)`), and
  * the per-class intersect / union loops of /root/reference/03a_sec-dsrg/model.py:698-719,736 and
    /root/reference/03c_hsn/demo.py:185-191,234 (`mIoU = mean(I / (U + 1e-7))`, where the union
    also counts predicted-k pixels whose GT is an ignored / unlisted value).

Counts are int64 and integer addition is order independent, so the sharded result is bit-identical
to a single-process NumPy bincount.  PyTorch is used only to own the device buffer and for the
`torch.distributed` collective (NCCL on GPUs; gloo in the CPU tests, which exercise the host logic
with host-side counting disabled -- the counting itself always runs in the CUDA kernel).
"""
import ctypes as C

import numpy as np

from . import _lib


class CollectiveComm(object):
    """An NCCL communicator made through the C ABI (dcrf_nccl_comm_create): what the C-level
    collective `dcrf_confusion_allreduce` runs on.  The 128-byte ncclUniqueId is generated on rank 0 and
    handed to the other ranks either by the caller (`unique_id=`, any transport) or, by default,
    through `torch.distributed.broadcast_object_list` of an already initialised process group -- the
    only use of PyTorch here is that hand-off."""

    def __init__(self, rank, world_size, device, unique_id=None, group=None):
        self._lib = _lib.load()
        self.rank, self.world_size, self.device = int(rank), int(world_size), int(device)
        if unique_id is None:
            import torch.distributed as dist

            obj = [self.new_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(obj, src=0, group=group)
            unique_id = obj[0]
        assert len(unique_id) == 128
        comm = C.c_void_p()
        _lib.check(self._lib.dcrf_nccl_comm_create(self.world_size, self.rank, unique_id, self.device, C.byref(comm)))
        self._comm = comm

    @staticmethod
    def new_unique_id():
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().dcrf_nccl_unique_id(buf))
        return bytes(buf.raw)

    def all_reduce_i64(self, tensor, stream=None):
        """In-place SUM all-reduce of an int64 CUDA tensor on `stream` (default: torch's current stream)."""
        import torch

        assert tensor.is_cuda and tensor.dtype == torch.int64 and tensor.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream(tensor.device).cuda_stream
        _lib.check(self._lib.dcrf_confusion_allreduce(self._comm, tensor.data_ptr(), tensor.numel(),
                                                      tensor.device.index, stream))

    def close(self):
        comm, self._comm = self._comm, None
        if comm:
            _lib.check(self._lib.dcrf_nccl_comm_destroy(comm))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConfusionAccumulator(object):
    """(C+1, C) int64 device matrix: row = GT class, column = predicted class; row C collects pixels
    whose GT is outside [0, C) (VOC's 255, chainercv's -1) and is what the IRN convention ignores."""

    def __init__(self, n_classes, device=None):
        import torch

        self.C = int(n_classes)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.conf = torch.zeros((self.C + 1, self.C), dtype=torch.int64, device=self.device)
        self.bad = torch.zeros((1,), dtype=torch.int64, device=self.device)
        self._lib = _lib.load()

    def _dev_i32(self, x):
        import torch

        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32))
        if x.dtype != torch.int32:
            x = x.to(torch.int32)
        return x.to(self.device, non_blocking=False).contiguous().view(-1)

    def update(self, gt, pred):
        """gt, pred: same-size integer label maps (numpy or torch, host or device)."""
        import torch

        g, p = self._dev_i32(gt), self._dev_i32(pred)
        if g.numel() != p.numel():
            raise ValueError("gt and pred differ in size: %d vs %d" % (g.numel(), p.numel()))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dcrf_confusion_accumulate(g.data_ptr(), p.data_ptr(), g.numel(), self.C,
                                                       self.conf.data_ptr(), self.bad.data_ptr(),
                                                       self.device.index, stream))
        return self

    def synchronize(self):
        import torch

        torch.cuda.synchronize(self.device)

    def all_reduce(self, group=None, comm=None):
        """Sum over all ranks (NCCL over NVLink on GPUs).  comm: a CollectiveComm -> the C-level
        collective (dcrf_confusion_allreduce, ncclAllReduce int64 SUM); otherwise
        torch.distributed.all_reduce of the initialised process group (no-op without one)."""
        import torch.distributed as dist

        if comm is not None:
            comm.all_reduce_i64(self.conf)
            comm.all_reduce_i64(self.bad)
            return self
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.conf, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.bad, op=dist.ReduceOp.SUM, group=group)
        return self

    def result(self):
        """(C+1, C) int64 ndarray."""
        return self.conf.cpu().numpy()

    def bad_predictions(self):
        return int(self.bad.item())


def resize_nearest(labels, size_wh, device=None):
    """`cv2.resize(labels, (W, H), interpolation=cv2.INTER_NEAREST)` on the GPU
    (/root/reference/03b_irn/step/eval_sem_seg.py:36, 03c_hsn/demo.py:181-183).  labels: (h, w) integer
    map (numpy or torch); returns an int32 torch tensor (H, W) on the GPU."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    if not isinstance(labels, torch.Tensor):
        labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32))
    src = labels.to(dev, torch.int32).contiguous()
    W, H = int(size_wh[0]), int(size_wh[1])
    dst = torch.empty((H, W), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().dcrf_resize_nearest_i32(src.data_ptr(), src.shape[0], src.shape[1], dst.data_ptr(), H, W,
                                                    dev.index, torch.cuda.current_stream(dev).cuda_stream))
    return dst


def resize_bilinear(featmap, size_wh, device=None):
    """`cv2.resize(featmap, (W, H))` (INTER_LINEAR) of a float32 (h, w, C) map on the GPU
    (/root/reference/03a_sec-dsrg/model.py:686-687).  Returns a float32 torch tensor (H, W, C)."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    if not isinstance(featmap, torch.Tensor):
        featmap = torch.from_numpy(np.ascontiguousarray(featmap, dtype=np.float32))
    src = featmap.to(dev, torch.float32).contiguous()
    if src.dim() == 2:
        src = src.unsqueeze(-1)
    W, H = int(size_wh[0]), int(size_wh[1])
    dst = torch.empty((H, W, src.shape[2]), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().dcrf_resize_bilinear_f32(src.data_ptr(), src.shape[0], src.shape[1], src.shape[2],
                                                     dst.data_ptr(), H, W, dev.index,
                                                     torch.cuda.current_stream(dev).cuda_stream))
    return dst


def all_reduce_confusion_host(conf, group=None):
    """Host-side (gloo) form of the same collective for CPU-only tests of the sharding logic."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(conf, dtype=np.int64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


def iou_irn(conf):
    """IRN / chainercv convention (eval_sem_seg.py:43-50): ignored-GT row dropped,
    iou = diag / (gt_total + pred_total - diag), miou = nanmean."""
    c = np.asarray(conf)[:-1].astype(np.int64) if conf.shape[0] == conf.shape[1] + 1 else np.asarray(conf)
    gtj = c.sum(axis=1)
    resj = c.sum(axis=0)
    gtjresj = np.diag(c)
    denominator = gtj + resj - gtjresj
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = gtjresj / denominator
    return iou, float(np.nanmean(iou))


def iou_sec(conf):
    """03a / 03c convention (model.py:716-719,736): intersect[k] = #(gt==k & pred==k),
    union[k] = #(gt==k | pred==k) INCLUDING predicted-k pixels on ignored GT,
    mIoU = mean(I / (U + 1e-7))."""
    conf = np.asarray(conf)
    assert conf.shape[0] == conf.shape[1] + 1, "needs the (C+1, C) matrix with the ignored-GT row"
    C_ = conf.shape[1]
    inter = np.diag(conf[:C_]).astype(np.float64)
    gt_count = conf[:C_].sum(axis=1).astype(np.float64)
    pred_count = conf.sum(axis=0).astype(np.float64)  # includes the ignored-GT row
    union = gt_count + pred_count - inter
    iou = inter / (union + 1e-7)
    return iou, float(np.mean(iou))


def shard_indices(n_items, rank, world_size):
    """Item i -> rank i mod world_size: the `split_dataset` striding of
    /root/reference/03b_irn/step/cam_to_ir_label.py:114-117."""
    return list(range(rank, n_items, world_size))


def shard_balanced(pixel_counts, rank, world_size):
    """Length-balanced alternative for mixed image sizes: items are dealt in
    decreasing pixel count to the currently lightest rank (ties: lowest rank, lowest index), the
    same deterministic assignment on every rank; returns this rank's indices in ascending order.
    The confusion matrix is a sum over images, so any assignment gives the same all-reduced result."""
    counts = [int(c) for c in pixel_counts]
    order = sorted(range(len(counts)), key=lambda i: (-counts[i], i))
    load = [0] * int(world_size)
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += counts[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)
