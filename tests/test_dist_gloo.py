"""CPU, world_size 2 over gloo: images shard over ranks by striding, per-rank integer confusion
matrices are summed with one all-reduce, and the result is bit-identical to a single process
(SURVEY.md section 8e).  On GPUs the same code path runs over NCCL (ConfusionAccumulator.all_reduce);
here the per-rank counting uses the oracle as a stand-in for the CUDA kernel."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _make(n_img, C_):
    rng = np.random.default_rng(0)
    gts, preds = [], []
    for i in range(n_img):
        n = int(rng.integers(200, 400))
        g = rng.integers(0, C_, n).astype(np.int32)
        g[:: 11 + i] = 255
        gts.append(g)
        preds.append(rng.integers(0, C_, n).astype(np.int32))
    return gts, preds


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import oracle as O
    from wsss_analysis_b200 import evaluation as E

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C_ = 5
    gts, preds = _make(13, C_)
    conf = np.zeros((C_ + 1, C_), np.int64)
    for i in E.shard_indices(len(gts), rank, world):
        conf += O.confusion(gts[i], preds[i], C_)
    total = E.all_reduce_confusion_host(conf)
    # same sweep with the shards balanced by pixel count instead of strided
    conf_b = np.zeros((C_ + 1, C_), np.int64)
    for i in E.shard_balanced([g.size for g in gts], rank, world):
        conf_b += O.confusion(gts[i], preds[i], C_)
    total_b = E.all_reduce_confusion_host(conf_b)
    assert np.array_equal(total, total_b)
    q.put((rank, total.tolist()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_confusion_allreduce_is_bit_exact():
    import torch.multiprocessing as mp

    from oracle import oracle as O
    from wsss_analysis_b200 import evaluation as E

    C_ = 5
    gts, preds = _make(13, C_)
    single = O.confusion(np.concatenate(gts), np.concatenate(preds), C_)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for _, total in results:
        assert np.array_equal(np.array(total, np.int64), single)
    assert E.iou_irn(single)[1] == E.iou_irn(np.array(results[0][1], np.int64))[1]
