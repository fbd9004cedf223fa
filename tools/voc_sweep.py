#!/usr/bin/env python
"""BASELINE config 5: VOC2012-val-shaped sweep sharded over the GPUs of one box.

    python tools/voc_sweep.py --images 1449                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/voc_sweep.py --images 1449                  # 8 GPUs

Rank 0 prints one JSON line: images/s (max time over ranks), both mIoU conventions and a SHA-256 of
the all-reduced int64 confusion matrix -- the digest must be identical for every GPU count."""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))


def main():
    import torch
    import torch.distributed as dist

    from wsss_analysis_b200 import sweep

    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1449)
    ap.add_argument("--labels", type=int, default=21)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    r = sweep.run_sweep(args.images, args.labels, rank, world, batch=args.batch, device=local)
    secs = torch.tensor([r["seconds"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    if rank == 0:
        conf = r["confusion"]
        print(json.dumps({
            "workload": "voc2012_val_shaped_sweep", "images": args.images, "n_gpus": world,
            "seconds_max_over_ranks": float(secs.item()), "images_per_s": args.images / float(secs.item()),
            "note": "time = CRF (host buffers in, labels stay on the GPU) + confusion; synthetic input generation excluded",
            "miou_irn": r["miou_irn"], "miou_sec": r["miou_sec"], "pixels_counted": int(conf.sum()),
            "confusion_sha256": hashlib.sha256(conf.tobytes()).hexdigest()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
