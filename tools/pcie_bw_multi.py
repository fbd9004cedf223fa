"""Raw pinned-memory PCIe bandwidth with all ranks copying at the same time (torchrun)."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 504_000_000 // 4
h1 = torch.empty(n, dtype=torch.float32).pin_memory(); h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d1 = torch.empty(n, dtype=torch.float32, device="cuda"); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bar():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
res = {}
for name in ("H2D", "D2H", "BOTH"):
    bar(); t = time.perf_counter()
    for _ in range(8):
        if name in ("H2D", "BOTH"):
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if name in ("D2H", "BOTH"):
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 8
    res[name] = n * 4 * (2 if name == "BOTH" else 1) / dt / 1e9
t = torch.tensor([res["H2D"], res["D2H"], res["BOTH"]], device="cuda", dtype=torch.float64)
if world > 1:
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(t, op=dist.ReduceOp.SUM)
else:
    lo = t
if rank == 0:
    print("ranks %d: aggregate GB/s H2D %.0f D2H %.0f BOTH %.0f | slowest rank H2D %.1f D2H %.1f BOTH %.1f" % (world, *t.tolist(), *lo.tolist()))
if world > 1: dist.destroy_process_group()
