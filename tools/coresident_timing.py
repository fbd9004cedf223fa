"""Experiment: two half-batch handles on two streams with capped-residency iteration kernels
(DCRF_CORESIDENT = resident threads per SM per kernel) against one full batch.  Device-resident
inputs/outputs, lattice build included.  Usage: DCRF_CORESIDENT=1024 python tools/coresident_timing.py [B] [slots]"""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200.pipeline import BatchPipeline
dev = torch.device("cuda", 0)
imgs, unaries = bench.make_inputs(32)
cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 80, "bi_srgb": 13, "bi_compat": 10, "iterations": 10}


def run(B, n_slots, cores):
    os.environ["DCRF_CORESIDENT"] = str(cores)
    def dev_batch(lo, hi):
        U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries[lo:hi]])).to(dev)
        I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs[lo:hi]])).to(dev)
        return U, I
    parts = [dev_batch(i, i + B) for i in range(0, 32, B)]
    sizes = [(bench.W_IMG, bench.H_IMG)] * B
    Q = [torch.empty(B * bench.L_LAB * bench.W_IMG * bench.H_IMG, dtype=torch.float32, device=dev)
         for _ in range(max(n_slots, len(parts)))]
    pipe = BatchPipeline(n_slots=n_slots, device=0)
    best = 1e9
    for rep in range(4):
        rounds = 5
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tk = []
        for r in range(rounds):
            for i, (U, I) in enumerate(parts):
                tk.append(pipe.submit(sizes, bench.L_LAB, U, I, cfg, out=Q[i % len(Q)]))
        for t in tk: pipe.result(t)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / rounds
        best = min(best, dt)
    print("coresident=%-5s B=%-2d slots=%d: %.2f ms per 32 images -> %.0f Mpix*iter/s" % (
        cores, B, n_slots, best * 1e3, 32 * bench.W_IMG * bench.H_IMG * 10 / best / 1e6), flush=True)
    pipe.close()
    del Q, parts
    torch.cuda.empty_cache()


for B, n_slots, cores in [(32, 1, 0), (16, 2, 0), (16, 2, 1024), (16, 2, 768), (16, 2, 1280), (16, 2, 512),
                          (8, 4, 1024), (8, 4, 512), (16, 3, 768), (32, 2, 1024), (32, 1, 1024), (32, 1, 2048)]:
    run(B, n_slots, cores)
