"""Diagnostic: device-resident batches through BatchPipeline (n slots) vs back to back."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from wsss_analysis_b200.pipeline import BatchPipeline
B = 32
dev = torch.device("cuda", 0)
imgs, unaries = bench.make_inputs(B)
sizes = [(bench.W_IMG, bench.H_IMG)] * B
U = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).to(dev)
I = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).to(dev)
cfg = {"g_sxy": 3, "g_compat": 3, "bi_sxy": 80, "bi_srgb": 13, "bi_compat": 10, "iterations": 10}
for n_slots in (1, 2, 3):
    Q = [torch.empty(B * bench.L_LAB * bench.W_IMG * bench.H_IMG, dtype=torch.float32, device=dev) for _ in range(n_slots)]
    torch.cuda.synchronize()
    pipe = BatchPipeline(n_slots=n_slots, device=0)
    for rep in range(2):
        steps = 9
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tk = [pipe.submit(sizes, bench.L_LAB, U, I, cfg, out=Q[i % n_slots]) for i in range(steps)]
        for t in tk: pipe.result(t)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    print("slots %d: %.2f ms/step -> %.0f Mpix*iter/s" % (n_slots, dt * 1e3, B * bench.W_IMG * bench.H_IMG * 10 / dt / 1e6))
    pipe.close()
