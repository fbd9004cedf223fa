#!/usr/bin/env python
"""bench.py -- DenseCRF mean-field throughput on B200 (metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic VOC2012-shaped images
(500x375 RGB, 21 labels, 10 mean-field iterations, Gaussian sxy=3 compat=3, bilateral sxy=80
srgb=13 compat=10 -- /root/reference/03a_sec-dsrg/SEC.py:20): for every image the two permutohedral
lattices are built (a new image means a new lattice) and 10 iterations are run.
`value`  : Mpix*iter/s with unaries and images already resident in HBM (device pointers in, device out)
`e2e`    : the same batch through the same C-ABI calls with HOST (pinned) buffers: H2D of unaries and
           images and D2H of the marginals Q are inside the timed region.
One process per GPU; images shard over ranks with no data-path collective ("weak" scaling: the
per-GPU batch is fixed); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W_IMG, H_IMG, L_LAB, N_ITER = 500, 375, 21, 10
G_SXY, G_COMPAT, B_SXY, B_SRGB, B_COMPAT = 3.0, 3.0, 80.0, 13.0, 10.0
METRIC, UNIT = "densecrf_mpix_iter_per_s", "Mpix*iter/s"
N_DISTINCT = 4  # distinct synthetic images generated on the host, tiled up to the batch size


def config_dict(batch, impl):
    return {
        "workload": "voc2012_shaped_batch: %d images/GPU/step of %dx%d RGB, %d labels, %d iterations, "
                    "gaussian sxy=%g compat=%g + bilateral sxy=%g srgb=%g compat=%g, lattice build included"
                    % (batch, W_IMG, H_IMG, L_LAB, N_ITER, G_SXY, G_COMPAT, B_SXY, B_SRGB, B_COMPAT),
        "images_per_gpu_per_step": batch,
        "image": "natural-like synthetic (smooth colour field + N(0,8) noise), seeds 0..%d tiled" % (N_DISTINCT - 1),
        "l2_policy": "inputs larger than L2 (unaries alone are %.0f MB per step per GPU)"
                     % (batch * L_LAB * W_IMG * H_IMG * 4 / 1e6),
        "impl": impl,
    }


def make_inputs(batch):
    from wsss_analysis_b200 import synthetic as S

    imgs = [S.natural_image(H_IMG, W_IMG, s) for s in range(N_DISTINCT)]
    unaries = [S.random_unary(L_LAB, W_IMG * H_IMG, s) for s in range(N_DISTINCT)]
    return [imgs[i % N_DISTINCT] for i in range(batch)], [unaries[i % N_DISTINCT] for i in range(batch)]


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement of pydensecrf (pydensecrf itself cannot be installed
# here: SURVEY.md section 0.2), one image per thread on all host cores.
# ------------------------------------------------------------------------------------------------
def cpu_one_image(args):
    from oracle import oracle as O

    img, U = args
    d = O.DenseCRF2D(W_IMG, H_IMG, L_LAB)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=G_SXY, compat=G_COMPAT)
    d.addPairwiseBilateral(sxy=B_SXY, srgb=B_SRGB, rgbim=img, compat=B_COMPAT)
    return d.inference(N_ITER)


# measured LSU row-gather ceiling (tools/micro/bulk_gather.cu, profiles/r1_micro_gather.txt)
GATHER_CEILING_GROWS = 106.7


def cpu_sample(n_images, threads):
    """Wall time of `n_images` VOC-shaped CRFs over `threads` host threads (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    O.lib()
    imgs, unaries = make_inputs(min(n_images, N_DISTINCT))
    work = [(imgs[i % len(imgs)], unaries[i % len(unaries)]) for i in range(n_images)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(cpu_one_image, work))
    return time.perf_counter() - t0


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n_images = threads  # one image per thread per step: a bounded sample of the batch workload
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_sample(min(n_images, threads), threads)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_sample(n_images, threads)   # CRF time only: generating the synthetic inputs is not the path
    value = args.steps * n_images * W_IMG * H_IMG * N_ITER / dt / 1e6
    sample = "%d VOC-shaped images per step (one per host thread), %d steps" % (n_images, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(n_images, "oracle port of pydensecrf on host cores"),
        "images_per_s": args.steps * n_images / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", os.environ.get("BENCH_CLOCK_MS", "200")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi's start-up stalls driver calls of other processes for ~100 ms: wait for
            # its first sample so that the stall is outside the timed region
            t0 = time.time()
            while not self.lines and time.time() - t0 < 10.0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(cls, d, N, L, M, n_terms_info=None):
    """Algorithmic bytes of ONE launch (DESIGN.md section 'Roofline accounting'); float32 values,
    int32 ids, L real labels (padding lanes are not counted)."""
    # (the pre- / post-normalisation vectors are folded into the packed entry weights at build time,
    # so no iteration kernel reads them and they are not counted)
    if cls == "splat":   # Q read + (pixel id, weight) per entry + row starts + lattice write
        return 4 * L * N + 8 * (d + 1) * N + 4 * M + 4 * L * M
    if cls == "blur":    # lattice read + write + two neighbour ids
        return 8 * L * M + 8 * M
    if cls == "slice":   # unary read + Q write + per term: (vertex id, weight) per entry + lattice read
        b = 8 * L * N
        for (dk, Mk) in n_terms_info:
            b += 8 * (dk + 1) * N + 4 * L * Mk
        return b
    raise KeyError(cls)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from wsss_analysis_b200 import densecrf as G

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own banner ("NCCL version ...") goes to stdout by default; stdout carries the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    N = W_IMG * H_IMG
    imgs, unaries = make_inputs(B)
    sizes = [(W_IMG, H_IMG)] * B
    # host (pinned) and device copies of the step's inputs / outputs
    U_host = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).pin_memory()
    I_host = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).pin_memory()
    Q_host = torch.empty(B * L_LAB * N, dtype=torch.float32).pin_memory()
    U_dev, I_dev = U_host.to(dev), I_host.to(dev)
    Q_dev = torch.empty(B * L_LAB * N, dtype=torch.float32, device=dev)
    # a real (non-default) stream: the library launches on it and the CUDA events below bracket it
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    prof = {"splat": {}, "blur": {}, "slice": {}}
    lattice_M = {}

    # e2e leg: the same batch through the host-buffer API, driven by wsss_analysis_b200.pipeline
    # (H2D | build + iterations | D2H as three overlapped stages over n_slots host threads); every
    # batch pays its own H2D of unaries + image and its own D2H of Q inside the timed region.
    from wsss_analysis_b200.pipeline import BatchPipeline

    n_slots = int(os.environ.get("BENCH_SLOTS", "3"))
    slot_Q = [Q_host] + [torch.empty_like(Q_host).pin_memory() for _ in range(n_slots - 1)]
    crf_cfg = {"g_sxy": G_SXY, "g_compat": G_COMPAT, "bi_sxy": B_SXY, "bi_srgb": B_SRGB, "bi_compat": B_COMPAT,
               "iterations": N_ITER}
    # every 32-image step goes through the pipeline as two 16-image sub-batches (upload of one
    # overlaps the kernels of the other inside the step; measured 28.8 vs 29.7 ms per step)
    chunk = int(os.environ.get("BENCH_CHUNK", "16")) or None
    pipe = BatchPipeline(n_slots=n_slots, device=local, chunk_images=chunk)

    def step(device_resident, profile=False):
        crf = G.DenseCRFBatch(sizes, L_LAB, device=local, stream=stream)
        if profile:
            crf.profile_enable(True)
        crf.setUnaryEnergy(U_dev)
        crf.addPairwiseGaussian(sxy=G_SXY, compat=G_COMPAT)
        crf.addPairwiseBilateral(sxy=B_SXY, srgb=B_SRGB, rgbim=I_dev, compat=B_COMPAT)
        crf.inference_device(N_ITER, out=Q_dev)
        if profile:
            for k in range(2):
                d, M, _ = crf.lattice_info(k)
                lattice_M[d] = M
            for cls, cid in (("splat", 0), ("blur", 1), ("slice", 2)):
                for tag in ((2, 5) if cls != "slice" else (2,)):
                    ms, n = crf.profile_read(cid, tag)
                    a = prof[cls].setdefault(tag, [0.0, 0])
                    a[0] += ms
                    a[1] += n
        crf.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(device_resident, steps, profile):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = G.launch_count()
        e0.record(stream)
        for _ in range(steps):
            t_step = time.perf_counter()
            step(device_resident, profile)
            if os.environ.get("BENCH_VERBOSE"):
                print("step %.2f ms" % ((time.perf_counter() - t_step) * 1e3), file=sys.stderr)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = G.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def timed_e2e(steps):
        """`steps` batches through the host-buffer pipeline (n_slots batches in flight)."""
        barrier()
        t0 = time.perf_counter()
        tickets = [pipe.submit(sizes, L_LAB, U_host.numpy(), I_host.numpy(), crf_cfg, out=slot_Q[i % n_slots].numpy())
                   for i in range(steps)]
        for t_ in tickets:
            pipe.result(t_)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3   # results are in host memory: wall clock is the e2e clock
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # each leg is warmed right before it is timed (the stream-ordered memory pool re-balances when
    # the allocating stream changes, which would otherwise land in the first timed step)
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(True)
    ms_dev, launches = timed(True, args.steps, False)
    # per-kernel pass for the roofline: same steps with the library's CUDA-event pairs around every
    # launch; profiling serialises the two pairwise filters (they overlap on two streams otherwise)
    ms_prof, _ = timed(True, args.steps, True)
    timed_e2e(2 * n_slots)  # warm-up of the slot threads (their memory pools, pinned buffers)
    ms_e2e = timed_e2e(args.steps)
    pipe.close()
    clocks = sampler.stop() if rank == 0 else None

    total_pix_iter = world * B * N * N_ITER * args.steps
    value = total_pix_iter / (ms_dev * 1e-3) / 1e6
    e2e_value = total_pix_iter / (ms_e2e * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel class (largest share of device time in the timed region)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    Ntot = B * N
    kernels = []
    for cls in ("splat", "blur", "slice"):
        for tag, (ms, n) in prof[cls].items():
            if n == 0:
                continue
            if cls == "slice":
                info = [(d, lattice_M[d]) for d in sorted(lattice_M)]
                by = algorithmic_bytes("slice", None, Ntot, L_LAB, None, info)
                name = "slice_softmax_kernel (fused %d terms)" % tag
            else:
                by = algorithmic_bytes(cls, tag, Ntot, L_LAB, lattice_M[tag])
                name = "%s_kernel d=%d" % (cls, tag)
            # row gathers per launch (one 96-byte row per lattice entry): the splat gathers E = N(d+1)
            # pixel rows, the fused slice E_gauss + E_bilat vertex rows; compared below with the
            # measured LSU gather ceiling of tools/micro/bulk_gather.cu
            rows = None
            if cls == "slice":
                rows = sum(Ntot * (d + 1) for d in lattice_M)
            elif cls == "splat":
                rows = Ntot * (tag + 1)
            kernels.append({"kernel": name, "launches": n, "total_ms": ms, "avg_us": ms / n * 1e3,
                            "algorithmic_bytes_per_launch": by, "achieved_gbs": by / (ms / n * 1e-3) / 1e9,
                            "gather_rows_per_launch": rows})
    kernels.sort(key=lambda k: -k["total_ms"])
    top = kernels[0]
    kernel_ms = sum(k["total_ms"] for k in kernels)
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture (same batch size only)
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("batch") == B:
            traffic = tj["kernels"].get(top["kernel"])
    roofline = {
        "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": top["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
        "share_of_step": top["total_ms"] / ms_prof,
        "timing": "CUDA events recorded by the library on the launching stream around every launch, in a separate "
                  "pass of the same %d steps run right after the timed region (%.2f ms/step profiled and "
                  "serialised vs %.2f ms/step timed)" % (args.steps, ms_prof / args.steps, ms_dev / args.steps),
        "per_kernel": [{"kernel": k["kernel"], "launches": k["launches"], "avg_us": round(k["avg_us"], 2),
                        "achieved_gbs": round(k["achieved_gbs"], 1), "frac": round(k["achieved_gbs"] / peak, 4),
                        "share_of_step": round(k["total_ms"] / ms_prof, 4),
                        "gather_grows_per_s": (None if not k["gather_rows_per_launch"] else
                                               round(k["gather_rows_per_launch"] / (k["avg_us"] * 1e-6) / 1e9, 1)),
                        "gather_frac_of_lsu_ceiling": (None if not k["gather_rows_per_launch"] else round(
                            k["gather_rows_per_launch"] / (k["avg_us"] * 1e-6) / 1e9 / GATHER_CEILING_GROWS, 3))}
                       for k in kernels],
        "gather_ceiling": {"value": GATHER_CEILING_GROWS, "unit": "G rows/s (96-byte rows, LDG.128 by 6-lane groups)",
                           "source": "tools/micro/bulk_gather.cu measured on B200 (L1- or L2-resident table, "
                                     "same rate); profiles/r1_micro_gather.txt"},
        "iteration_kernels_share_of_step": kernel_ms / ms_prof,
    }

    cpu = None
    if world == 1:
        threads = host_threads()
        n_img = threads
        dt = cpu_sample(n_img, threads)
        cpu = {"value": n_img * N * N_ITER / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d VOC-shaped images, one per host thread, oracle restatement of pydensecrf "
                         "(%.2f s wall)" % (n_img, dt),
               "images_per_s": n_img / dt}
        dt1 = cpu_sample(1, 1)   # the same port on ONE core
        cpu["single_core"] = {"value": N * N_ITER / dt1 / 1e6, "unit": UNIT, "images_per_s": 1.0 / dt1}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(B, "wsss_analysis_b200 (libdcrf_b200.so, sm_100a)"),
        "images_per_s": world * B * args.steps / (ms_dev * 1e-3),
        # SURVEY.md 8d asks for both figures: `value` includes the per-image lattice build; this one
        # counts the mean-field iteration kernels only (rank 0's serialised per-kernel event times)
        "iteration_only": {"value": B * N * N_ITER * args.steps / (kernel_ms * 1e-3) / 1e6, "unit": UNIT + " per GPU",
                           "ms_per_step": kernel_ms / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "images_per_s": world * B * args.steps / (ms_e2e * 1e-3),
                "h2d_bytes_per_step": int(U_host.numel() * 4 + I_host.numel()),
                "d2h_bytes_per_step": int(Q_host.numel() * 4),
                "api": "DenseCRFBatch.setUnaryEnergy/addPairwiseGaussian/addPairwiseBilateral/inference "
                       "with pinned host buffers (dcrf_set_unary / dcrf_add_pairwise_* / dcrf_inference, on_device=0); "
                       "driven by wsss_analysis_b200.pipeline.BatchPipeline: %d handles in flight on dedicated streams "
                       "(DCRF_OPT_ASYNC_HOST), each step cut into sub-batches of %s images, copies of one sub-batch "
                       "overlap kernels of the others" % (n_slots, chunk if chunk else "all")},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
