#!/usr/bin/env python
"""bench.py -- DenseCRF mean-field throughput on B200 (metric of BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)
    python bench.py --config dg612x8                         # another BASELINE configuration as the headline

A "step" is one pass of the hot path over one batch of synthetic images of the named configuration:
for every image both permutohedral lattices are built (a new image means a new lattice) and the
configuration's mean-field iterations are run.  The headline workload is `voc32`: BASELINE config 1's
geometry (500x375 RGB, 21 labels, 10 iterations, Gaussian sxy=3 compat=3, bilateral sxy=80 srgb=13
compat=10 -- /root/reference/03a_sec-dsrg/SEC.py:20) as a batch of 32 images per GPU per step.
`value`      : Mpix*iter/s with unaries and images already resident in HBM (device pointers in / out)
`e2e`        : the same batch through the same C-ABI calls with HOST (pinned) buffers: H2D of unaries and
               images and D2H of the marginals Q are inside the timed region
`e2e_labels` : the same with uint8 label maps coming back instead of float32 marginals (what
               dcrf_process / crf_inference_label / the evaluation loops consume)
`configs`    : (N = 1 only) the same measurements for every other BASELINE configuration
`sweep`      : BASELINE config 5 -- 1449 VOC-val-shaped images striped over the ranks, int64 confusion
               matrix summed with one NCCL all-reduce; its SHA-256 must not depend on N
One process per GPU; images shard over ranks with no data-path collective ("weak" scaling: the
per-GPU batch is fixed); rank 0 prints ONE JSON line.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "densecrf_mpix_iter_per_s", "Mpix*iter/s"

# Every BASELINE.json configuration; file:line into /root/reference.
CONFIGS = OrderedDict([
    ("voc32", dict(sizes=[(500, 375)] * 32, L=21, iters=10, g_sxy=3.0, g_compat=3.0, b_sxy=80.0, b_srgb=13.0,
                   b_compat=10.0, image="natural",
                   what="configs[0] geometry (VOC2012 500x375, SEC/DSRG test CRF, 03a_sec-dsrg/SEC.py:20), 32 images per step")),
    ("voc1", dict(sizes=[(500, 375)], L=21, iters=10, g_sxy=3.0, g_compat=3.0, b_sxy=80.0, b_srgb=13.0, b_compat=10.0,
                  image="natural", what="configs[0] literally: one VOC-shaped image per step (03c_hsn/utilities.py:424-442 loop body)")),
    ("sec41x32", dict(sizes=[(41, 41)] * 32, L=21, iters=5, g_sxy=3.0 / 12, g_compat=3.0, b_sxy=80.0 / 12, b_srgb=13.0,
                      b_compat=10.0, image="natural",
                      what="configs[1]: SEC constrain-to-boundary step, 2 x 16 maps of 41x41 (SEC.py:19,270-283)")),
    ("hsn321x16", dict(sizes=[(321, 321)] * 16, L=21, iters=10, g_sxy=1.5, g_compat=3.0, b_sxy=40.0, b_srgb=13.0,
                       b_compat=10.0, image="histo", what="HistoSegNet as run: 321x321, literal CRF set of 03c_hsn/demo.py:159")),
    ("adp1088_morph", dict(sizes=[(1088, 1088)] * 4, L=29, iters=5, g_sxy=1.0, g_compat=20.0, b_sxy=10.0, b_srgb=40.0,
                           b_compat=50.0, image="histo",
                           what="configs[2]: ADP 1088x1088 patches, 29 ADP-morph labels, SEC.py:24-25 CRF, 4 per step")),
    ("adp1088_func", dict(sizes=[(1088, 1088)] * 8, L=5, iters=5, g_sxy=3.0, g_compat=40.0, b_sxy=10.0, b_srgb=4.0,
                          b_compat=25.0, image="histo",
                          what="configs[2]: ADP 1088x1088 patches, 5 ADP-func labels, SEC.py:29-30 CRF, 8 per step")),
    ("dg612x8", dict(sizes=[(612, 612)] * 8, L=6, iters=10, g_sxy=3.0, g_compat=3.0, b_sxy=50.0, b_srgb=5.0,
                     b_compat=10.0, image="natural",
                     what="DeepGlobe as run: 612x612 (03b_irn/step/cam_to_ir_label.py:61,67), IRN label CRF, 8 per step")),
    ("dg2448", dict(sizes=[(2448, 2448)], L=6, iters=10, g_sxy=3.0, g_compat=3.0, b_sxy=80.0, b_srgb=13.0, b_compat=10.0,
                    image="natural", what="configs[3]: one DeepGlobe 2448x2448 tile, 6 labels, bilateral CRF")),
])
HEADLINE = "voc32"
SWEEP_IMAGES, SWEEP_LABELS, SWEEP_BATCH = 1449, 21, 32  # configs[4]; 03b_irn/voc12/val.txt has 1449 lines

# measured LSU row-gather ceiling (tools/micro/bulk_gather.cu, profiles/r1_micro_gather.txt; the TMA
# gather4 path tops out at 127 G rows/s: profiles/r2_micro_gather4.txt)
GATHER_CEILING_GROWS = 106.7
# lattice-construction phases timed by the library (include/dcrf_b200.h, DCRF_K_BUILD_*)
BUILD_PHASES = ("point", "hash", "number", "neigh", "sort", "csr", "norm", "repl")


def npix(cfg):
    return sum(w * h for w, h in cfg["sizes"])


def config_dict(name, cfg):
    w, h = cfg["sizes"][0]
    return {
        "workload": "%s: %d image(s)/GPU/step of %dx%d RGB, %d labels, %d iterations, gaussian sxy=%g compat=%g + "
                    "bilateral sxy=%g srgb=%g compat=%g, lattice build included -- %s"
                    % (name, len(cfg["sizes"]), w, h, cfg["L"], cfg["iters"], cfg["g_sxy"], cfg["g_compat"],
                       cfg["b_sxy"], cfg["b_srgb"], cfg["b_compat"], cfg["what"]),
        "name": name,
        "images_per_gpu_per_step": len(cfg["sizes"]),
        "image": "%s synthetic images, a distinct seed per image (wsss_analysis_b200/synthetic.py)" % cfg["image"],
        "l2_policy": "inputs larger than L2 (unaries alone are %.0f MB per step per GPU)" % (cfg["L"] * npix(cfg) * 4 / 1e6)
                     if cfg["L"] * npix(cfg) * 4 > 126e6 else
                     "an L2 flush (256 MB memset) between timed steps: the step's inputs are smaller than L2",
    }


W_IMG, H_IMG, L_LAB, N_ITER = 500, 375, 21, 10   # the headline geometry (tools/*.py)


def make_inputs(cfg, seed0=0):
    """One distinct seeded image + unary per batch slot.  cfg: a CONFIGS entry, or a batch size of the
    headline geometry (tools/*.py)."""
    from wsss_analysis_b200 import synthetic as S

    if isinstance(cfg, int):
        cfg = dict(CONFIGS[HEADLINE], sizes=[(W_IMG, H_IMG)] * cfg)

    gen = getattr(S, cfg["image"] + "_image")
    imgs = [gen(h, w, seed0 + i) for i, (w, h) in enumerate(cfg["sizes"])]
    unaries = [S.random_unary(cfg["L"], w * h, seed0 + i) for i, (w, h) in enumerate(cfg["sizes"])]
    return imgs, unaries


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement of pydensecrf (pydensecrf itself cannot be installed
# here: SURVEY.md section 0.2), one image per thread on all host cores.
# ------------------------------------------------------------------------------------------------
def cpu_one_image(args):
    from oracle import oracle as O

    cfg, w, h, img, U = args
    d = O.DenseCRF2D(w, h, cfg["L"])
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
    d.addPairwiseBilateral(sxy=cfg["b_sxy"], srgb=cfg["b_srgb"], rgbim=img, compat=cfg["b_compat"])
    return d.inference(cfg["iters"])


def cpu_sample(cfg, inputs, n_images, threads):
    """Wall time of `n_images` CRFs of `cfg` over `threads` host threads (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    O.lib()
    imgs, unaries = inputs
    B = len(cfg["sizes"])
    work = [(cfg,) + tuple(cfg["sizes"][i % B]) + (imgs[i % B], unaries[i % B]) for i in range(n_images)]
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
    name = args.config
    cfg = CONFIGS[name]
    threads = host_threads()
    B = len(cfg["sizes"])
    inputs = make_inputs(cfg)
    pix = npix(cfg)
    for _ in range(1 if args.warmup >= 1 else 0):
        cpu_sample(cfg, inputs, min(B, threads), threads)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_sample(cfg, inputs, B, threads)   # CRF time only: generating the synthetic inputs is not the path
    value = args.steps * pix * cfg["iters"] / dt / 1e6
    sample = "%d steps of the full %d-image batch over %d host threads (one image per task)" % (args.steps, B, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(name, cfg),
        "implementation": "oracle port of pydensecrf on the host cores (oracle/densecrf_oracle.c)",
        "images_per_s": args.steps * B / dt,
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
# roofline accounting
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(cls, d, N, L, M, n_terms_info=None):
    """Algorithmic bytes of ONE launch; float32 values, int32 ids, L real labels
    (padding lanes are not counted)."""
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


def iteration_bytes(N, L, lattices):
    """bytes_iter = 12 L N + sum_k [16 (d_k+1) N + 8 L M_k + (d_k+1)(8 L M_k + 8 M_k)]."""
    b = 12 * L * N
    for d, M in lattices:
        b += 16 * (d + 1) * N + 8 * L * M + (d + 1) * (8 * L * M + 8 * M)
    return b


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def csrc_digest():
    """SHA-256 over the CUDA sources that define the captured (per-iteration) kernels and their data
    layout: profiles/r2_traffic.json is only quoted while the kernels it was captured on are unchanged."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "wsss_analysis_b200", "csrc")
    for f in ("common.cuh", "filter.cu", "softmax_ref.cuh"):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
class Runner(object):
    """Measurements of one configuration on one rank."""

    def __init__(self, name, cfg, local, world, dist, stream, n_slots=3):
        import torch

        self.torch, self.dist = torch, dist
        self.name, self.cfg, self.local, self.world = name, cfg, local, world
        self.dev = torch.device("cuda", local)
        self.stream = stream
        self.n_slots = n_slots
        self.B, self.N = len(cfg["sizes"]), npix(cfg)
        L = cfg["L"]
        self.inputs = make_inputs(cfg)
        imgs, unaries = self.inputs
        self.U_host = torch.from_numpy(np.concatenate([u.ravel() for u in unaries])).pin_memory()
        self.I_host = torch.from_numpy(np.concatenate([im.ravel() for im in imgs])).pin_memory()
        self.U_dev, self.I_dev = self.U_host.to(self.dev), self.I_host.to(self.dev)
        self.Q_dev = torch.empty(self.N * L, dtype=torch.float32, device=self.dev)
        self.flush = None
        if L * self.N * 4 <= 126e6:   # inputs fit L2: flush it between timed steps
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.prof = {"splat": {}, "blur": {}, "slice": {}}
        self.build_prof = OrderedDict()   # (phase, d) -> total ms over the profiled steps
        self.lattice_M = {}
        self.arith = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step(self, profile=False, ev_mid=None):
        from wsss_analysis_b200 import densecrf as G

        cfg = self.cfg
        crf = G.DenseCRFBatch(cfg["sizes"], cfg["L"], device=self.local, stream=self.stream)
        if profile:
            crf.profile_enable(True)
        crf.setUnaryEnergy(self.U_dev)
        crf.addPairwiseGaussian(sxy=cfg["g_sxy"], compat=cfg["g_compat"])
        crf.addPairwiseBilateral(sxy=cfg["b_sxy"], srgb=cfg["b_srgb"], rgbim=self.I_dev, compat=cfg["b_compat"])
        if ev_mid is not None:
            ev_mid.record(self.stream)
        crf.inference_device(cfg["iters"], out=self.Q_dev)
        if profile:
            self.arith = crf.arithmetic()
            for k in range(2):
                d, M, _ = crf.lattice_info(k)
                self.lattice_M[d] = M
            for cls, cid in (("splat", 0), ("blur", 1), ("slice", 2)):
                for tag in ((2, 5) if cls != "slice" else (2,)):
                    ms, n = crf.profile_read(cid, tag)
                    a = self.prof[cls].setdefault(tag, [0.0, 0])
                    a[0] += ms
                    a[1] += n
            for cid, phase in enumerate(BUILD_PHASES, start=3):
                for tag in (2, 5):
                    ms, n = crf.profile_read(cid, tag)
                    if n:
                        self.build_prof[(phase, tag)] = self.build_prof.get((phase, tag), 0.0) + ms
        crf.close()

    def timed(self, steps, profile):
        """-> (total ms, build ms, launches): CUDA events on the launching stream, max over ranks."""
        from wsss_analysis_b200 import densecrf as G

        torch = self.torch
        self.barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        l0 = G.launch_count()
        for i in range(steps):
            if self.flush is not None:
                self.flush.zero_()
                torch.cuda.synchronize()
            ev[i][0].record(self.stream)
            self.step(profile, ev[i][1])
            ev[i][2].record(self.stream)
        self.barrier()
        ms = sum(e[0].elapsed_time(e[2]) for e in ev)
        build = sum(e[0].elapsed_time(e[1]) for e in ev)
        self.build_share = build / ms
        if self.flush is None:   # back-to-back steps: one event pair over the whole region
            ms = ev[0][0].elapsed_time(ev[-1][2])
        return self.max_over_ranks(ms), build, G.launch_count() - l0

    def timed_e2e(self, pipe, steps, labels):
        """`steps` batches through the host-buffer pipeline (n_slots batches in flight): wall clock from
        the first submit to the last result in host memory."""
        cfg = self.cfg
        crf_cfg = {"g_sxy": cfg["g_sxy"], "g_compat": cfg["g_compat"], "bi_sxy": cfg["b_sxy"], "bi_srgb": cfg["b_srgb"],
                   "bi_compat": cfg["b_compat"], "iterations": cfg["iters"]}
        outs = self.slot_lab if labels else self.slot_Q
        self.barrier()
        t0 = time.perf_counter()
        tickets = [pipe.submit(cfg["sizes"], cfg["L"], self.U_host.numpy(), self.I_host.numpy(), crf_cfg,
                               out=outs[i % self.n_slots].numpy(), labels=labels) for i in range(steps)]
        for t_ in tickets:
            pipe.result(t_)
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) * 1e3)

    def measure(self, steps, warmup, e2e=True):
        from wsss_analysis_b200.pipeline import BatchPipeline

        torch, cfg = self.torch, self.cfg
        # warm-up: at least max(W, 3) steps, then until two consecutive steps agree within 5 % (the
        # library's stream-ordered memory pool keeps growing for a few steps on the larger
        # configurations, and a growing pool stalls cudaMallocFromPoolAsync for 10-100 ms), at most 12
        prev, n_warm = None, 0
        while n_warm < 12:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self.step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            n_warm += 1
            if n_warm >= max(warmup, 3) and prev is not None and abs(dt - prev) <= 0.05 * prev:
                break
            prev = dt
        self.n_warm = n_warm
        ms_dev, build_ms, launches = self.timed(steps, False)
        build_share = self.build_share
        # per-kernel pass for the roofline: same steps with the library's CUDA-event pairs around every
        # launch; profiling serialises the two pairwise filters (they overlap on two streams otherwise)
        ms_prof, _, _ = self.timed(steps, True)
        r = {"ms_dev": ms_dev, "build_ms": build_ms, "build_share": build_share, "ms_prof": ms_prof, "launches": launches}
        if e2e:
            L = cfg["L"]
            self.slot_Q = [torch.empty(self.N * L, dtype=torch.float32).pin_memory() for _ in range(self.n_slots)]
            self.slot_lab = [torch.empty(self.N, dtype=torch.uint8).pin_memory() for _ in range(self.n_slots)]
            # every 32-image step goes through the pipeline as two 16-image sub-batches (upload of one
            # overlaps the kernels of the other inside the step; measured 28.8 vs 29.7 ms per step)
            chunk = int(os.environ.get("BENCH_CHUNK", "16")) or None
            if self.B < 32:
                chunk = None
            pipe = BatchPipeline(n_slots=self.n_slots, device=self.local, chunk_images=chunk)
            self.timed_e2e(pipe, 2 * self.n_slots, False)  # warm-up of the slot streams (memory pools, pinned buffers)
            r["ms_e2e"] = self.timed_e2e(pipe, steps, False)
            self.timed_e2e(pipe, self.n_slots, True)
            r["ms_e2e_labels"] = self.timed_e2e(pipe, steps, True)
            r["chunk"] = chunk
            pipe.close()
            self.slot_Q = self.slot_lab = None
        return r

    def report(self, r, steps, peak):
        """Per-configuration JSON object (rank 0)."""
        cfg, N, L, world = self.cfg, self.N, self.cfg["L"], self.world
        pix_iter = world * N * cfg["iters"] * steps
        value = pix_iter / (r["ms_dev"] * 1e-3) / 1e6
        lattices = [(d, self.lattice_M[d]) for d in sorted(self.lattice_M)]
        kernels = []
        for cls in ("splat", "blur", "slice"):
            for tag, (ms, n) in self.prof[cls].items():
                if n == 0:
                    continue
                if cls == "slice":
                    by = algorithmic_bytes("slice", None, N, L, None, lattices)
                    name = "slice_softmax_kernel (fused %d terms)" % tag
                    rows = sum(N * (d + 1) for d, _ in lattices)
                else:
                    by = algorithmic_bytes(cls, tag, N, L, self.lattice_M[tag])
                    name = "%s_kernel d=%d" % (cls, tag)
                    rows = N * (tag + 1) if cls == "splat" else None
                kernels.append({"kernel": name, "launches": n, "total_ms": ms, "avg_us": ms / n * 1e3,
                                "algorithmic_bytes_per_launch": by, "achieved_gbs": by / (ms / n * 1e-3) / 1e9,
                                "gather_rows_per_launch": rows})
        kernels.sort(key=lambda k: -k["total_ms"])
        kernel_ms = sum(k["total_ms"] for k in kernels)
        b_iter = iteration_bytes(N, L, lattices)
        ceiling = peak * 1e9 / (b_iter / N) / 1e6   # Mpix*iter/s per GPU if every algorithmic byte moved at the HBM peak
        iter_only = N * cfg["iters"] * steps / (kernel_ms * 1e-3) / 1e6
        build_ms, build_share = r["build_ms"] / steps, r["build_share"]
        build_timing = "CUDA events on the handle's stream around the add-pairwise calls (timed pass)"
        if N <= 1000000:
            # small handles enqueue the first halves of their lattice builds on side streams and finish them
            # inside the inference call (api.cu, concurrent builds): an event after the add calls sees nothing
            build_ms = sum(self.build_prof.values()) / steps
            build_share = build_ms / (r["ms_dev"] / steps)
            build_timing = ("sum of the per-phase times of the profiled pass (serial); in the timed pass the lattices "
                            "of this configuration build concurrently on side streams and overlap")
        out = {
            "value": value, "unit": UNIT, "ms_per_step": r["ms_dev"] / steps,
            "images_per_s": world * self.B * steps / (r["ms_dev"] * 1e-3),
            "arithmetic": self.arith,
            "lattice": {"pixels": N, "labels": L, "vertices": {("d%d" % d): M for d, M in lattices}},
            "build_ms_per_step": build_ms, "build_share": build_share, "build_timing": build_timing,
            "build_phases_ms_per_step": {"%s_d%d" % k: round(v / steps, 4) for k, v in self.build_prof.items()},
            "iteration_only": {"value": iter_only, "unit": UNIT + " per GPU", "ms_per_step": kernel_ms / steps},
            "step_roofline": {
                "bytes_per_pixel_iteration": b_iter / N, "ceiling": ceiling, "unit": UNIT + " per GPU",
                "frac": value / world / ceiling, "iteration_only_frac": iter_only / ceiling,
                "formula": "SURVEY.md 8(d): 12LN + sum_k[16(d+1)N + 8LM + (d+1)(8LM + 8M)] at the measured HBM peak"},
            "per_kernel": [{"kernel": k["kernel"], "launches": k["launches"], "avg_us": round(k["avg_us"], 2),
                            "achieved_gbs": round(k["achieved_gbs"], 1), "frac": round(k["achieved_gbs"] / peak, 4),
                            "share_of_step": round(k["total_ms"] / r["ms_prof"], 4),
                            "gather_frac_of_lsu_ceiling": (None if not k["gather_rows_per_launch"] else round(
                                k["gather_rows_per_launch"] / (k["avg_us"] * 1e-6) / 1e9 / GATHER_CEILING_GROWS, 3))}
                           for k in kernels],
        }
        if "ms_e2e" in r:
            h2d = int(self.U_host.numel() * 4 + self.I_host.numel())
            out["e2e"] = {"value": pix_iter / (r["ms_e2e"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": r["ms_e2e"] / steps,
                          "images_per_s": world * self.B * steps / (r["ms_e2e"] * 1e-3),
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(N * L * 4)}
            out["e2e_labels"] = {"value": pix_iter / (r["ms_e2e_labels"] * 1e-3) / 1e6, "unit": UNIT,
                                 "ms_per_step": r["ms_e2e_labels"] / steps,
                                 "images_per_s": world * self.B * steps / (r["ms_e2e_labels"] * 1e-3),
                                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(N),
                                 "result": "uint8 argmax label maps (dcrf_map_u8)"}
        return out, kernels, kernel_ms


def host_dma_probe(torch, dev, world, dist):
    """Aggregate pinned-memory copy rate of all ranks at once, both directions (the ceiling of the
    host-buffer legs at N > 1: the host's DMA rate, not the GPUs, bounds them)."""
    n = 256 << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = 0.0
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if rep:
            best = max(best, 8 * n / dt / 1e9)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def run_ours(args):
    import torch
    import torch.distributed as dist

    from wsss_analysis_b200 import densecrf as G
    from wsss_analysis_b200 import sweep as SW

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1
    # (NCCL prints its version banner there when a communicator is created) are sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own banner ("NCCL version ...") goes to stdout by default; stdout carries the JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    # a real (non-default) stream: the library launches on it and the CUDA events bracket it
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    peak, peak_src = hbm_peak()
    steps = args.steps

    name = args.config
    head = Runner(name, CONFIGS[name], local, world, dist, stream)
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_CLOCKS"):
        sampler.start()
    r = head.measure(steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    dma = host_dma_probe(torch, dev, world, dist)

    # BASELINE config 5: the sharded sweep with its one collective (every N, all ranks)
    if world > 1:
        dist.barrier()
    sw = None
    if not args.no_sweep:
        from wsss_analysis_b200.evaluation import CollectiveComm

        comm = CollectiveComm(rank, world, local) if world > 1 else None   # ncclComm_t made through the C ABI
        # untimed warm-up on the first images of this rank's shard (memory pool, kernel images)
        SW.run_sweep_device(3 * SWEEP_BATCH * world, SWEEP_LABELS, rank, world, batch=SWEEP_BATCH, device=local,
                            all_reduce=False, verify=False)
        sr = SW.run_sweep_device(SWEEP_IMAGES, SWEEP_LABELS, rank, world, batch=SWEEP_BATCH, device=local, comm=comm)
        if comm is not None:
            comm.close()
        secs = torch.tensor([sr["seconds"]], dtype=torch.float64, device=dev)
        tot = torch.tensor([sr["pixels"], sr["images"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        conf = sr["confusion"]
        sw = {"workload": "configs[4]: %d VOC2012-val-shaped images (mixed sizes), %d labels, 10 iterations, image i -> rank "
                          "i mod N, batches of %d, device-resident inputs" % (SWEEP_IMAGES, SWEEP_LABELS, SWEEP_BATCH),
              "images": int(tot[1].item()), "seconds_max_over_ranks": float(secs.item()),
              "images_per_s": int(tot[1].item()) / float(secs.item()),
              "value": int(tot[0].item()) * 10 / float(secs.item()) / 1e6, "unit": UNIT,
              "collective": "one SUM all-reduce of the (C+1, C) int64 confusion matrix (%s)" % (
                  "dcrf_confusion_allreduce = ncclAllReduce(int64) on a communicator of %d ranks made through the C ABI" % world
                  if world > 1 else "single rank: no-op"),
              "confusion_sha256": hashlib.sha256(conf.tobytes()).hexdigest(), "pixels_counted": int(conf.sum()),
              "miou_irn": sr["miou_irn"], "miou_sec": sr["miou_sec"],
              "verified": "every rank's local matrix equals np.bincount of its downloaded label maps (bit-exact)"}

    # the other BASELINE configurations (one GPU: they are per-GPU figures)
    others = OrderedDict()
    if world == 1 and not args.no_configs:
        for nm, cfg in CONFIGS.items():
            if nm == name:
                continue
            rn = Runner(nm, cfg, local, world, dist, stream)
            rr = rn.measure(max(3, min(steps, 5)), 3)
            obj, _, _ = rn.report(rr, max(3, min(steps, 5)), peak)
            if not args.no_cpu:
                dt1 = cpu_sample(cfg, rn.inputs, 1, 1)   # the CPU port on ONE image, one core
                w0, h0 = cfg["sizes"][0]
                obj["cpu_port_one_core"] = {"value": w0 * h0 * cfg["iters"] / dt1 / 1e6, "unit": UNIT,
                                            "images_per_s": 1.0 / dt1}
            obj["what"] = cfg["what"]
            others[nm] = obj
            del rn
            torch.cuda.empty_cache()
            G.trim_memory()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    obj, kernels, kernel_ms = head.report(r, steps, peak)
    top = kernels[0]
    traffic, traffic_note = None, "no ncu capture of this build committed"
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("config") == name and tj.get("csrc_sha256_16") == csrc_digest():
            traffic = tj["kernels"].get(top["kernel"])
            traffic_note = "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch, %s" % tj.get("source", "")
        else:
            traffic_note = "profiles/r2_traffic.json was captured on other sources / another config: not quoted"
    roofline = {
        "bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak, "unit": "GB/s",
        "frac": top["achieved_gbs"] / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": top["algorithmic_bytes_per_launch"],
        "share_of_step": top["total_ms"] / r["ms_prof"],
        "whole_step_frac_of_hbm_ceiling": obj["step_roofline"]["frac"],
        "timing": "CUDA events recorded by the library on the launching stream around every launch, in a separate "
                  "pass of the same %d steps run right after the timed region (%.2f ms/step profiled and "
                  "serialised vs %.2f ms/step timed)" % (steps, r["ms_prof"] / steps, r["ms_dev"] / steps),
        "per_kernel": obj["per_kernel"],
        "gather_ceiling": {"value": GATHER_CEILING_GROWS, "unit": "G rows/s (96-byte rows, LDG.128 by 6-lane groups)",
                           "source": "tools/micro/bulk_gather.cu measured on B200 (L1- or L2-resident table, "
                                     "same rate); profiles/r1_micro_gather.txt; TMA gather4: profiles/r2_micro_gather4.txt"},
        "iteration_kernels_share_of_step": kernel_ms / r["ms_prof"],
    }

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = host_threads()
        cfg = CONFIGS[name]
        dt = cpu_sample(cfg, head.inputs, head.B, threads)
        cpu = {"value": head.N * cfg["iters"] / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "one full %d-image step, one image per task over %d host threads, oracle restatement of "
                         "pydensecrf (%.2f s wall)" % (head.B, threads, dt),
               "images_per_s": head.B / dt}
        dt1 = cpu_sample(cfg, head.inputs, 1, 1)   # the same port on ONE core
        w0, h0 = cfg["sizes"][0]
        cpu["single_core"] = {"value": w0 * h0 * cfg["iters"] / dt1 / 1e6, "unit": UNIT, "images_per_s": 1.0 / dt1}

    e2e = obj.pop("e2e")
    e2e["api"] = ("DenseCRFBatch.setUnaryEnergy/addPairwiseGaussian/addPairwiseBilateral/inference with pinned host "
                  "buffers (dcrf_set_unary / dcrf_add_pairwise_* / dcrf_inference, on_device=0); driven by "
                  "wsss_analysis_b200.pipeline.BatchPipeline: %d handles in flight on dedicated streams "
                  "(DCRF_OPT_ASYNC_HOST), each step cut into sub-batches of %s images" % (head.n_slots, r.get("chunk") or "all"))
    per_gpu_bytes = e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]
    e2e["host_dma"] = {"aggregate_gbs_all_ranks_both_directions": dma,
                       "needed_gbs_at_device_rate": world * per_gpu_bytes / (r["ms_dev"] / steps * 1e-3) / 1e9,
                       "frac_of_host_dma_ceiling": (world * per_gpu_bytes / (r["ms_e2e"] / steps * 1e-3) / 1e9) / dma,
                       "limiter": "host DMA (aggregate pinned-copy rate of the box)" if
                                  world * per_gpu_bytes / (r["ms_dev"] / steps * 1e-3) / 1e9 > 0.9 * dma else "GPU kernels"}
    e2e_labels = obj.pop("e2e_labels")
    line = {
        "metric": METRIC, "value": obj["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "warmup_steps_run": head.n_warm, "ms_per_step": obj["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, CONFIGS[name]),
        "implementation": "wsss_analysis_b200 (libdcrf_b200.so, hand-written CUDA for sm_100a) through the C ABI",
        "arithmetic": obj["arithmetic"],
        "images_per_s": obj["images_per_s"],
        "lattice": obj["lattice"], "build_ms_per_step": obj["build_ms_per_step"],
        "build_phases_ms_per_step": obj["build_phases_ms_per_step"],
        # SURVEY.md 8d asks for both figures: `value` includes the per-image lattice build; this one
        # counts the mean-field iteration kernels only (rank 0's serialised per-kernel event times)
        "iteration_only": obj["iteration_only"],
        "step_roofline": obj["step_roofline"],
        "clocks": clocks,
        "e2e": e2e, "e2e_labels": e2e_labels,
        "gpu_launches": int(r["launches"]),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sweep": sw,
        "configs": others if others else None,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=HEADLINE, choices=list(CONFIGS.keys()))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration summary (N = 1)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 1449-image sharded sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-port legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
