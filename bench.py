#!/usr/bin/env python
"""Benchmark of the motion-blur synthesis hot path (BASELINE.json metric: blurred images/sec, 800x1333 RGB).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3] [--impl reference]

Own arm: one "step" blurs one batch of `--batch` synthetic 3x800x1333 fp32 images (BASELINE config 2: stored-format
128x128 PSFs, param_index 1, low exposure) with ONE tiled-kernel launch; inputs, tap set and outputs are resident in
HBM when the timed region starts (`value`).  `e2e` repeats the measurement through the reference-facing call
(`blur_image_list`) with pinned HOST buffers: H2D of images + dense PSFs, tap compaction, blur, D2H of the results.
`roofline` is the tiled kernel against the measured HBM copy bandwidth; `cpu_baseline` is the oracle's port of the
reference's CPU Fourier blur (--cpu_blur path) on the host cores.

Reference arm (`--impl reference`): the same CPU Fourier port on all host cores, same metric and config.
Multi-GPU: one process per GPU under torchrun, images sharded by rank, no collective in the timed region (weak scaling);
an NCCL all-gather of one checksum per rank runs after it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, C = 800, 1333, 3
ALGO_BYTES_PER_IMAGE = 2 * C * H * W * 4          # read once + write once, fp32
PARAMS = [0.005, 0.001, 0.00005]
FRACTIONS = [1 / 18, 1 / 10, 1 / 5, 1 / 2, 1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-overlap", action="store_true",
                    help="order every step after the previous one (default: consecutive steps are independent batches and "
                         "are launched so that the tail of one overlaps the ramp-up of the next)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg5", "cfg2h"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3 / config-5 / fp16 records and the reference GPU loop")
    ap.add_argument("--min-ms", type=float, default=50.0, help="device time to measure before the median K-step region is taken")
    return ap.parse_args()


def workload_spec(name, batch):
    if name == "cfg2":
        return dict(name="cfg2", batch=batch or 8, param_index=1, exposures=[0, 1, 2],
                    desc="gpu_blur of batch %dx3x800x1333 fp32, stored-format 128x128 PSFs, param_index 1 (expl 0.005), "
                         "low exposure (1/18, 1/10, 1/5)" % (batch or 8))
    if name == "cfg2h":
        return dict(name="cfg2h", batch=batch or 8, param_index=1, exposures=[0, 1, 2], half=True,
                    desc="gpu_blur of batch %dx3x800x1333 fp16 (the dtype the reference's engines pass, engine.py:80): half rows "
                         "widened while they are staged, fp32 accumulation, one rounding at the store; stored-format 128x128 "
                         "PSFs, param_index 1, low exposure" % (batch or 8))
    if name == "cfg5":
        return dict(name="cfg5", batch=batch or 8, param_index=1, exposures=[0, 1, 2], fused_normalize=True,
                    desc="fused blur->normalize of batch %dx3x800x1333 fp32 into the zero-padded %dx3x800x1344 batch that feeds "
                         "GeneralizedRCNNTransform (canonical mean/std), stored-format PSFs, param_index 1, low exposure" % (
                             batch or 8, batch or 8))
    return dict(name="cfg3", batch=batch or 16, param_index=3, exposures=[3, 4],
                desc="gpu_blur of batch %dx3x800x1333 fp32, 128x128 PSFs, param_index 3 (expl 0.00005), "
                     "high exposure (1/2, 1)" % (batch or 16))


def make_trajectories(spec, seed):
    """Seeded camera-shake trajectories, drawn like dataset_utils/generate_PSFs.py:47-48 (Trajectory.fit().fit())."""
    from detectinblur_b200.motion_blur.generate_trajectory import Trajectory
    import random
    np.random.seed(seed)
    random.seed(seed)
    xs, fr = [], []
    for k in range(spec["batch"]):
        e = spec["exposures"][k % len(spec["exposures"])]
        tr = Trajectory(canvas=256, max_len=96, expl=PARAMS[spec["param_index"] - 1]).fit().fit()
        xs.append(tr.x)
        fr.append(FRACTIONS[e])
    return np.stack(xs), np.array(fr)


# ------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        allrows = []
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                row = (ts, float(parts[1]), float(parts[2]), parts[5:9])
            except ValueError:
                continue
            allrows.append(row)
        inside = [r for r in allrows if t0 - 0.05 <= r[0] <= t1 + 0.05] or allrows
        for ts, s, m, flags in inside:
            sm.append(s)
            mx.append(m)
            for nm, fl in zip(names, flags):
                if fl.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    """One host core blurring `count` 800x1333 images with the reference's CPU Fourier algorithm (oracle port)."""
    seed, count, psf = args
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    from oracle import fourier_oracle as fo
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    t0 = time.perf_counter()
    for _ in range(count):
        fo.fourier_blur(img, psf)
    return time.perf_counter() - t0


def cpu_fourier_throughput(psf, images_per_core, pool=None, cores=None):
    """images/s of the CPU Fourier blur with one worker per host core (mirrors DataLoader workers, train.py:203-205)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_worker, [(k, 0, psf) for k in range(cores)])     # start the workers, import scipy
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(k, images_per_core, psf) for k in range(cores)], chunksize=1)
    wall = time.perf_counter() - t0
    if own_pool:
        pool.close()
        pool.join()
    return cores * images_per_core / wall, wall, cores


def baseline_psf(spec):
    """A PSF of the workload's class for the CPU arm, from the oracle's generator (no GPU needed)."""
    from oracle import psf_oracle as po
    np.random.seed(1337)
    e = spec["exposures"][-1]
    p16, _ = po.stored_psf(PARAMS[spec["param_index"] - 1], FRACTIONS[e], np.random)
    return po.crop128(p16).astype(np.float32)


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    psf = baseline_psf(spec)
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_worker, [(k, 0, psf) for k in range(cores)])
    per_core = max(1, -(-spec["batch"] // cores))        # a step = one batch spread over the host cores
    # A step costs ~0.5 s of all host cores, so K and W are honoured up to a wall-clock budget: the run always ends within
    # a couple of minutes, whatever K the caller passes (the own arm's K is sized for 60-us steps).
    budget_s = float(os.environ.get("DIB_REFERENCE_BUDGET_S", "90"))
    t_start = time.perf_counter()
    for _ in range(min(args.warmup, 2)):
        cpu_fourier_throughput(psf, per_core, pool, cores)
    t0 = time.perf_counter()
    n_img = 0
    steps_timed = 0
    for _ in range(args.steps):
        cpu_fourier_throughput(psf, per_core, pool, cores)
        n_img += per_core * cores
        steps_timed += 1
        if steps_timed >= 3 and time.perf_counter() - t_start > budget_s:
            break
    wall = time.perf_counter() - t0
    pool.close()
    pool.join()
    value = n_img / wall
    line = {
        "impl": "reference", "metric": "blurred images/sec (800x1333 RGB)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "steps_timed": steps_timed, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / steps_timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": spec["desc"], "path": "CPU Fourier blur (motion_blur/blur_image.py port, --cpu_blur)",
                   "images_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d images of 800x1333x3 uint8 (%.0f s; a %d-step request is cut at a %.0f s budget), "
                                   "one process per core, scipy fftconvolve" % (steps_timed, per_core * cores, wall, args.steps,
                                                                                budget_s)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- own arm
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class Workload(object):
    """Device-resident inputs, tap set and prepared launches of one BASELINE config on this rank's GPU."""

    def __init__(self, spec, dev, rank, n_rot=3, dense_only=False, pitched=False):
        import torch
        import detectinblur_b200.blur_functions as bf
        import detectinblur_b200.psf_ops as ops
        self.spec, self.dev, self.n_rot = spec, dev, n_rot
        B = self.B = spec["batch"]
        self.half = bool(spec.get("half"))
        self.esize = 2 if self.half else 4
        self.dtype = torch.float16 if self.half else torch.float32
        gen = torch.Generator(device="cpu").manual_seed(1337 + rank)
        # rotating input batches: 3 x 102 MB of inputs (+ outputs) > 126 MB L2, nothing survives between steps
        self.host_batches = [torch.rand((B, C, H, W), generator=gen).to(self.dtype).pin_memory() for _ in range(n_rot)]
        self.batches = [hb.to(dev) for hb in self.host_batches]
        if pitched:      # the same pixels as [:, :, :, :W] views of buffers whose rows start 16-byte aligned (pitch 1344 elements)
            wide = [torch.zeros((B, C, H, 1344), dtype=self.dtype, device=dev) for _ in range(n_rot)]
            for wbuf, b in zip(wide, self.batches):
                wbuf[:, :, :, :W] = b
            self.batches = [wbuf[:, :, :, :W] for wbuf in wide]
        self.fused = bool(spec.get("fused_normalize"))
        # results land in rows that start 16-byte aligned (pitch 1336 floats; 1344 for the padded batch of the fused workload),
        # handed out as [:, :, :W] views -- what blur_batch allocates by default, and like the reference, whose result is a
        # crop view of its padded accumulator (blur_functions.py:69)
        quad = self.quad = 16 // self.esize
        self.out_w = 1344 if self.fused else (W + quad - 1) // quad * quad
        self.outs_rot = [torch.zeros((B, C, H, self.out_w), dtype=self.dtype, device=dev) for _ in range(n_rot)]
        views = [[o[i, :, :, :W] for i in range(B)] for o in self.outs_rot]
        norm_kw = dict(mean=[[0.485, 0.456, 0.406]] * B, std=[[0.229, 0.224, 0.225]] * B) if self.fused else {}
        traj, fracs = make_trajectories(spec, seed=1337 * rank)
        psfs16 = ops.rasterize_psfs(traj, fracs, dev, canvas=256, center=True, out_side=128, dtype=torch.float16)
        self.psfs = psfs16.to(self.dtype)                 # stored-format values (fp16 grid) in the image dtype
        self.host_psfs = self.psfs.cpu().pin_memory()
        self.tapset = ops.compact_taps(self.psfs, normalize=True, dense_only=dense_only)
        self.taps = self.tapset.counts
        self.kernels = sorted({"blur_masked_kernel" if m.prog_group_w == 0 else "blur_tiled_kernel" for m in self.tapset.meta})
        self.plans = [bf.prepare_blur([self.batches[r][i] for i in range(B)], self.tapset, list(range(B)), outs=views[r], **norm_kw)
                      for r in range(n_rot)]

    def step(self, k, overlap):
        self.plans[k % self.n_rot].run(overlap=overlap)      # one dib_blur_batch call: host planning + one launch per tiled kernel in use

    @property
    def algo_bytes(self):
        return ALGO_BYTES_PER_IMAGE * self.B * self.esize // 4

    @property
    def fmas(self):
        return float(sum(self.taps)) * C * H * W


def time_steps(step, K, warmup, min_ms=50.0, max_regions=400):
    """Times regions of exactly K steps with CUDA events until at least `min_ms` of device time and 5 regions have been
    measured; the K launches of a region are replayed from a CUDA graph when capture works, so that the host's launch path
    (ctypes, validation, descriptors) is out of a window that may be only a millisecond long.  Returns the region times."""
    import torch
    for k in range(max(warmup, 3)):
        step(k)
    torch.cuda.synchronize()
    graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(3):
                step(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread-local capture: other threads of the process (NCCL's watchdog, the clock sampler) keep making CUDA calls
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            for k in range(K):
                step(k)
        graph = g
    except Exception as exc:      # capture unsupported for this launch: time direct launches instead
        sys.stderr.write("bench: CUDA graph capture failed (%s); timing direct launches\n" % (str(exc).splitlines()[0],))
        graph = None
        torch.cuda.synchronize()

    def region():
        if graph is not None:
            graph.replay()
        else:
            for k in range(K):
                step(k)

    region()
    torch.cuda.synchronize()
    times = []
    total = 0.0
    while (total < min_ms or len(times) < 5) and len(times) < max_regions:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        region()
        e1.record()
        e1.synchronize()
        t = e0.elapsed_time(e1)
        times.append(t)
        total += t
    return times, graph is not None


def fp32_probe_tflops(dev):
    import torch
    import ctypes
    from detectinblur_b200 import _lib
    sink = torch.empty(148 * 8 * 256 * 2, device=dev)
    cnt = ctypes.c_uint64(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib.dib_fp32_probe(200, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(cnt), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib.dib_fp32_probe(2000, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(cnt), st))
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * cnt.value / (e0.elapsed_time(e1) * 1e-3) / 1e12


def roofline_block(wl, kernel_ms, pipelined_ms, hbm_peak, peak_src, fp32_tflops):
    """The roofline of SURVEY.md 8(d): the slower of bytes at the HBM peak and taps x pixels FMAs at the FP32 peak.
    `kernel_ms` is the duration of one launch ordered after its predecessor (what ncu reports); `pipelined_ms` the time
    per launch when consecutive independent batches overlap tail-to-head (what `value` is made of)."""
    t_hbm = wl.algo_bytes / (hbm_peak * 1e9)
    t_fma = 2.0 * wl.fmas / (fp32_tflops * 1e12)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % wl.spec["name"])
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    common = {"traffic": traffic, "kernel": "dib::" + "+".join(wl.kernels), "kernel_ms": kernel_ms, "kernel_ms_pipelined": pipelined_ms,
              "algorithmic_bytes_per_launch": wl.algo_bytes, "fma_per_launch": wl.fmas, "fp32_probe_tflops": fp32_tflops,
              "hbm_peak_gbs": hbm_peak, "t_hbm_ms": t_hbm * 1e3, "t_fp32_ms": t_fma * 1e3,
              "frac_pipelined": max(t_hbm, t_fma) / (pipelined_ms * 1e-3) if pipelined_ms else None}
    if t_hbm >= t_fma:
        ach = wl.algo_bytes / (kernel_ms * 1e-3) / 1e9
        return dict({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": peak_src}, **common)
    ach = 2.0 * wl.fmas / (kernel_ms * 1e-3) / 1e12
    return dict({"bound": "fp32", "achieved": ach, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": ach / fp32_tflops,
                 "peak_source": "dib_fp32_probe measured in this run (FFMA, non-tensor); nominal 74.4"}, **common)


def e2e_variant(kind, wl, dev, world, steps):
    """The metric end to end through the reference-facing call with pinned HOST buffers: H2D of the step's images and dense
    PSFs, tap compaction, blur, D2H of the results, every step, on three streams taking turns (PCIe is full duplex).
    kind: 'f32' float32 both ways (what the reference's float tensors are), 'f16' half both ways (the dtype engine.py:80
    uploads), 'u8' bytes both ways (uploaded bytes are scaled on the device as to_tensor does, results come back as the
    uint8 image of the --cpu_blur path)."""
    import torch
    import torch.distributed as dist
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops
    B = wl.B
    n_streams = int(os.environ.get("DIB_E2E_STREAMS", "3"))
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    if kind == "u8":
        hin = [(hb.float() * 255).to(torch.uint8).pin_memory() for hb in wl.host_batches]
        hout = [torch.empty((B, C, H, W), dtype=torch.uint8).pin_memory() for _ in range(n_streams)]
        dout = [torch.empty((B, C, H, W), dtype=torch.uint8, device=dev) for _ in range(n_streams)]
        comp = torch.float32
        in_bytes, out_bytes = B * C * H * W, B * C * H * W
    else:
        comp = torch.float32 if kind == "f32" else torch.float16
        es = 4 if kind == "f32" else 2
        quad = 16 // es
        pitch = (W + quad - 1) // quad * quad
        hin = [hb.to(comp).pin_memory() for hb in wl.host_batches]
        hout = [torch.empty((B, C, H, pitch), dtype=comp).pin_memory() for _ in range(n_streams)]
        in_bytes, out_bytes = B * C * H * W * es, B * C * H * pitch * es
    hpsf = wl.host_psfs.to(comp).pin_memory()
    in_bytes += hpsf.numel() * hpsf.element_size()

    host_s = []

    def step(k):
        st = streams[k % n_streams]
        st.synchronize()                      # the step that used this stream's buffers three steps ago is complete
        t_enq = time.perf_counter()
        with torch.cuda.stream(st):
            dbatch = hin[k % wl.n_rot].to(dev, non_blocking=True)
            dpsf = hpsf.to(dev, non_blocking=True)
            if kind == "u8":
                dbatch = ops.u8_to_float(dbatch)
            images = [dbatch[i] for i in range(B)]
            # fp32 compute: no read-back of the PSF summaries (the launch is planned on the device), so the host enqueues the
            # whole step without waiting for the step's own uploads
            bf.blur_image_list(images, [{"blurring": True}] * B, [dpsf[i] for i in range(B)], sync=False)
            # the results of a same-shape batch are [k, :, :, :W] views of ONE row-aligned buffer (blur_batch allocates them
            # so): convert / copy the whole batch in one go instead of image by image
            base = images[0]._base
            whole = base is not None and base.dim() == 4 and all(im._base is base for im in images)
            if kind == "u8":
                if whole:
                    ops.float_to_u8(base.view(B * C, H, base.shape[3])[:, :, :W], out=dout[k % n_streams].view(B * C, H, W))
                else:
                    for i in range(B):
                        ops.float_to_u8(images[i], out=dout[k % n_streams][i])
                hout[k % n_streams].copy_(dout[k % n_streams], non_blocking=True)
            elif whole:
                hout[k % n_streams].copy_(base, non_blocking=True)
            else:
                for i in range(B):
                    r = images[i]
                    full = r.as_strided((C, H, r.stride(1)), (r.stride(0), r.stride(1), 1))
                    hout[k % n_streams][i, :, :, :r.stride(1)].copy_(full, non_blocking=True)
        host_s.append(time.perf_counter() - t_enq)

    for k in range(6):
        step(k)
    torch.cuda.synchronize()
    del host_s[:]
    regions = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            step(k)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        regions.append(t)
    med = float(np.median(regions))
    return {"value": world * B * steps / med, "unit": "images/s", "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
            "steps": steps, "regions_s": regions, "io": kind, "host_enqueue_us_per_step": round(float(np.median(host_s)) * 1e6, 1),
            "api": "blur_image_list(images, blur_dicts, psfs%s) on pinned host buffers, %d streams%s" % (
                ", sync=False", n_streams, "; psf_ops.u8_to_float / float_to_u8 around it" if kind == "u8" else "")}


def reference_gpu_loop(wl, dev):
    """The reference's own GPU loop (unmodified models/blur_functions.py:92-100, staged under baseline/_ref by
    __graft_entry__.build()) on the same device-resident config-2 batch, fp32 and fp16: the path --gpu_blur users run today.
    Reported only."""
    import torch
    ref_root = os.environ.get("DIB_REFERENCE_ROOT") or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "models")):
        return {"unavailable": "no reference tree at %s" % ref_root}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import refshim
        refshim.install(ref_root)
        import models.blur_functions as rbf
    except Exception as exc:
        return {"unavailable": "reference import failed: %s" % (str(exc).splitlines()[0],)}
    out = {"source": "unmodified models/blur_functions.py blur_image_list from %s" % os.path.relpath(ref_root, ROOT), "batches": 3}
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        psfs = [wl.psfs[i].to(dt) for i in range(wl.B)]
        times = []
        for r in range(4):
            images = [wl.batches[r % wl.n_rot][i].to(dt) for i in range(wl.B)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rbf.blur_image_list(images, [{"blurring": True}] * wl.B, list(psfs))
            torch.cuda.synchronize()
            if r > 0:
                times.append(time.perf_counter() - t0)
        out[name] = {"ms_per_batch": 1000.0 * float(np.median(times)), "images_per_s": wl.B / float(np.median(times))}
    return out


def run_own_arm(args, spec):
    import torch
    import torch.distributed as dist
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, max(args.warmup, 3)
    overlap = not args.no_overlap
    wl = Workload(spec, dev, rank)
    B = wl.B

    # ---- the timed region: regions of exactly K steps, repeated until >= 50 ms have been measured; the median region counts
    for k in range(Wm):
        wl.step(k, overlap)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    l0 = bf.launch_count()
    wl.step(0, overlap)
    launches_per_step = bf.launch_count() - l0           # own kernels per step: one per tiled kernel in use
    torch.cuda.synchronize()
    t_wall0 = time.time()
    regions, used_graph = time_steps(lambda k: wl.step(k, overlap), K, Wm, min_ms=args.min_ms)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    mine = [float(np.median(regions)), float(min(regions)), float(max(regions))]
    per_rank = [mine]
    if world > 1:
        t = torch.tensor(mine, device=dev)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank = [[float(v) for v in g.tolist()] for g in gathered]
    region_ms = max(p[0] for p in per_rank)               # MAX over ranks of each rank's median K-step region
    slowest = int(np.argmax([p[0] for p in per_rank]))
    value = world * B * K / (region_ms / 1000.0)

    # ---- the same steps with every launch ordered after the previous one: the kernel's own duration (what ncu reports)
    ordered_regions, _ = time_steps(lambda k: wl.step(k, False), K, 3, min_ms=min(args.min_ms, 30.0)) if overlap else (regions, used_graph)
    kernel_ms = float(np.median(ordered_regions)) / K
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp32_tflops = fp32_probe_tflops(dev)
    roofline = roofline_block(wl, kernel_ms, region_ms / K, hbm_peak, peak_src, fp32_tflops)

    # ---- the other BASELINE configs, measured the same way (records beside the headline, not the metric).  N = 1: all of
    #      them; N > 1: BASELINE config 5 only -- "8 ranks x batch 8, fused blur -> normalize" -- on every rank, MAX over ranks
    extra = {}
    if not args.no_extras and spec["name"] == "cfg2":
        variants = [("cfg3", "cfg3", {}), ("cfg5", "cfg5", {}), ("cfg2h", "cfg2h", {}),
                    # config 2 again: inputs pitched (rows 16-byte aligned), and through the TMA-staged dense kernel (which the
                    # default routing reserves for large PSFs) in both input layouts
                    ("cfg2_pitched_inputs", "cfg2", {"pitched": True}),
                    ("cfg2_tma_kernel", "cfg2", {"dense_only": True}),
                    ("cfg2_tma_kernel_pitched_inputs", "cfg2", {"dense_only": True, "pitched": True})]
        if world > 1:
            variants = [v for v in variants if v[0] == "cfg5"]
        for name, base, kw in variants:
            w2, err, ms_o, ms_s, g_o, k2 = None, None, float("inf"), float("inf"), False, 0
            try:
                sp = workload_spec(base, None)
                w2 = Workload(sp, dev, rank, **kw)
                k2 = max(5, min(K, 40 if name == "cfg3" else 200))
                reg_o, g_o = time_steps(lambda k: w2.step(k, True), k2, 3, min_ms=40.0)
                reg_s, _ = time_steps(lambda k: w2.step(k, False), k2, 3, min_ms=30.0)
                ms_o, ms_s = float(np.median(reg_o)) / k2, float(np.median(reg_s)) / k2
            except Exception as exc:
                err = str(exc).splitlines()[0]
            if world > 1:       # every rank takes part, whatever happened to it: MAX over ranks (inf = a rank failed)
                both = torch.tensor([ms_o, ms_s], device=dev, dtype=torch.float64)
                dist.all_reduce(both, op=dist.ReduceOp.MAX)
                ms_o, ms_s = float(both[0].item()), float(both[1].item())
                if err is None and not np.isfinite(ms_o):
                    err = "another rank failed"
            if err is None:
                extra[name] = {"workload": sp["desc"], "kernels": w2.kernels, "value": world * w2.B / (ms_o / 1000.0), "unit": "images/s",
                               "n_gpus": world, "steps": k2,
                               "ms_per_step": ms_o, "ms_per_step_ordered": ms_s, "taps": w2.taps,
                               "dtype": "f16 i/o, f32 accumulate" if w2.half else "f32", "cuda_graph": g_o,
                               "roofline": roofline_block(w2, ms_s, ms_o, hbm_peak, peak_src, fp32_tflops)}
            else:
                extra[name] = {"error": err}
            del w2
            torch.cuda.empty_cache()

    # ---- end to end through the reference-facing API with HOST buffers
    e2e, e2e_variants = None, {}
    if not args.no_e2e:
        e2e_steps = max(12, min(K, 48))
        e2e = e2e_variant("f32", wl, dev, world, e2e_steps)
        e2e["pcie_note"] = "measured on this pool: 46 GB/s per direction with both directions busy -> 3.6 k img/s ceiling for fp32"
        for kind in ("f16", "u8"):
            try:
                e2e_variants[kind] = e2e_variant(kind, wl, dev, world, e2e_steps)
            except Exception as exc:
                e2e_variants[kind] = {"error": str(exc).splitlines()[0]}

    # ---- the reference's own GPU loop on the same batch (rank 0, N = 1 only)
    ref_loop = None
    if world == 1 and rank == 0 and not args.no_extras and spec["name"] == "cfg2":
        try:
            ref_loop = reference_gpu_loop(wl, dev)
        except Exception as exc:
            ref_loop = {"unavailable": str(exc).splitlines()[0]}

    # ---- optional cross-shard verification: all-gather one checksum per rank (outside every timed region)
    csum = ops.checksum(wl.outs_rot[0])
    sums = [int(csum.item()) & 0xFFFFFFFFFFFFFFFF]
    if world > 1:
        gathered = [torch.zeros_like(csum) for _ in range(world)]
        dist.all_gather(gathered, csum)
        sums = [int(g.item()) & 0xFFFFFFFFFFFFFFFF for g in gathered]

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            psf_host = wl.host_psfs[B - 1].float().numpy().copy()
            cores = os.cpu_count() or 1
            per_core = 3
            v, wall, cores = cpu_fourier_throughput(psf_host, per_core, None, cores)
            cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "%d images of 800x1333x3 (%d per core, one process per core), CPU Fourier blur port "
                             "(oracle/fourier_oracle.py), %.1f s wall" % (cores * per_core, per_core, wall)}
        line = {
            "metric": "blurred images/sec (800x1333 RGB)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": region_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 i/o, f32 accumulate" if wl.half else "f32",
            "data": "synthetic",
            "config": {"workload": spec["desc"], "batch_per_gpu": B, "taps": wl.taps, "kernels": wl.kernels,
                       "timing": ("regions of exactly %d steps, each bracketed by CUDA events, repeated until >= %.0f ms of device time "
                                  "(%d regions here); ms_per_step = MAX over ranks of each rank's MEDIAN region / %d; the launches of a "
                                  "region are %s" % (K, args.min_ms, len(regions), K,
                                                     "replayed from a CUDA graph (host launch path outside the window)" if used_graph
                                                     else "issued directly (graph capture unavailable)")),
                       "step_overlap": ("consecutive steps are independent batches (own inputs, own outputs) launched with "
                                        "programmatic dependent launch: the tail of step k overlaps the ramp-up of step k + 1"
                                        if overlap else "every step is ordered after the previous one"),
                       "ms_per_step_ordered": kernel_ms,
                       "input_layout": "the reference's: a list of contiguous CHW tensors (row pitch 1333 floats = 5332 B)",
                       "output_layout": "rows 16-byte aligned (pitch %d elements), returned as [:, :, :W] views" % wl.out_w,
                       "l2": "3 rotating input batches (3 x %.0f MB in, %.0f MB out) > 126 MB L2" % (
                           B * C * H * W * wl.esize / 1e6, B * C * H * W * wl.esize / 1e6),
                       "parallelism": "images sharded by rank, no collective on the hot path"},
            "regions_ms": {"n": len(regions), "median": mine[0], "min": mine[1], "max": mine[2]},
            "per_rank_region_ms": {"median_min_max": per_rank, "slowest_rank": slowest},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_variants": e2e_variants,
            "gpu_launches": launches_per_step * K,
            "roofline": roofline,
            "extra": extra,
            "reference_gpu_loop": ref_loop,
            "cpu_baseline": cpu,
            "checksums": ["%016x" % s for s in sums],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    spec = workload_spec(args.workload, args.batch)
    if args.impl == "reference":
        run_reference_arm(args, spec)
    else:
        run_own_arm(args, spec)


if __name__ == "__main__":
    main()
