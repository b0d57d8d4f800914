#!/usr/bin/env python
"""Headline benchmark: megapixels/s of (normalise + 7-step enhancement with safeguards +
compute_metrics + compute_validation) over a synthetic 512x512x1024 uint16 CT stack per GPU
(BASELINE.json configs[1]; weak scaling: every rank processes its own 1024-slice stack, C4 style,
followed by one NCCL all-gather of the per-slice result rows).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference ...                     # the reference's CPU path (oracle
                                                             # restatement: skimage/pywt absent)
Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "megapixels/s (enhance+metrics)"
WORKLOAD = ("C2: synthetic 512x512x1024 uint16 CT stack per GPU; normalize_image + "
            "apply_enhancements_from_params(P_full: denoise, clahe, gamma, unsharp, post_denoise, "
            "bilateral d=5, tv_denoise w=0.05; 3 safeguards) + compute_metrics(enhanced) + "
            "compute_validation(original, enhanced)")
H = W = 512
HBM_FALLBACK_GBS = 6650.0


def _gen_slice(args):
    from mdimg_b200 import synth
    seed, z = args
    return synth.ct_slice(seed, z)


def make_stack(n: int, seed0: int) -> np.ndarray:
    """[n, 512, 512] uint16, seeds seed0 + z (cached under the temp dir)."""
    cache = Path(tempfile.gettempdir()) / f"mdimg_ct_{n}_{seed0}.npy"
    if cache.exists():
        try:
            arr = np.load(cache)
            if arr.shape == (n, H, W):
                return arr
        except Exception:  # noqa: BLE001
            pass
    from multiprocessing import get_context
    jobs = [(seed0 + z, z / n) for z in range(n)]
    cores = len(os.sched_getaffinity(0))
    with get_context("fork").Pool(min(cores, 16)) as pool:
        slices = pool.map(_gen_slice, jobs, chunksize=8)
    arr = np.stack(slices)
    try:
        np.save(cache, arr)
    except Exception:  # noqa: BLE001
        pass
    return arr


# ------------------------------------------------------------------------------------------
# CPU reference arm (oracle restatement of the reference's own functions)
# ------------------------------------------------------------------------------------------
def _cpu_one(raw_slice):
    from mdimg_b200 import synth
    from oracle import ref_enhancement as oenh
    from oracle import ref_metrics as omet
    import warnings
    warnings.filterwarnings("ignore")
    x = omet.normalize_image(raw_slice)
    enh, _ = oenh.apply_enhancements_from_params(x, synth.plan_full())
    omet.compute_metrics(enh)
    omet.compute_validation(x, enh)
    return x.size


def sample_slices(n_slices: int, n_total: int = 1024, seed0: int = 1000):
    """`n_slices` slices spread evenly over the bench stack (same seeds / anatomy as make_stack)."""
    from mdimg_b200 import synth
    idx = np.linspace(0, n_total - 1, n_slices).astype(int)
    return np.stack([synth.ct_slice(seed0 + int(z), int(z) / n_total) for z in idx])


_cpu_pool = None


def cpu_pool(cores: int):
    """One process per core, created once: the oracle (numpy / scipy) is imported in the parent
    first, so the forked workers start warm and no interpreter / import time lands in a step."""
    global _cpu_pool
    if _cpu_pool is None and cores > 1:
        from multiprocessing import get_context
        import oracle.ref_enhancement  # noqa: F401  (pre-import for the forked workers)
        import oracle.ref_metrics  # noqa: F401
        from mdimg_b200 import synth
        synth.plan_full()
        _cpu_pool = get_context("fork").Pool(cores)
        _cpu_pool.map(_cpu_one, [synth.ct_slice(999, 0.5, size=64)] * cores, chunksize=1)   # touch every worker
    return _cpu_pool


def cpu_throughput(stack: np.ndarray, n_slices: int, cores: int):
    """Mpx/s of the restated reference on `n_slices` slices using `cores` processes."""
    sample = [stack[i] for i in np.linspace(0, stack.shape[0] - 1, n_slices).astype(int)]
    pool = cpu_pool(cores) if cores > 1 else None
    t0 = time.perf_counter()
    if pool is not None:
        px = sum(pool.map(_cpu_one, sample, chunksize=1))
    else:
        px = sum(_cpu_one(s) for s in sample)
    dt = time.perf_counter() - t0
    return px / dt / 1e6, dt


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    per_step = max(4 * cores, 32)            # ~10 s of CPU work per step
    stack = sample_slices(per_step)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_throughput(stack, min(per_step, cores), cores)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt = cpu_throughput(stack, per_step, cores)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    _cpu_one(stack[0])
    v1, dt1 = cpu_throughput(stack, 6, 1)        # single-core rate (BASELINE.md 4.3 asks for both)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_slices_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": cores, "kind": "port",
                         "one_core": {"value": v1, "unit": "Mpx/s", "cores": 1,
                                      "sample": f"6 slices in one process, {dt1:.1f} s wall"},
                         "sample": f"{per_step} of 1024 slices per step, one warm process per core; "
                                   "restated reference (numpy+scipy oracle; scikit-image/PyWavelets "
                                   "are not installed, so the reference itself cannot be imported)"},
        "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.

    NVML is read in-process (nvidia_ml_py) from a thread every 50 ms: two cheap driver calls per
    sample.  A polling `nvidia-smi -lms` child was measurably intrusive here -- its start-up and its
    multi-field queries stalled the second timed step by 10-100 ms -- and remains only the fallback."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (monotonic time, sm_mhz, max_mhz, set of reason names)
        self.thread = None
        self._stop = threading.Event()
        self.source = None

    # ---- NVML in-process ----
    def _nvml_loop(self, nv, handle):
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            mx = None
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                bits = int(get_reasons(handle))
                self.lines.append((time.monotonic(), sm, mx, {n for b, n in names.items() if bits & b}))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:  # noqa: BLE001
                pass
        return self.index

    def wait_ready(self, timeout: float = 5.0) -> None:
        """Block until the first sample has arrived (the sampler's start-up is over)."""
        t_end = time.monotonic() + timeout
        while self.thread is not None and not self.lines and time.monotonic() < t_end:
            time.sleep(0.02)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 9:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            reasons = {n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")}
            self.lines.append((time.monotonic(), sm, mx, reasons))

    def stop(self, t0: float = float("-inf"), t1: float = float("inf")):
        """Summary of the samples taken inside [t0, t1] (monotonic clock) -- the timed region."""
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        inside = [ln for ln in self.lines if t0 <= ln[0] <= t1]
        if not inside:                      # region shorter than the sampling period: nearest samples
            inside = self.lines[-2:]
        sm = [ln[1] for ln in inside]
        mx = [ln[2] for ln in inside if ln[2] is not None]
        reasons = set()
        for ln in inside:
            reasons |= ln[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def measured_hbm_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# algorithmic bytes per pixel and call of every C-ABI operator of the step
ALG_BYTES_PER_PX = {
    "mdimg_normalize_u16": 8, "mdimg_metrics": 4, "mdimg_fullref": 8, "mdimg_wavelet_denoise": 12,
    "mdimg_light_denoise": 16, "mdimg_clahe": 16, "mdimg_clahe_gamma": 16, "mdimg_gamma": 8, "mdimg_unsharp": 8,
    "mdimg_bilateral": 8, "mdimg_clip01": 8, "mdimg_estimate_sigma": 4, "mdimg_quality": 4, "mdimg_axpby": 12,
    "mdimg_copy": 8, "mdimg_export_u16": 6,
}


def operator_roofline(ops, raw_chunk, plan, peak):
    """One chunk through the step with CUDA events around every C-ABI call (the torch-side control flow issues
    the same operator calls that `mdimg_enhance` issues internally): per operator the device time, SURVEY
    8(d)'s algorithmic bytes, achieved GB/s and the fraction of the measured HBM peak.  TV-Chambolle is
    counted per launch as built (20 B/px per two-body launch -> 10 B/px per body, + 4 B/px result write)."""
    import torch
    from collections import defaultdict
    from mdimg_b200.batch import process_stack
    from mdimg_b200.stack import StackOps

    n = int(raw_chunk.shape[0])
    px = float(n) * H * W
    prev_env = os.environ.get("MDIMG_NATIVE_ENGINE")
    os.environ["MDIMG_NATIVE_ENGINE"] = "0"
    orig_call = StackOps._call
    try:
        process_stack(raw_chunk, plan, chunk=n, ops=ops)          # warm-up in this mode
        torch.cuda.synchronize()
        events = []

        def timed_call(self, fn, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(self, fn, *a)
            e1.record()
            events.append((getattr(fn, "__name__", str(fn)), e0, e1))

        StackOps._call = timed_call
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        res = process_stack(raw_chunk, plan, chunk=n, ops=ops)
        t1.record()
        torch.cuda.synchronize()
    finally:
        StackOps._call = orig_call
        if prev_env is None:
            os.environ.pop("MDIMG_NATIVE_ENGINE", None)
        else:
            os.environ["MDIMG_NATIVE_ENGINE"] = prev_env
    ms, calls = defaultdict(float), defaultdict(int)
    for name, e0, e1 in events:
        ms[name] += e0.elapsed_time(e1)
        calls[name] += 1
    it_mean = float(res.tv_iterations.mean())
    total = t0.elapsed_time(t1)
    out = []
    for name, t in sorted(ms.items(), key=lambda kv: -kv[1]):
        b = 10.0 * it_mean + 4.0 if name == "mdimg_tv_chambolle" else ALG_BYTES_PER_PX.get(name)
        entry = {"op": name, "calls": calls[name], "ms": round(t, 4), "share": round(t / total, 4)}
        if b is not None:
            gbs = b * px * calls[name] / (t / 1e3) / 1e9
            entry.update(alg_bytes_per_px_per_call=b, achieved_gbs=round(gbs, 1), frac=round(gbs / peak, 4))
        out.append(entry)
    return {"slices": n, "pixels": px, "total_ms": round(total, 3), "tv_iterations_mean": it_mean, "peak_gbs": peak,
            "basis": "SURVEY.md 8(d) algorithmic bytes x pixels x calls / CUDA-event time; "
                     "frac = achieved / peak; TV per launch as built (10 B/px per body + 4)",
            "ops": out}


def gpu_cpu_affinity(index: int):
    """Bind this rank (its host threads and the pages of its pinned buffers, which are allocated afterwards)
    to the CPUs NVML reports as local to its GPU.  On a single-NUMA virtual machine this is the whole
    machine and changes nothing; on a two-socket host it keeps the copies off the inter-socket link.
    Returns the CPU list that was set, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis else index
        handle = nv.nvmlDeviceGetHandleByIndex(phys)
        words = nv.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:  # noqa: BLE001
        pass
    return None


def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    from mdimg_b200 import synth
    from mdimg_b200.batch import PACK_COLS, default_chunk, tapered_schedule
    from mdimg_b200.shard import process_cohort
    from mdimg_b200.stack import get_ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    affinity = gpu_cpu_affinity(local) if (world > 1 and args.affinity) else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    ops = get_ops(device)
    n = args.slices
    chunk = args.chunk or default_chunk(H, W)
    plan = synth.plan_full()

    stack = make_stack(n, 1000 + 4096 * rank)                 # every rank its own volume (C4 style)
    pinned_in = torch.from_numpy(stack.view(np.int16)).pin_memory()
    pinned_out = torch.empty((n, H, W), dtype=torch.float32, pin_memory=True)
    raw_dev = pinned_in.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        # the product's multi-GPU entry point: this rank's own volume (C4 style, inputs per rank) through the
        # stack pipeline, then ONE all-gather of the device-resident result rows (no host round trip)
        coh = process_cohort([raw_dev], plan, already_sharded=True, chunk=chunk, keep_enhanced=True, ops=ops,
                             workers=args.workers)
        assert coh.rows.shape[0] == world * n
        return coh.local[0][4]

    # several chunks per stack so copies overlap compute; the first copy-in and the last copy-out
    # cannot overlap anything, so the chunks of the end-to-end path are smaller than the resident ones
    e2e_chunk = args.e2e_chunk or max(1, min(chunk, n // 8))
    if args.e2e_schedule:
        e2e_schedule = [int(v) for v in args.e2e_schedule.split(",")]
    elif args.e2e_chunk:
        e2e_schedule = None
    else:
        e2e_schedule = tapered_schedule(n, args.workers)

    pinned_outs = [pinned_out, torch.empty((n, H, W), dtype=torch.float32, pin_memory=True)]

    pinned_outs16 = [torch.empty((n, H, W), dtype=torch.int16, pin_memory=True) for _ in range(2)]

    def cohort_e2e(k, u16=False):
        """k stacks through the public host-buffer API in one call: the chunks of all k stacks form one
        queue, so a stack's last copy-out overlaps the next stack's compute (outputs double-buffered)."""
        coh = process_cohort([stack] * k, plan, already_sharded=True, chunk=e2e_chunk, ops=ops,
                             pinned_ins=[pinned_in] * k,
                             pinned_outs=[(pinned_outs16 if u16 else pinned_outs)[i % 2] for i in range(k)],
                             workers=args.workers, schedule=e2e_schedule,
                             out_dtype=np.uint16 if u16 else np.float32)
        assert coh.rows.shape[0] == world * n * k
        return coh.local[-1][4]

    def timed_e2e(steps, warmup, u16=False):
        for _ in range(warmup):
            # same number of stacks as the timed call: the device and pinned caching allocators reach the timed
            # region's steady state here (a third stack in flight otherwise cudaMallocs inside the timed region)
            cohort_e2e(steps, u16)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cohort_e2e(steps, u16)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    per_step = []
    region = [0.0, 0.0]          # monotonic clock around the last timed region

    def timed(fn, steps, warmup):
        last = None
        for _ in range(warmup):
            # keep the previous result alive while the next step runs, exactly as the timed loop does:
            # the caching allocator then reaches its steady state (two live result stacks) during
            # warm-up instead of cudaMalloc-ing a second 1 GiB block inside the second timed step
            last = fn()
        barrier()
        l0 = ops.lib.mdimg_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        region[0] = time.monotonic()
        e0.record()
        for i in range(steps):
            last = fn()
            marks[i].record()
        e1.record()
        barrier()
        region[1] = time.monotonic()
        ms = e0.elapsed_time(e1)
        prev = e0
        per_step.clear()
        for mk in marks:
            per_step.append(round(prev.elapsed_time(mk), 2))
            prev = mk
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), int(ops.lib.mdimg_launch_count() - l0), last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    ms_total, launches, last = timed(step_resident, args.steps, args.warmup)
    steps_ms = list(per_step)
    clocks = sampler.stop(region[0], region[1]) if rank == 0 else None
    e2e_steps = max(1, args.steps)
    # the end-to-end leg shares the box's host side (pinned copies, worker threads) with everything else running
    # there: it is timed twice, both runs are reported, the faster one is the value
    e2e_runs = [timed_e2e(e2e_steps, max(1, min(args.warmup, 2))), timed_e2e(e2e_steps, 0)]
    ms_e2e = min(e2e_runs)
    ms_e2e_single = timed_e2e(1, 0)          # one stack alone: its last copy-out overlaps nothing
    ms_e2e_u16 = min(timed_e2e(e2e_steps, 1, u16=True), timed_e2e(e2e_steps, 0, u16=True))   # same path, 16-bit export formed on the device (half the D2H bytes); faster of two runs

    px_per_step = float(n) * H * W * world
    value = px_per_step * args.steps / (ms_total / 1e3) / 1e6
    e2e_value = px_per_step * e2e_steps / (ms_e2e / 1e3) / 1e6

    # ---- roofline of the dominant kernel + per-operator list --------------------------------------
    roof = None
    cpu_base = None
    if rank == 0:
        peak, peak_src = measured_hbm_gbs()
        x = ops.normalize(raw_dev[:chunk])
        y = torch.empty_like(x)

        def tv_ms(iters):
            ops.tv_chambolle(x, y, 0.05, eps=0.0, max_iter=iters)     # eps = 0: never stops, no polling
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.tv_chambolle(x, y, 0.05, eps=0.0, max_iter=iters)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b)
        t_long, t_short = tv_ms(41), tv_ms(1)
        bodies_per_launch = 2                             # k_tvp<2>: two Chambolle bodies per launch
        per_launch_ms = (t_long - t_short) / 40.0 * bodies_per_launch
        # Algorithmic bytes of ONE LAUNCH of the kernel as built (a lower bound on its DRAM traffic, so the
        # fraction cannot exceed 1): x read once (4), p read once (8) and written once (8) = 20 B/px per
        # launch, whatever the number of bodies fused inside.  SURVEY 8(d)'s 20 B/px PER BODY describes the
        # unfused iteration; on that basis the same launch moves 40 B/px "algorithmically" (kept below).
        bytes_per_launch = 20.0 * chunk * H * W
        achieved = bytes_per_launch / (per_launch_ms / 1e3) / 1e9
        tv_iters = last.tv_iterations
        traffic = None
        tj = ROOT / "profiles" / "r02_tvp_traffic.json"          # one ncu --set full capture of this kernel (256 slices)
        if not tj.exists():
            tj = ROOT / "profiles" / "r01_tvp_traffic.json"
        if tj.exists():
            try:
                traffic = (float(json.loads(tj.read_text())["dram_bytes_per_pixel_per_body"]) * bodies_per_launch
                           * chunk * H * W)
            except Exception:  # noqa: BLE001
                traffic = None
        roof = {"bound": "hbm", "kernel": "k_tvp<2> (TV-Chambolle, packed two-pixel kernel, 2 bodies per launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per pixel and body "
                                  f"(profiles/{tj.name}: 20.2 B/px per launch measured against 20 algorithmic) "
                                  "x 2 bodies x pixels of this launch",
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "launch_ms": per_launch_ms,
                "unfused_basis": {"bytes_per_px_per_body": 20.0,
                                  "achieved_gbs": 2.0 * achieved, "frac": 2.0 * achieved / peak,
                                  "note": "SURVEY 8(d) counts 20 B/px for every loop body; two bodies share one "
                                          "pass over x and p here, so this figure is not a lower bound on traffic "
                                          "and may exceed 1 -- reported for comparison only"},
                "note": f"one launch = 2 bodies over {chunk} slices x 512x512: x 4 B read + p 8 B read + p 8 B written "
                        f"per pixel; timed as (41 - 1 bodies) / 20 launches with CUDA events on the launching stream; "
                        f"measured limiter: FP32 pipe + issue, DRAM ~55 % busy; "
                        f"mean TV iterations/slice in the workload = {float(tv_iters.mean()):.1f}"}
        del x, y
        roof["operators"] = operator_roofline(ops, raw_dev[:chunk], plan, peak)
        if not args.no_cpu and world == 1:     # the CPU leg is an N=1 measurement (rank 0 would share the host with 7 busy ranks)
            cores = len(os.sched_getaffinity(0))
            sample = max(4 * cores, 32)
            cpu_throughput(stack, cores, cores)          # warm the pool (not timed)
            v, dt = cpu_throughput(stack, sample, cores)
            _cpu_one(stack[0])                           # warm this process
            v1, dt1 = cpu_throughput(stack, 6, 1)        # the single-core rate beside it
            cpu_base = {"value": v, "unit": "Mpx/s", "cores": cores, "kind": "port",
                        "one_core": {"value": v1, "unit": "Mpx/s", "cores": 1,
                                     "sample": f"6 of {n} slices in this process, {dt1:.1f} s wall"},
                        "sample": f"{sample} of {n} slices, one warm process per core, {dt:.1f} s wall; restated "
                                  "reference (numpy+scipy oracle; scikit-image/PyWavelets not installed)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "slices_per_gpu": n, "chunk_slices": chunk, "workers": args.workers,
                       "l2": "input stack (512 MiB u16 per GPU) is larger than L2; no flush needed",
                       "api": "shard.process_cohort (per-rank volume -> stack pipeline -> all-gather of device-resident rows)",
                       "cpu_affinity_rank0": (f"{len(affinity)} CPUs local to the GPU" if affinity else None),
                       "images_per_s": value * 1e6 / (H * W), "ms_each_step_rank0": steps_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpx/s",
                    "h2d_bytes_per_step": int(n * H * W * 2),
                    "d2h_bytes_per_step": int(n * H * W * 4 + n * PACK_COLS * 8),
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "ms_per_step_runs": [round(v / e2e_steps, 3) for v in e2e_runs],
                    "pipelining": "the K stacks go through one process_stacks_host call (one chunk queue, "
                                  "outputs double-buffered): every stack's copies are inside the timed region, "
                                  "a stack's last copy-out overlaps the next stack's compute",
                    "ms_single_stack": ms_e2e_single,
                    "uint16_export": {"value": px_per_step * e2e_steps / (ms_e2e_u16 / 1e3) / 1e6, "unit": "Mpx/s",
                                      "ms_per_step": ms_e2e_u16 / e2e_steps,
                                      "d2h_bytes_per_step": int(n * H * W * 2 + n * PACK_COLS * 8),
                                      "note": "optional extension, not the headline: the reference returns float32; "
                                              "out_dtype=uint16 returns uint16(clip(rint(x * 65535), 0, 65535))"},
                    "chunk_slices": e2e_schedule if e2e_schedule else e2e_chunk},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu_base,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, written to the process's real stdout (see main: while the benchmark runs, file
    descriptor 1 points at stderr, so that nothing a library prints -- NCCL announces its version on
    stdout when the first communicator is created -- can land next to it)."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


def main() -> None:
    global _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slices", type=int, default=1024, help="slices per GPU")
    ap.add_argument("--chunk", type=int, default=0, help="slices per chunk (0 = auto)")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="slices per chunk of the end-to-end leg (0 = auto)")
    ap.add_argument("--e2e-schedule", default="", help="comma-separated chunk sizes of the end-to-end leg")
    ap.add_argument("--workers", type=int, default=4, help="host threads / CUDA streams driving chunks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--affinity", action="store_true",
                    help="bind each rank to the CPUs NVML reports as local to its GPU (N > 1; off by default: the pool's "
                         "boxes are single-NUMA virtual machines where the set is the whole machine)")
    args = ap.parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)
    finally:
        if _cpu_pool is not None:
            _cpu_pool.close()
            _cpu_pool.join()


if __name__ == "__main__":
    main()
