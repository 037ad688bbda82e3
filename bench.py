#!/usr/bin/env python
"""bench.py — scans/s of the voxelise + MinkUNet hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference path's CPU restatement

A step = one pass of the hot path over one batch: coordinate hash + kernel-map build for the batch,
MinkUNet34 (Stage-1 model, ref modules/exp.py:249-267) forward, cross-entropy, backward (dgrad + wgrad),
data-parallel gradient all-reduce, SGD update.  ``value`` times it with inputs resident in HBM;
``e2e`` times the public API from pinned HOST point clouds (H2D copy, GPU quantise + dedup, the step,
D2H of the loss).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")   # scan sizes vary step to step: avoid cudaMalloc churn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import _paths  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (scan kind, scans per GPU, points per scan (None = full sweep), classes, step)
    "kitti_b4": ("kitti", 4, None, 17, "stage1"),          # BASELINE.json configs[1]: the headline line
    "nuscenes_b16": ("nuscenes", 16, None, 14, "stage1"),  # configs[2]
    "stage2": ("kitti", 4, None, 17, "stage2"),            # configs[3]: mean-teacher step, 2 labelled + 2 unlabelled scans per GPU
    "dense": ("dense", 1, None, 17, "stage1"),             # configs[4]: one ~1.2 M-point aggregated scan per GPU (hashing / kernel-map stress)
}
METRIC = "MinkUNet fwd+bwd scans/sec"


def load_traffic(kernel_class):
    """DRAM bytes of one captured launch of the kernel class (profiles/traffic.json, from an ncu --set full capture), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel_class)
    except (OSError, ValueError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t_mark = index, None, [], 0.0

    def mark(self):
        """Samples that arrive from now on belong to the timed region."""
        self.t_mark = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout=5.0):
        """Blocks until the sampler has delivered its first samples: nvidia-smi's start-up (NVML / driver initialisation,
        ~0.5 s) disturbs kernel launches of this process while it lasts and must be over before anything is timed."""
        t0 = time.perf_counter()
        while self.proc is not None and len(self.lines) < 3 and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t_arrived, ln in self.lines:
            if t_arrived < self.t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


class NvmlClockSampler:
    """The same samples through NVML in-process (nvidia_ml_py): a thread reads the SM clock and the clocks-event reasons every
    100 ms (each NVML query takes 10-50 ms of driver time on this box, tools/nvml_probe.py: a faster poll keeps the driver busy
    for no gain).  An `nvidia-smi -lms` child process re-queries the driver from outside on every poll, and every poll that fell into
    the timed region showed up as one step of 25-60 ms among 20 steps of 9.6 ms (r2 call 10: mean 10.8-11.1 ms against a
    median of 9.5-9.7 ms); NVML calls from inside the process do not stall the launch path."""

    def __init__(self, index):
        self.index, self.samples, self.t_mark, self._stop, self.thread, self.max_mhz = index, [], 0.0, False, None, None

    def start(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = self.index
        if vis:
            try:
                phys = int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                phys = self.index
        self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), float(mhz), int(reasons)))
            except Exception:       # noqa: BLE001  (a failed sample is a missing sample)
                pass
            time.sleep(0.1)

    def wait_ready(self, timeout=5.0):
        t0 = time.perf_counter()
        while len(self.samples) < 3 and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        self._stop = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        sm, reasons = [], set()
        for t, mhz, bits in self.samples:
            if t < self.t_mark:
                continue
            sm.append(mhz)
            for name, bit in names:
                if bits & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml"}


def make_clock_sampler(index):
    """NVML in-process when nvidia_ml_py imports and initialises, else the nvidia-smi child process (GCDLSS_BENCH_CLOCKS=smi forces it)."""
    if os.environ.get("GCDLSS_BENCH_CLOCKS", "nvml") != "smi":
        try:
            s = NvmlClockSampler(index)
            s.start()
            return s, True
        except Exception:           # noqa: BLE001
            pass
    return ClockSampler(index), False


# ------------------------------------------------------------------------------------------ data
def make_host_batches(kind, scans_per_gpu, n_points, n_classes, rank, n_batches):
    """Pinned host point clouds [N, 4] (x, y, z, remission) + per-point labels, ``n_batches`` distinct batches."""
    from gcdlss_b200 import synth
    batches = []
    for b in range(n_batches):
        scans = []
        for s in range(scans_per_gpu):
            idx = scan_index(rank, b, s, n_batches, scans_per_gpu)
            xyz, feat = synth.make_dense_scan(idx) if kind == "dense" else synth.make_scan(kind, idx, n_points=n_points)
            pts = torch.from_numpy(np.concatenate([xyz, feat], 1)).pin_memory()
            lab = torch.from_numpy(np.random.default_rng(idx).integers(0, n_classes, xyz.shape[0])).pin_memory()
            scans.append((pts, lab))
        batches.append(scans)
    return batches


def quantize_batch_on_gpu(scans, q, dev):
    """H2D + GPU quantise/dedup of one batch -> (bcoords int32 [M,4], feats [M,1], labels [M]) on the device."""
    from gcdlss_b200.quantize import sparse_quantize_gpu
    coords, feats, labels = [], [], []
    for b, (pts, lab) in enumerate(scans):
        p = pts.to(dev, non_blocking=True)
        l = lab.to(dev, non_blocking=True)
        c, um, _ = sparse_quantize_gpu(p[:, :3], q)
        coords.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32, device=dev), c], 1))
        feats.append(p[:, 3:4].index_select(0, um))
        labels.append(l.index_select(0, um))
    return torch.cat(coords), torch.cat(feats), torch.cat(labels)


# ------------------------------------------------------------------------------------------ reference arm
# Nothing below imports the product package (no gcdlss_b200, no models, no MinkowskiEngine shim, no .so): the scan
# generator is loaded from its file (it needs numpy only) and the parameters come from the oracle's own table of the
# architecture (oracle/minkunet.py:random_params).
def load_synth():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gcdlss_synth", os.path.join(_paths.PKG_DIR, "gcdlss_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scan_index(rank, batch, slot, n_batches, scans_per_gpu):
    """Index (= generator seed offset) of scan ``slot`` of batch ``batch`` on rank ``rank``: both arms use this."""
    return (rank * n_batches + batch) * scans_per_gpu + slot


def oracle_scan_step(synth, kind, index, n_points, n_classes, params, arch):
    """One scan through the CPU restatement: quantise (numpy) + MinkUNet fwd + CE + bwd (torch CPU)."""
    from oracle import quantize as oq
    from oracle.minkunet import OracleMinkUNet
    xyz, feat = synth.make_scan(kind, index, n_points=n_points)
    c, um, _ = oq.sparse_quantize_me(xyz, synth.voxel_size(kind))
    bc = oq.batched_coordinates([c])
    labels = torch.from_numpy(np.random.default_rng(index).integers(0, n_classes, xyz.shape[0])[um])
    for p in params.values():
        if p.is_floating_point() and p.requires_grad:
            p.grad = None
    logits, _, _ = OracleMinkUNet(params, arch, training=True).forward(bc, torch.from_numpy(feat[um]))
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    return float(loss), bc.shape[0]


def time_reference(kind, n_points, n_classes, steps, warmup, scans_per_gpu=4, n_batches=3):
    """Times the CPU oracle on the first scans of the GPU arm's batch rotation (rank 0: batch 0 slot 0, 1, ...), one scan
    per step.  Returns (scans/s, seconds per scan, threads, voxels per scan)."""
    from oracle.minkunet import random_params
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    synth = load_synth()
    params = random_params("MinkUNet34C", 1, n_classes, seed=1234)
    order = [scan_index(0, b, s, n_batches, scans_per_gpu) for b in range(n_batches) for s in range(scans_per_gpu)]
    for i in range(warmup):
        oracle_scan_step(synth, kind, order[-1 - i], n_points, n_classes, params, "MinkUNet34C")
    t0 = time.perf_counter()
    vox = 0
    for i in range(steps):
        vox += oracle_scan_step(synth, kind, order[i % len(order)], n_points, n_classes, params, "MinkUNet34C")[1]
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, cores, vox / max(steps, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, scans, n_points, n_classes = WORKLOADS[args.workload][:4]
    steps, warmup = min(args.steps, 4), min(args.warmup, 1)
    sps, dt, cores, vox = time_reference(kind, n_points, n_classes, steps, warmup, scans)
    sample = (f"{steps} steps of 1 {kind}-like scan each (the first {steps} scans of the GPU arm's rotation; quantise + MinkUNet34C fwd + CE + bwd), "
              f"torch CPU fp32, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": "scans/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "voxels_per_scan": vox, "voxels_per_s": sps * vox,
                       "note": "CPU restatement of the reference path (MinkowskiEngine itself is not installable here); one scan per step"},
            "cpu_baseline": {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "product_modules_imported": sorted(m for m in sys.modules if m.split(".")[0] in ("gcdlss_b200", "MinkowskiEngine", "models"))}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    import gcdlss_b200
    import MinkowskiEngine as ME
    from gcdlss_b200 import ops, synth
    from gcdlss_b200.config import nvtx_range
    from gcdlss_b200.ddp import GradBucketReducer
    from gcdlss_b200.steps import point_cross_entropy
    from models.multiheadminkunet import MinkUNetBase

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    gcdlss_b200.set_math_mode("bf16" if args.dtype == "bf16" else "fp32")
    kind, scans_per_gpu, n_points, n_classes = WORKLOADS[args.workload][:4]
    q = synth.voxel_size(kind)
    peaks = load_peaks()

    # Every rank and every N runs the library's default path (gcdlss_b200.config; GCDLSS_* environment variables override it
    # for A/B runs) and prints it: no run-time self-check, no timing race deciding the code path.
    from gcdlss_b200 import config as gcfg
    path_cfg = {"tile_sort": {"enabled": bool(gcfg.get_tile_sort() and args.dtype == "bf16"), "min_rows": gcfg.tile_sort_min_rows()},
                "kmap_search": gcfg.get_kmap_search()}

    stage = WORKLOADS[args.workload][4]
    torch.manual_seed(1234)
    if stage == "stage2":
        # Stage-2 model pair (ref modules/exp_merge_mean_teacher.py:95-160): MinkUNet34RC student + EMA teacher with the
        # novel-class heads bolted on by the Lightning module
        import copy
        from gcdlss_b200.steps import Stage2Harness, make_stage2_half
        from models.multiheadminkunet import MinkUNetRC
        model = MinkUNetRC(n_classes).to(dev).train()
        for name, n_out in (("final2", 3), ("final3", 2)):
            setattr(model.encoder, name, ME.MinkowskiConvolution(96, n_out, kernel_size=1, bias=True, dimension=3).to(dev))
        teacher = copy.deepcopy(model)
    else:
        model = MinkUNetBase(num_classes=n_classes).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused=True)     # ref modules/exp.py:155-174
    reducer = GradBucketReducer(model.parameters()) if world > 1 else None
    if reducer is not None:
        reducer.broadcast_buffers(model)

    n_batches = 3
    host_batches = make_host_batches(kind, scans_per_gpu, n_points, n_classes, rank, n_batches)
    if stage == "stage2":
        half = scans_per_gpu // 2

        def stage2_batch(b):               # first half of the scans labelled, second half unlabelled
            return (make_stage2_half(kind, 0, half, None, dev, True, n_classes, host_scans=b[:half]),
                    make_stage2_half(kind, 0, scans_per_gpu - half, None, dev, False, n_classes, host_scans=b[half:]))
        resident = [stage2_batch(b) for b in host_batches]
        harness = Stage2Harness(model, teacher, opt, voxel_size=q, reducer=reducer)
        voxels = int(np.mean([r[0]["coords"].shape[0] + r[1]["coords"].shape[0] for r in resident]))
    else:
        resident = [quantize_batch_on_gpu(b, q, dev) for b in host_batches]
        voxels = int(np.mean([r[0].shape[0] for r in resident]))
    torch.cuda.synchronize()
    h2d_bytes = int(np.mean([sum(p.numel() * 4 + l.numel() * 8 for p, l in b) for b in host_batches]))
    points = int(np.mean([sum(p.shape[0] for p, _ in b) for b in host_batches]))

    from gcdlss_b200.prefetch import BatchPrefetcher
    prefetcher = BatchPrefetcher(dev)

    def train_step(st, labels):
        if reducer is None:
            opt.zero_grad(set_to_none=True)        # gradients are taken over from the wgrad kernels' buffers, no accumulate pass
        else:
            reducer.reset()                        # .grad are views of the flat all-reduce buckets: zero them in place
        with nvtx_range("step:forward"):
            out = model(st)
        with nvtx_range("step:loss"):
            loss = point_cross_entropy(out["logits"], labels)
        with nvtx_range("step:backward"):
            loss.backward()
        if reducer is not None:
            with nvtx_range("step:gradient all-reduce (wait)"):
                reducer.finish()
        with nvtx_range("step:optimizer"):
            opt.step()
        return loss

    # Both loops are software-pipelined by one batch: while step i trains on the main stream, batch i+1 is prepared on
    # the prefetcher's side stream (resident: coordinate hash + kernel maps; e2e: also the H2D copy and the GPU
    # quantisation).  Every step still pays for one full preparation, it just overlaps the previous step's GPU tail.
    pending = {}

    def prepare_resident(i):
        bc, f, l = resident[i % n_batches]
        return prefetcher.submit(lambda: (f, bc, l))

    def prepare_e2e(i):
        def make():
            bc, f, l = quantize_batch_on_gpu(host_batches[i % n_batches], q, dev)
            return f, bc, l
        return prefetcher.submit(make)

    def step_resident(i):
        if stage == "stage2":              # the harness builds its two SparseTensors (shared batch, LaserMix batch) itself
            return harness.step(*resident[i % n_batches])
        cur = pending.pop("r", None) or prepare_resident(i)
        pending["r"] = prepare_resident(i + 1)
        st, labels = cur.get()
        return train_step(st, labels)

    # D2H read of every step's result, one step late: the loss of step i is copied to pinned memory asynchronously and read
    # by the host at step i + 1 (the last one by finish_e2e, still inside the timed region), so the host keeps running ahead
    # of the GPU the way a training loop that logs its loss does.
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_log = []

    def read_pending_loss():
        ev = pending.pop("loss_event", None)
        if ev is not None:
            ev.synchronize()
            loss_log.append(float(loss_host[pending.pop("loss_slot")]))

    def step_e2e(i):
        if stage == "stage2":              # H2D of the four scans + GPU quantisation of both halves + the step
            loss = harness.step(*stage2_batch(host_batches[i % n_batches]))
        else:
            cur = pending.pop("e", None) or prepare_e2e(i)
            pending["e"] = prepare_e2e(i + 1)
            st, labels = cur.get()
            loss = train_step(st, labels)
        read_pending_loss()                   # result of the previous step
        slot = i & 1
        loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending["loss_event"], pending["loss_slot"] = ev, slot

    def finish_e2e():
        read_pending_loss()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    max_ahead = int(os.environ.get("GCDLSS_BENCH_MAX_AHEAD", "2"))
    host_t = []

    def timed(fn, steps, finish=None):
        """(total ms, per-step device intervals in ms): K steps between two CUDA events after a barrier + synchronize on both
        sides (max over ranks), plus one event after every step for the step-time percentiles."""
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        e1 = torch.cuda.Event(enable_timing=True)
        host_t.clear()
        marks[0].record()
        for i in range(steps):
            # The host stays at most `max_ahead` steps ahead of the GPU, as a training loop that reads its loss does (the e2e
            # loop reads it one step late).  Unthrottled the host (6.9 ms of launch work per 9.6 ms step) runs further ahead
            # every step until a driver queue fills, and the wait that follows showed up as single steps of 20-180 ms in
            # about half of the runs (r2 calls 10-13, with and without the clock sampler).
            if max_ahead > 0 and i >= max_ahead:
                marks[i - max_ahead + 1].synchronize()
            host_t.append(time.perf_counter())
            fn(i)
            marks[i + 1].record()
        host_t.append(time.perf_counter())
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([marks[0].elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]

    def warm_up(fn, finish=None):
        """At least W steps and every batch shape twice, then windows of 2 * n_batches steps until a window is no more than
        3 % faster than the one before it (at most 4 s): a fresh process / fresh box keeps speeding up for a second or two
        (allocator, clocks, host caches).  All ranks take the same decision (max over ranks).  Used by BOTH timed loops."""
        n = max(args.warmup, 2 * n_batches + 2)
        for i in range(n):
            fn(i)
        prev, spent = None, 0.0
        fixed = bool(os.environ.get("GCDLSS_BENCH_FIXED_WARMUP"))   # profiling runs (ncu --launch-skip needs a fixed launch count)
        while spent < 4.0 and not fixed:                            # `spent` is built from all-reduced times: same on every rank
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(2 * n_batches):
                fn(n + i)
            torch.cuda.synchronize()
            win = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(win, op=dist.ReduceOp.MAX)
            win = float(win.item())
            spent += win
            n += 2 * n_batches
            if prev is not None and win > 0.97 * prev:
                break
            prev = win
        if finish is not None:
            finish()
        return n

    # The sampler (an nvidia-smi process polling every 100 ms) is started BEFORE the warm-up: its start-up (NVML / driver
    # initialisation, ~0.5 s) slows kernel launches of this process while it lasts, which used to fall exactly into the
    # timed region (17 ms instead of 11-13 ms per step); only samples that arrive after mark() are reported.
    clocks, started = (make_clock_sampler(local_rank) if rank == 0 else (ClockSampler(local_rank), True))
    if rank == 0 and not os.environ.get("GCDLSS_BENCH_NO_CLOCKS"):       # (diagnosis only: the sampler is part of the contract)
        if not started:
            clocks.start()
        clocks.wait_ready()
    if world > 1:
        dist.barrier()                    # the other ranks wait for rank 0's sampler too
    n_warm = warm_up(step_resident)
    import gc
    gc.collect()
    gc.freeze()            # model / maps / library objects are long-lived: keep the cyclic GC from re-walking them every few steps
    gc.disable()           # no cyclic collections inside the timed regions (reference counting frees every per-step object; a
                           # generation-2 pass over the frozen heap costs tens of milliseconds); collected between the regions
    if args.debug_steps and rank == 0:
        gc_t = [0.0, 0.0]

        def _gc_cb(phase, info):
            if phase == "start":
                gc_t[1] = time.perf_counter()
            else:
                gc_t[0] += time.perf_counter() - gc_t[1]
        gc.callbacks.append(_gc_cb)
        for i in range(12):
            ms0 = torch.cuda.memory_stats()
            gc_t[0] = 0.0
            torch.cuda.synchronize(); t0 = time.perf_counter()
            step_resident(i)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            ms1 = torch.cuda.memory_stats()
            print(f"debug step {i}: host {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms, gc {1e3 * gc_t[0]:.2f} ms, "
                  f"segments +{ms1['segment.all.allocated'] - ms0['segment.all.allocated']}, reserved {ms1['reserved_bytes.all.current'] >> 20} MiB, "
                  f"alloc_retries {ms1['num_alloc_retries']}", file=sys.stderr)
        gc.callbacks.remove(_gc_cb)
    clocks.mark()
    launches0 = ops.launch_counter["calls"]
    if reducer is not None:
        reducer.measure = True
    total_ms, step_ms = timed(step_resident, args.steps)
    host_step_ms = [1e3 * (b - a) for a, b in zip(host_t, host_t[1:])]
    launches = ops.launch_counter["calls"] - launches0
    comm_exposed_ms = None
    if reducer is not None:                # time the training stream spent waiting for the gradient all-reduce (per step)
        reducer.measure = False
        comm_exposed_ms = sum(a.elapsed_time(b) for a, b in reducer.exposed_events) / max(len(reducer.exposed_events), 1)
    if args.no_e2e:
        e2e_ms, e2e_step_ms = float("nan"), [float("nan")]
    else:
        gc.collect()
        warm_up(step_e2e, finish_e2e)          # the e2e path allocates its own shapes: same warm-up rule as above
        loss_log.clear()
        seg0 = torch.cuda.memory_stats().get("segment.all.allocated", 0)
        e2e_ms, e2e_step_ms = timed(step_e2e, args.steps, finish_e2e)
        e2e_new_segments = torch.cuda.memory_stats().get("segment.all.allocated", 0) - seg0      # cudaMalloc calls inside the timed e2e region
        e2e_host_step_ms = [1e3 * (b - a) for a, b in zip(host_t, host_t[1:])]
        assert len(loss_log) == args.steps and all(np.isfinite(loss_log)), "every timed e2e step must have delivered its loss to the host"
    clock_info = clocks.stop() if rank == 0 else None
    if os.environ.get("GCDLSS_BENCH_PROFILE_STEP") and world == 1:
        # ONE more step between cudaProfilerStart/Stop, outside every timed region: `ncu --profile-from-start off --metrics
        # gpu__time_duration.sum ...` then lists exactly the launches of one steady-state step (no --launch-skip arithmetic)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident(args.steps)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    gc.enable()

    scans_total = scans_per_gpu * world * args.steps
    value = scans_total / (total_ms / 1e3)
    e2e_value = scans_total / (e2e_ms / 1e3)

    # ---- live roofline of the dominant kernel class ----------------------------------------------------------------
    # The timed region issues whole blocks through gcd_block_forward/backward (no per-launch hooks on the host).  For the
    # kernel-only durations one more step runs through the per-launch path (same kernels, same order), its convolution
    # launches are captured and then replayed back to back on the same tensors, all launches of a class between two CUDA
    # events; the share of the step is that kernel time over the measured step time.
    roofline = None
    # the capture step contains the gradient all-reduce, so every rank runs it -- through the SAME (per-launch) path: ranks on
    # different paths complete their all-reduce buckets in different orders, and NCCL collectives issued in different orders
    # hang (seen at 8 GPUs with GCDLSS_DDP_DIRECT=0, r2 call 18).  Only rank 0 keeps the record and replays it.
    from gcdlss_b200 import coords as gcoords
    ops.kernel_timer.captured.clear()
    ops.kernel_timer.capture = True
    gcoords.TABLE_LOG = []                 # every neighbour table built during the capture step (one or two coordinate managers)
    if stage == "stage2":
        harness.step(*resident[0])
    else:
        train_step(ME.SparseTensor(features=resident[0][1], coordinates=resident[0][0]), resident[0][2])
    ops.kernel_timer.capture = False
    tables, gcoords.TABLE_LOG = gcoords.TABLE_LOG, None
    torch.cuda.synchronize()
    if rank != 0:
        ops.kernel_timer.captured.clear()
    if rank == 0:
        pair_count = {}
        for table in tables:                             # every pointer a convolution launch may carry -> pairs of that table
            if table.nbr is not None:
                n_pairs = int(table.pairs[2][-1].item())
                for t in table.device_tensors():
                    pair_count[t.data_ptr()] = n_pairs
        classes = {}
        for kind_k, ptr, n_out, kv, c_in, c_out, fn in ops.kernel_timer.captured:
            pairs = n_out if ptr == 0 else pair_count.get(ptr, n_out)
            cls = classes.setdefault(kind_k, {"flops": 0.0, "fns": [], "launches": 0})
            cls["flops"] += 2.0 * pairs * c_in * c_out
            cls["fns"].append(fn)
            cls["launches"] += 1
        reps = 5
        for cls in classes.values():
            for fn in cls["fns"]:
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for fn in cls["fns"]:
                    fn()
            e1.record()
            torch.cuda.synchronize()
            cls["time"] = e0.elapsed_time(e1) * 1e-3 / reps
        ops.kernel_timer.captured.clear()
        if classes:
            name = max(classes, key=lambda k: classes[k]["time"])
            info = classes[name]
            tensor = name.endswith("_tc")
            peak = peaks["bf16_tflops"] if tensor else 75.0
            achieved = info["flops"] / info["time"] / 1e12
            names = {"conv_tc": "conv_fwd_tc_kernel (forward + dgrad launches)", "wgrad_tc": "conv_wgrad_tc_kernel",
                     "conv_simt": "conv_fwd_simt_kernel", "wgrad_simt": "conv_wgrad_simt_kernel"}
            captured = load_traffic(name)
            roofline = {"kernel": names[name], "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": captured["bytes"] if captured else None,
                        "traffic_of": (f"one launch: {captured['launch']}; compulsory {captured['compulsory_bytes']} B; {captured['source']}"
                                       if captured else None),
                        "peak_source": (peaks["source"] + " bf16_tflops (burst: kernels timed alone, back to back)") if tensor else "nominal fp32 FMA peak",
                        "algorithmic_gflop_per_step": info["flops"] / 1e9, "launches_per_step": info["launches"],
                        "avg_launch_us": info["time"] / info["launches"] * 1e6, "kernel_ms_per_step": info["time"] * 1e3,
                        "share_of_step": {k: v["time"] * 1e3 / (total_ms / args.steps) for k, v in classes.items()},
                        "by_kernel": {k: {"kernel_ms_per_step": v["time"] * 1e3, "tflops": v["flops"] / v["time"] / 1e12,
                                          "launches_per_step": v["launches"]} for k, v in classes.items()}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_timed = 3                                   # bounded sample: ~10 s of CPU work on the box's host cores
        sps, dt, cores, _ = time_reference(kind, n_points, n_classes, n_timed, 1, scans_per_gpu, n_batches)
        cpu_baseline = {"value": sps, "unit": "scans/s", "cores": cores, "kind": "port",
                        "sample": f"1 warm-up + {n_timed} timed {kind}-like scans, one per step (quantise + MinkUNet34C fwd + CE + bwd) through the CPU oracle, torch CPU fp32, {cores} threads"}

    def pct(v):
        return {"p10": float(np.percentile(v, 10)), "p50": float(np.percentile(v, 50)), "p90": float(np.percentile(v, 90)), "max": float(np.max(v))}

    if os.environ.get("GCDLSS_BENCH_DUMP_STEPS") and rank == 0:
        print("step_ms resident:", " ".join(f"{v:.2f}" for v in step_ms), file=sys.stderr)
        print("host_ms resident:", " ".join(f"{v:.2f}" for v in host_step_ms), file=sys.stderr)
        print("step_ms e2e:     ", " ".join(f"{v:.2f}" for v in e2e_step_ms), file=sys.stderr)
        if not args.no_e2e:
            print("host_ms e2e:     ", " ".join(f"{v:.2f}" for v in e2e_host_step_ms), file=sys.stderr)
            print("e2e allocator segments allocated inside the timed region:", e2e_new_segments, file=sys.stderr)

    descr = {"stage1": ("MinkUNet34RC backbone + final head (Stage-1 MinkUNetBase, ref modules/exp.py:249-267)",
                        "hash + kernel maps (side stream, one batch ahead) + fwd + CE + bwd + grad all-reduce + SGD"),
             "stage2": ("MinkUNet34RC student + EMA teacher with final/final2/final3 heads (ref modules/exp_merge_mean_teacher.py:2772-2875)",
                        "teacher fwd + student fwd on the shared 4-scan batch, CE + 200 MSE, pseudo labels -> points, LaserMix, GPU quantise, "
                        "student fwd on the mixed batch, bwd through both, grad all-reduce, SGD, EMA")}[stage]
    if rank == 0:
        total_gflop = sum(v["flops"] for v in classes.values()) / 1e9 if roofline else None
        line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": n_warm,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scans_per_gpu": scans_per_gpu, "points_per_batch": points, "voxels_per_batch": voxels,
                           "voxels_per_s": voxels * world * args.steps / (total_ms / 1e3), "conv_gflop_per_step_per_gpu": total_gflop,
                           "model": descr[0], "classes": n_classes, "step": descr[1], "parallelism": f"dp{world}",
                           "l2": f"{n_batches} distinct batches rotate; per-step activations + maps exceed the 126 MB L2",
                           "host_max_steps_ahead": max_ahead,
                           **path_cfg},
                "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4 + 8 * scans_per_gpu,
                        "ms_per_step": e2e_ms / args.steps, "step_ms": pct(e2e_step_ms)},
                "step_ms": pct(step_ms), "comm_exposed_ms_per_step": comm_exposed_ms,
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clock_info,
                "grad_allreduce_bytes": reducer.grad_bytes() if reducer is not None else 0}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti_b4", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-steps", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
