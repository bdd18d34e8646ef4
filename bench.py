#!/usr/bin/env python
"""bench.py -- throughput of libacgpu's hot path on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric): ac_imgconvert IMG_YUV420P -> IMG_RGB24 at 1920x1080.  A *step* is one
pass of the hot path over one batch of synthetic frames: ONE acgpu_imgconvert_batch launch over
`batch` device-resident frames (batch * 9.33 MB >> the 126 MB L2, so every step streams from HBM).
  value     whole-job frames/s, inputs already resident in HBM, timed with CUDA events on the launching stream
  e2e       the same metric through acgpu_imgconvert_frames_host with pinned HOST buffers: H2D + kernels + D2H inside
            the timed region (PCIe-bound by construction)
  roofline  achieved algorithmic GB/s of the conversion kernel vs the measured HBM copy peak
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores, bounded sample
Multi-GPU (--gpus N under torchrun): frames are sharded, each rank converts its own batch on its own GPU,
no collective on the data path (frames are independent); weak scaling; time = max over ranks.
`--impl reference` times the reference's own CPU path (aclib, stock ac_init(AC_ALL) => SSE2) on the host.
Other workloads (--workload) exist for the remaining BASELINE configs and for profiling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
F = pkg.F

# name -> (kind, srcfmt, dstfmt, w, h, default batch)
WORKLOADS = {
    "yuv420p_rgb24_1080p": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 1920, 1080, 256),      # headline, config 2
    "yuv420p_rgb24_pal": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 720, 576, 1000),          # config 1 size
    "yuv420p_rgb24_uhd": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 3840, 2160, 64),          # config 4, leg 1
    "rgb24_yuv422p_uhd": ("convert", F.IMG_RGB24, F.IMG_YUV422P, 3840, 2160, 64),          # config 4, leg 2
    "yuy2_yuv420p_720p": ("convert", F.IMG_YUY2, F.IMG_YUV420P, 1280, 720, 64 * 8),        # config 5 (8 batches of 64)
    "deinterlace_1080p_y": ("deint", 1, 0, 1920, 1080, 512),                                # config 3 (interpolate, Bpp 1)
    "deinterlace_blend_1080p_rgb": ("deint", 3, 1, 1920, 1080, 192),                        # config 3 (linear blend, Bpp 3)
    "resize_1080_720_y": ("resize", 1, 0, 1920, 1080, 512),                                 # config 3 (1080 -> 720 rows)
}
METRIC = {
    "yuv420p_rgb24_1080p": "1080p frames/s (ac_imgconvert YUV420P->RGB24)",
}


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


class Dist:
    """torch.distributed plumbing for the multi-rank bench: barrier, max / sum over ranks.  The data path has no
    collective (frames are independent); this only synchronises the timed region and reduces the timing scalar.
    backend "nccl" on GPUs; "gloo" is used by the CPU tests (tests/test_bench_dist.py)."""

    def __init__(self, backend: str = "nccl"):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.backend = backend
        self.dist = None
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend)

    def _tensor(self, v, dtype):
        dev = "cuda" if self.backend == "nccl" else "cpu"
        return self.torch.tensor([v], device=dev, dtype=dtype)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
            if self.backend == "nccl":
                self.torch.cuda.synchronize()

    def max_float(self, v: float) -> float:
        if self.dist is None:
            return v
        t = self._tensor(v, self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int(self, v: int) -> int:
        if self.dist is None:
            return v
        t = self._tensor(v, self.torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def aggregate_value(units_all_ranks: int, steps: int, ms_max: float) -> float:
    """Whole-job throughput: units every rank processed / the slowest rank's device time."""
    return units_all_ranks * steps / (ms_max / 1000.0)


def cpubench(lib: str, accel: int, threads: int, seconds: float, kind: str, a: int, b: int, w: int, h: int, oracle=False):
    exe = os.path.join(ROOT, "oracle", "cpubench")
    if kind == "convert":
        args = ["--op", "convert", "--src", hex(a), "--dst", hex(b)]
    elif kind == "deint":
        args = ["--op", "average", "--bpp", str(a)]
    else:
        args = ["--op", "rescale", "--bpp", str(a)]
    cmd = [exe, lib, "--accel", str(accel), "--threads", str(threads), "--seconds", str(seconds), "-w", str(w), "-h", str(h)] + args
    if oracle:
        cmd.insert(2, "--oracle")
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return json.loads(out)


def cpu_libs():
    """(lib path, accel, kind, is_oracle) for the C-path and the stock (SSE2) path."""
    ref_c = os.path.join(ROOT, "oracle", "_ref", "libac_ref_c.so")
    ref_s = os.path.join(ROOT, "oracle", "_ref", "libac_ref_sse2.so")
    if not os.path.exists(os.path.join(ROOT, "oracle", "cpubench")) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=subprocess.DEVNULL)
    if os.path.exists(ref_c) and os.path.exists(ref_s):
        return {"c": (ref_c, 0, "reference", False), "stock": (ref_s, -1, "reference", False)}
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    return {"c": (orc, 0, "port", True), "stock": (orc, 0, "port", True)}


def workload_bytes(kind, a, b, w, h):
    """Algorithmic bytes per frame."""
    if kind == "convert":
        return F.algorithmic_bytes(a, b, w, h)
    # frame-granular row shapes: unique bytes touched by ONE fused pass (source rows read once + rows written),
    # not aclib's per-call 3 B per blended byte -- the fused kernel gets the row reuse from cache.
    bpl = w * a
    if kind == "deint":
        if b == 0:   # interpolate reads only the even rows (tcvideo.c:353-364)
            return bpl * ((h + 1) // 2 + h)
        return bpl * (h + h)                 # linear blend reads every row once
    new_h = h * 2 // 3
    return bpl * (h + new_h)                 # 3:2 shrink touches every source row


def run_reference(args, kind, a, b, w, h, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    libs = cpu_libs()
    lib, accel, kd, is_o = libs["stock"]
    cores = os.cpu_count() or 1
    # a step = a bounded slice of frame-parallel conversion; the whole run stays near two minutes whatever K and W are
    per_step = max(0.2, min(1.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpubench(lib, accel, cores, per_step, kind, a, b, w, h, oracle=is_o)
        if i >= args.warmup:
            vals.append(r["frames_per_s"])
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC.get(name, name + " frames/s"), "value": round(v, 2), "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "width": w, "height": h, "path": "aclib stock path ac_init(AC_ALL) => SSE2 asm" if accel else "aclib C path",
                   "threads": cores, "step": f"{per_step:.1f} s of frame-parallel conversion on all host cores"},
        "cpu_baseline": {"value": round(v, 2), "unit": "frames/s", "cores": cores, "kind": kd, "cpu_model": cpu_model(),
                         "sample": f"{args.steps} x {per_step:.1f} s, {cores} pthreads, each cycling through 4 distinct frames"},
        "e2e": {"value": round(v, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="acgpu", choices=["acgpu", "reference"])
    ap.add_argument("--workload", default="yuv420p_rgb24_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (0 = workload default)")
    ap.add_argument("--tier", type=int, default=0, help="force a kernel tier (profiling)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"], help="process-group backend (gloo: CPU tests)")
    ap.add_argument("--dry-run", action="store_true",
                    help="exercise the multi-rank plumbing with no GPU work (tests/test_bench_dist.py); not a measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "acgpu":
        args.warmup = 3
    name = args.workload
    kind, a, b, w, h, batch = WORKLOADS[name]
    if args.batch:
        batch = args.batch

    if args.impl == "reference":
        run_reference(args, kind, a, b, w, h, name)
        return

    dc = Dist(args.backend)
    rank, local_rank, world = dc.rank, dc.local_rank, dc.world
    abytes = workload_bytes(kind, a, b, w, h)

    if args.dry_run:
        # no device: each rank pretends its steps took (rank+1) ms so the max-over-ranks path is observable
        dc.barrier()
        ms = dc.max_float(float(rank + 1) * args.steps)
        total_units = dc.sum_int(batch)
        dc.barrier()
        if rank == 0:
            print(json.dumps({"dry_run": True, "n_gpus": world, "steps": args.steps, "ms_max": ms,
                              "units_all_ranks": total_units, "value": aggregate_value(total_units, args.steps, ms)}), flush=True)
        dc.close()
        return

    import numpy as np
    ac = pkg.AcGpu()
    if not ac.lib.acgpu_set_device(local_rank):
        raise SystemExit("libacgpu: " + ac.last_error())
    if ac.ac_init(pkg.AC_CUDA) != 1:
        raise SystemExit("libacgpu: ac_init(AC_CUDA) failed: " + ac.last_error() + " (no CPU fallback)")
    ac.lib.acgpu_force_tier(args.tier)
    lib = ac.lib

    # ---- synthetic, device-resident input -------------------------------------------------------------
    if kind == "convert":
        sfb, dfb = F.frame_bytes(a, w, h), F.frame_bytes(b, w, h)
    elif kind == "deint":
        sfb = dfb = w * h * a
    else:
        sfb, dfb = w * h * a, w * (h * 2 // 3) * a
    rng = np.random.default_rng(1234 + rank)
    uniq = min(batch, 8)
    host = rng.integers(0, 256, size=(uniq, sfb), dtype=np.uint8)
    dsrc = ac.malloc(batch * sfb + w * 4)
    ddst = ac.malloc(batch * dfb)
    for i in range(batch):
        lib.acgpu_memcpy_h2d(dsrc.ptr + i * sfb, host[i % uniq].ctypes.data, sfb, None)
    ac.sync()
    ddst.fill(0)
    stream = lib.acgpu_stream_create()

    def step():
        if kind == "convert":
            ok = ac.imgconvert_batch(dsrc.ptr, a, sfb, ddst.ptr, b, dfb, w, h, batch, stream)
        elif kind == "deint":
            ok = lib.acgpu_deinterlace_batch(dsrc.ptr, ddst.ptr, w, h, a, b, sfb, dfb, batch, stream)
        else:
            ok = lib.acgpu_resize_batch(dsrc.ptr, ddst.ptr, w, h, a, 0, -(h // 3) // 8, 8, 8, sfb, dfb, batch, stream)
        if ok != 1:
            raise SystemExit("libacgpu: " + ac.last_error())

    def barrier():
        ac.sync(stream)
        dc.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = lib.acgpu_event_create(), lib.acgpu_event_create()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.acgpu_launch_count(1)
    # ---- timed region: exactly K steps -----------------------------------------------------------------
    lib.acgpu_event_record(e0, stream)
    for _ in range(args.steps):
        step()
    lib.acgpu_event_record(e1, stream)
    ac.sync(stream)
    launches = int(lib.acgpu_launch_count(0))
    ms_local = float(lib.acgpu_event_elapsed_ms(e0, e1))
    # keep the GPU busy a little longer if the region was too short for the 100 ms clock sampler
    t_end = time.time() + max(0.0, 0.5 - ms_local / 1000.0)
    while time.time() < t_end:
        step()
    ac.sync(stream)
    sampler.stop_flag = True
    sampler.join()
    barrier()
    tier = lib.acgpu_last_kernel_tier()
    ms = dc.max_float(ms_local)
    total_units = dc.sum_int(batch)

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------------
    e2e = None
    if kind == "convert" and not args.no_e2e:
        eb = min(batch, 256)
        hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
        for i in range(eb):
            hs.array[i * sfb:(i + 1) * sfb] = host[i % uniq]
        for _ in range(2):
            ac._ok(lib.acgpu_imgconvert_frames_host(hs.ptr, a, hd.ptr, b, w, h, eb))
        dc.barrier()
        esteps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(esteps):
            ac._ok(lib.acgpu_imgconvert_frames_host(hs.ptr, a, hd.ptr, b, w, h, eb))
        dt = dc.max_float(time.perf_counter() - t0)
        # PCIe ceiling of this box, measured the same way: the larger direction alone, pinned, one big copy
        big_dir_bytes = max(sfb, dfb) * eb
        probe = ac.malloc(big_dir_bytes)
        cp = lib.acgpu_memcpy_d2h if dfb >= sfb else lib.acgpu_memcpy_h2d
        hp = hd if dfb >= sfb else hs
        args_cp = (hp.ptr, probe.ptr) if dfb >= sfb else (probe.ptr, hp.ptr)
        cp(args_cp[0], args_cp[1], big_dir_bytes, None); ac.sync()
        t1 = time.perf_counter()
        for _ in range(3):
            cp(args_cp[0], args_cp[1], big_dir_bytes, None)
        ac.sync()
        pcie = 3 * big_dir_bytes / (time.perf_counter() - t1) / 1e9
        probe.free()
        e2e_gbs = eb * esteps * max(sfb, dfb) / dt / 1e9        # this rank's dominant direction
        e2e = {"value": round(world * eb * esteps / dt, 2), "unit": "frames/s", "h2d_bytes_per_step": eb * sfb,
               "d2h_bytes_per_step": eb * dfb, "frames_per_step": eb, "steps": esteps,
               "bound": "pcie", "pcie_dominant_direction": "d2h" if dfb >= sfb else "h2d",
               "pcie_achieved_gbs": round(e2e_gbs, 2), "pcie_peak_gbs": round(pcie, 2), "pcie_frac": round(e2e_gbs / pcie, 4),
               "note": "acgpu_imgconvert_frames_host on pinned host buffers (3-slot H2D/kernel/D2H pipeline); wall clock "
                       "around the synchronous call, max over ranks; pcie_peak = the dominant direction alone, measured here"}
        hs.free(); hd.free()

    if rank != 0:
        dc.close()
        return

    peak, peak_src = read_peaks()
    launches_per_step = max(1, launches // max(1, args.steps))
    gbs = batch * abytes * args.steps / (ms_local / 1000.0) / 1e9            # rank 0's kernel(s)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(name)
        if t:
            traffic = int(t["dram_bytes_per_launch"] * batch / t["frames_per_launch"] / launches_per_step)
    except Exception:
        pass
    line = {
        "metric": METRIC.get(name, name + " frames/s"),
        "value": round(aggregate_value(total_units, args.steps, ms), 1), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "width": w, "height": h, "frames_per_step_per_gpu": batch,
                   "bytes_per_frame_algorithmic": abytes, "kernel_tier": tier,
                   "l2": "inputs larger than L2: %.0f MB touched per step" % (batch * abytes / 1e6),
                   "sharding": "frames sharded across ranks, no collective (frames are independent)"},
        "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                     "traffic": traffic, "peak_source": peak_src, "launches_per_step": launches_per_step,
                     "algorithmic_bytes_per_launch": batch * abytes // launches_per_step},
        "gpu_launches": launches,
        "clocks": sampler.result(),
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        libs = cpu_libs()
        cores = os.cpu_count() or 1
        res = {}
        for key, secs, thr in (("stock", 5.0, cores), ("c", 5.0, cores), ("stock", 2.5, 1), ("c", 2.5, 1)):
            lb, accel, kd, is_o = libs[key]
            r = cpubench(lb, accel, thr, secs, kind, a, b, w, h, oracle=is_o)
            res[(key, thr)] = r["frames_per_s"]
        line["cpu_baseline"] = {
            "value": round(res[("stock", cores)], 2), "unit": "frames/s", "cores": cores, "kind": libs["stock"][2],
            "sample": f"5 s per variant on {cores} pthreads + 2.5 s on 1 thread, same frame size, each thread cycling through 4 distinct frames",
            "path": "aclib stock ac_init(AC_ALL) => SSE2",
            "c_path_all_cores": round(res[("c", cores)], 2), "sse2_1_thread": round(res[("stock", 1)], 2),
            "c_path_1_thread": round(res[("c", 1)], 2), "cpu_model": cpu_model(),
        }
    print(json.dumps(line), flush=True)
    dc.close()


if __name__ == "__main__":
    main()
