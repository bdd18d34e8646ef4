#!/usr/bin/env python
"""bench.py -- throughput of libacgpu's hot path on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric): ac_imgconvert IMG_YUV420P -> IMG_RGB24 at 1920x1080.  A *step* is one
pass of the hot path over one batch of synthetic frames: ONE acgpu_imgconvert_batch launch over
`batch` device-resident frames (batch * 9.33 MB >> the 126 MB L2, so every step streams from HBM).
  value     whole-job frames/s, inputs already resident in HBM, timed with CUDA events on the launching stream
  e2e       the same metric through acgpu_imgconvert_frames_host with pinned HOST buffers: H2D + kernels + D2H inside
            the timed region (PCIe-bound by construction); its ceiling is measured in the same run, both copy
            directions concurrent, all ranks at once behind a barrier
  roofline  achieved algorithmic GB/s of the conversion kernel vs the measured HBM copy peak
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores, bounded sample
  extra     the other BASELINE configs, each with the same keys (device-resident value, roofline, e2e and -- at N=1 -- the
            reference on the host cores): config 4 as ONE resident chain (3840x2160 YUV420P -> RGB24 -> YUV422P, one
            stream per GPU: acgpu_chain_*), config 5 (1280x720 YUY2 -> YUV420P, batches of 64), config 3's row shapes, and a
            do_process_frame-shaped chain (-I 5 -B 45,80 -G 0.8: 1080p YUV420P -> deinterlaced, gamma-corrected 720p)
Multi-GPU (--gpus N under torchrun): frames are sharded, each rank converts its own batch on its own GPU,
no collective on the data path (frames are independent); weak scaling; time = max over ranks.
`--impl reference` times the reference's own CPU path (aclib, stock ac_init(AC_ALL) => SSE2) on the host, and carries the
same `extra` workloads timed the same way.
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

# name -> (kind, a, b, w, h, default batch)
#   convert: a = srcfmt, b = dstfmt        chain: a = srcfmt, b = [(kind, args...)]        deint / resize: a = Bpp, b = mode
WORKLOADS = {
    "yuv420p_rgb24_1080p": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 1920, 1080, 256),      # headline, config 2
    "yuv420p_rgb24_pal": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 720, 576, 1000),          # config 1 size
    "yuv420p_rgb24_uhd": ("convert", F.IMG_YUV420P, F.IMG_RGB24, 3840, 2160, 64),          # config 4, leg 1
    "rgb24_yuv422p_uhd": ("convert", F.IMG_RGB24, F.IMG_YUV422P, 3840, 2160, 64),          # config 4, leg 2
    "uhd_roundtrip": ("chain", F.IMG_YUV420P, [(pkg.CHAIN_CONVERT, F.IMG_RGB24), (pkg.CHAIN_CONVERT, F.IMG_YUV422P)],
                      3840, 2160, 64),                                                      # config 4 as one resident chain
    "yuy2_yuv420p_720p": ("convert", F.IMG_YUY2, F.IMG_YUV420P, 1280, 720, 64 * 8),        # config 5 (8 batches of 64)
    "deinterlace_1080p_y": ("deint", 1, 0, 1920, 1080, 512),                                # config 3 (interpolate, Bpp 1)
    "deinterlace_blend_1080p_rgb": ("deint", 3, 1, 1920, 1080, 192),                        # config 3 (linear blend, Bpp 3)
    "resize_1080_720_y": ("resize", 1, 0, 1920, 1080, 512),                                 # config 3 (1080 -> 720 rows)
    "process_frame_1080p": ("chain", F.IMG_YUV420P,                                         # a do_process_frame-shaped chain
                            [(pkg.CHAIN_DEINTERLACE, 5), (pkg.CHAIN_RESIZE, -80, -45), (pkg.CHAIN_GAMMA, 0.8)],
                            1920, 1080, 256),                                               # -I 5 -B 45,80 -G 0.8: 1080p -> 720p
}
METRIC = {
    "yuv420p_rgb24_1080p": "1080p frames/s (ac_imgconvert YUV420P->RGB24)",
}
EXTRA_DEFAULT = ["uhd_roundtrip", "yuy2_yuv420p_720p", "deinterlace_blend_1080p_rgb", "resize_1080_720_y", "process_frame_1080p"]


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


# ---- the reference on the host cores ------------------------------------------------------------------------------------
CHAIN_SPEC = {1: "convert", 2: "clip", 3: "deint", 4: "resize", 5: "reduce", 6: "flipv", 7: "fliph", 10: "gamma", 11: "antialias"}


def chain_spec(stages):
    """stage tuples -> cpubench's --chain syntax (the reference side of a chain: libtcvideo + aclib calls, per plane)."""
    out = []
    for st in stages:
        if st[0] not in CHAIN_SPEC:
            return None              # -k / -K have no libtcvideo entry point of their own
        args = [hex(v) if st[0] == 1 else repr(v) for v in st[1:]]
        out.append(":".join([CHAIN_SPEC[st[0]]] + args))
    return ",".join(out)


def as_chain(kind, a, b, w, h):
    """Every workload as (source format, stage list) -- how the reference libtcvideo is driven for it."""
    if kind == "chain":
        return a, b
    if kind == "deint":
        return (F.IMG_RGB24 if a == 3 else F.IMG_Y8), [(pkg.CHAIN_DEINTERLACE, 5 if b else 1)]
    if kind == "resize":
        return (F.IMG_RGB24 if a == 3 else F.IMG_Y8), [(pkg.CHAIN_RESIZE, 0, -(h // 3) // 8)]
    return None


def cpubench(libs, key: str, threads: int, seconds: float, kind: str, a, b, w: int, h: int):
    """Times the reference on the host cores; `key` picks the plain-C or the stock (SSE2) build."""
    exe = os.path.join(ROOT, "oracle", "cpubench")
    lib, accel, _, oracle = libs[key]
    if kind == "convert":
        args = ["--op", "convert", "--src", hex(a), "--dst", hex(b)]
    else:
        tcv = libs.get("tcv_" + key)
        fmt, stages = as_chain(kind, a, b, w, h)
        spec = chain_spec(stages)
        if tcv and spec:              # whole frames through the reference's own libtcvideo functions
            lib = tcv
            args = ["--op", "chain", "--src", hex(fmt), "--chain", spec]
        elif kind == "deint" and not b:
            args = ["--op", "average", "--bpp", str(a)]
        elif kind == "resize":
            args = ["--op", "rescale", "--bpp", str(a)]
        else:
            return None
    cmd = [exe, lib, "--accel", str(accel), "--threads", str(threads), "--seconds", str(seconds), "-w", str(w), "-h", str(h)] + args
    if oracle:
        cmd.insert(2, "--oracle")
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    return json.loads(out)


def cpu_libs():
    """key -> (lib path, accel, kind, is_oracle) for the plain-C path ("c") and the stock SSE2 path ("stock");
    "tcv_c" / "tcv_stock" -> the reference libtcvideo built over each, when present."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    ref_c, ref_s = os.path.join(ref, "libac_ref_c.so"), os.path.join(ref, "libac_ref_sse2.so")
    if not os.path.exists(os.path.join(ROOT, "oracle", "cpubench")) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=subprocess.DEVNULL)
    if os.path.exists(ref_c) and os.path.exists(ref_s):
        libs = {"c": (ref_c, 0, "reference", False), "stock": (ref_s, -1, "reference", False)}
        for key, name in (("tcv_c", "libtcv_ref.so"), ("tcv_stock", "libtcv_ref_sse2.so")):
            if os.path.exists(os.path.join(ref, name)):
                libs[key] = os.path.join(ref, name)
        return libs
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    return {"c": (orc, 0, "port", True), "stock": (orc, 0, "port", True)}


def host_bytes(kind, a, b, w, h):
    """(bytes in, bytes out) per frame that cross the library boundary -- what an end-to-end caller moves."""
    if kind == "convert":
        return F.frame_bytes(a, w, h), F.frame_bytes(b, w, h)
    if kind == "chain":
        ac = _AC[0]
        import ctypes as C
        ops = pkg.chain_ops(b)
        of, ow, oh = C.c_int(0), C.c_int(0), C.c_int(0)
        if ac.lib.acgpu_chain_output(a, w, h, ops, len(b), C.byref(of), C.byref(ow), C.byref(oh)) != 1:
            raise SystemExit("libacgpu: " + ac.last_error())
        return F.frame_bytes(a, w, h), F.frame_bytes(of.value, ow.value, oh.value)
    if kind == "deint":
        return w * h * a, w * h * a
    return w * h * a, w * (h * 2 // 3) * a


def workload_bytes(kind, a, b, w, h):
    """Algorithmic bytes per frame."""
    if kind == "convert":
        return F.algorithmic_bytes(a, b, w, h), "bytes read + written by one fused pass"
    if kind == "chain":
        # every conversion of the chain counted as its own pass (37 324 800 + 41 472 000 B); the row / plane
        # stages by what they read and write.  Intermediates that stay in L2 make the DRAM traffic smaller than this.
        total, fmt, cw, ch = 0, a, w, h
        for st in b:
            if st[0] == pkg.CHAIN_CONVERT:
                total += F.algorithmic_bytes(fmt, st[1], cw, ch)
                fmt = st[1]
            elif st[0] == pkg.CHAIN_RESIZE:
                fb = F.frame_bytes(fmt, cw, ch)
                nw, nh = cw + st[1] * 8, ch + st[2] * 8
                mid = F.frame_bytes(fmt, cw, nh)
                total += (fb + mid if st[2] else 0) + (mid + F.frame_bytes(fmt, nw, nh) if st[1] else 0)
                cw, ch = nw, nh
            elif st[0] in (pkg.CHAIN_DEINTERLACE, pkg.CHAIN_GAMMA, pkg.CHAIN_ANTIALIAS):
                first = cw * ch * (3 if fmt == F.IMG_RGB24 else 1)
                total += 2 * first if st[0] == pkg.CHAIN_GAMMA else 2 * F.frame_bytes(fmt, cw, ch)
            else:
                total += 2 * F.frame_bytes(fmt, cw, ch)
        return total, "sum over the chain's stages of bytes read + written (each conversion its own pass)"
    # row shapes, SURVEY 8d: 3 B per blended output byte, 2 B per copied output byte -- what aclib's per-row calls move.
    bpl = w * a
    if kind == "deint":
        if b == 0:   # interpolate: even rows copied, odd rows blended, an odd last row copied (tcvideo.c:353-364)
            blended = (h - 1) // 2 if h % 2 else h // 2 - 1
            return bpl * (3 * blended + 2 * (h - blended)), "3 B per blended byte + 2 B per copied byte (SURVEY 8d)"
        # linear blend: interpolate (above) + 539 in-place row averages + copies + one whole-frame average (tcvideo.c:368-389)
        blended = h // 2 - 1
        return bpl * (3 * blended + 2 * (h - blended)) + bpl * (3 * ((h - 1) // 2) + 2 * 2) + 3 * bpl * h, \
            "aclib's three passes: 3 B per blended byte + 2 B per copied byte (SURVEY 8d); the fused kernel moves 2 B per byte"
    new_h = h * 2 // 3
    return bpl * 3 * new_h, "3 B per blended output byte (SURVEY 8d): every row of the 3:2 shrink is a two-row blend"


_AC = [None]


def measure(ac, dc, args, name, steps, warmup, do_e2e=True, do_cpu=True, clocks=True):
    """One workload through libacgpu on this rank's GPU; returns the fields of a bench line (rank 0 fills the aggregate)."""
    import numpy as np
    lib = ac.lib
    kind, a, b, w, h, batch = WORKLOADS[name]
    if args.batch and name == args.workload:
        batch = args.batch
    rank, local_rank, world = dc.rank, dc.local_rank, dc.world
    abytes, abytes_note = workload_bytes(kind, a, b, w, h)
    sfb, dfb = host_bytes(kind, a, b, w, h)
    ops = pkg.chain_ops(b) if kind == "chain" else None

    # ---- synthetic, device-resident input -------------------------------------------------------------
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
        elif kind == "chain":
            ok = lib.acgpu_chain_batch(dsrc.ptr, a, w, h, sfb, ddst.ptr, dfb, ops, len(b), batch, stream)
        elif kind == "deint":
            ok = lib.acgpu_deinterlace_batch(dsrc.ptr, ddst.ptr, w, h, a, b, sfb, dfb, batch, stream)
        else:
            ok = lib.acgpu_resize_batch(dsrc.ptr, ddst.ptr, w, h, a, 0, -(h // 3) // 8, 8, 8, sfb, dfb, batch, stream)
        if ok != 1:
            raise SystemExit("libacgpu: " + ac.last_error())

    def barrier():
        ac.sync(stream)
        dc.barrier()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = lib.acgpu_event_create(), lib.acgpu_event_create()
    sampler = ClockSampler(local_rank) if clocks else None
    if sampler:
        sampler.start()
    lib.acgpu_launch_count(1)
    # ---- timed region: exactly K steps -----------------------------------------------------------------
    lib.acgpu_event_record(e0, stream)
    for _ in range(steps):
        step()
    lib.acgpu_event_record(e1, stream)
    ac.sync(stream)
    launches = int(lib.acgpu_launch_count(0))
    ms_local = float(lib.acgpu_event_elapsed_ms(e0, e1))
    if sampler:
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
    lib.acgpu_event_destroy(e0); lib.acgpu_event_destroy(e1)
    dsrc.free(); ddst.free()

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------------
    e2e = None
    if do_e2e and not args.no_e2e:
        budget = 1 << 30                                        # pinned bytes per direction at most
        eb = max(1, min(batch, 256, budget // max(sfb, dfb)))
        hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
        for i in range(eb):
            hs.array[i * sfb:(i + 1) * sfb] = host[i % uniq]

        def call():
            if kind == "convert":
                ac._ok(lib.acgpu_imgconvert_frames_host(hs.ptr, a, hd.ptr, b, w, h, eb))
            elif kind == "chain":
                ac._ok(lib.acgpu_chain_frames_host(hs.ptr, a, w, h, hd.ptr, ops, len(b), eb))
            else:       # the row shapes as one-stage chains over whole planes: same kernels, pipelined upload / download
                fmt1 = F.IMG_RGB24 if a == 3 else F.IMG_Y8
                st = [(pkg.CHAIN_DEINTERLACE, 5 if b else 1)] if kind == "deint" else [(pkg.CHAIN_RESIZE, 0, -(h // 3) // 8)]
                ac._ok(lib.acgpu_chain_frames_host(hs.ptr, fmt1, w, h, hd.ptr, pkg.chain_ops(st), 1, eb))

        for _ in range(2):
            call()
        dc.barrier()
        esteps = max(3, min(steps, 10))
        t0 = time.perf_counter()
        for _ in range(esteps):
            call()
        dt = dc.max_float(time.perf_counter() - t0)
        # The ceiling of this box for this byte mix, measured the same way and at the same moment on every rank: the upload
        # of `eb` source frames and the download of `eb` result frames as two concurrent big pinned copies.
        pin, pout = ac.malloc(eb * sfb), ac.malloc(eb * dfb)
        s_in, s_out = lib.acgpu_stream_create(), lib.acgpu_stream_create()

        def copies(do_in, do_out, reps=3):
            dc.barrier()
            t1 = time.perf_counter()
            for _ in range(reps):
                if do_in:
                    lib.acgpu_memcpy_h2d(pin.ptr, hs.ptr, eb * sfb, s_in)
                if do_out:
                    lib.acgpu_memcpy_d2h(hd.ptr, pout.ptr, eb * dfb, s_out)
            ac.sync(s_in); ac.sync(s_out)
            return dc.max_float(time.perf_counter() - t1) / reps

        copies(True, True, 1)
        t_both, t_in, t_out = copies(True, True), copies(True, False), copies(False, True)
        lib.acgpu_stream_destroy(s_in); lib.acgpu_stream_destroy(s_out)
        pin.free(); pout.free()
        val = world * eb * esteps / dt
        ceil_fps = world * eb / t_both
        e2e = {"value": round(val, 2), "unit": "frames/s", "h2d_bytes_per_step": eb * sfb,
               "d2h_bytes_per_step": eb * dfb, "frames_per_step": eb, "steps": esteps,
               "bound": "pcie + host memory",
               "host_bytes_per_s": round(val * (sfb + dfb)),
               "ceiling": {"frames_per_s": round(ceil_fps, 2), "frac": round(val / ceil_fps, 4),
                           "h2d_gbs_alone": round(eb * sfb / t_in / 1e9, 2), "d2h_gbs_alone": round(eb * dfb / t_out / 1e9, 2),
                           "h2d_gbs_concurrent": round(eb * sfb / t_both / 1e9, 2), "d2h_gbs_concurrent": round(eb * dfb / t_both / 1e9, 2),
                           "how": "per rank: the same source and result bytes as two concurrent pinned copies (H2D stream + D2H "
                                  "stream), every rank at once behind a barrier, max over ranks"},
               "note": "through the host-buffer C-ABI call on pinned host buffers (4-slot upload / kernels / download pipeline); wall "
                       "clock around the synchronous call, max over ranks"}
        hs.free(); hd.free()
    lib.acgpu_stream_destroy(stream)

    if rank != 0:
        return None
    peak, peak_src = read_peaks()
    launches_per_step = max(1, launches // max(1, steps))
    gbs = batch * abytes * steps / (ms_local / 1000.0) / 1e9            # rank 0's kernel(s)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(name)
        if t:
            traffic = int(t["dram_bytes_per_launch"] * batch / t["frames_per_launch"] / launches_per_step)
            traffic_src = "static: profiles/ncu_traffic.json (one ncu --set full capture of this kernel), scaled to this batch"
    except Exception:
        pass
    res = {
        "metric": METRIC.get(name, name + " frames/s"),
        "value": round(aggregate_value(total_units, steps, ms), 1), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "width": w, "height": h, "frames_per_step_per_gpu": batch,
                   "bytes_per_frame_algorithmic": abytes, "bytes_per_frame_algorithmic_rule": abytes_note, "kernel_tier": tier,
                   "l2": "inputs larger than L2: %.0f MB of frames per step" % (batch * (sfb + dfb) / 1e6),
                   "sharding": "frames sharded across ranks, no collective (frames are independent)"},
        "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches_per_step": launches_per_step,
                     "algorithmic_bytes_per_launch": batch * abytes // launches_per_step},
        "gpu_launches": launches,
    }
    if kind == "chain":
        # A chain's roofline is stated in UNIQUE bytes (source read once + result written once): conversion pairs are fused
        # and what is not fused keeps its intermediates on the device, so the sum of the stages' own traffic (SURVEY 8d counts
        # config 4 as 37.3 + 41.5 MB per frame) is what the CPU pays, not a bound for the device.  Both are reported.
        uniq_gbs = batch * (sfb + dfb) * steps / (ms_local / 1000.0) / 1e9
        res["config"]["stages"] = len(b)
        res["config"]["bytes_per_frame_unique"] = sfb + dfb
        res["roofline"].update({"achieved": round(uniq_gbs, 1), "frac": round(uniq_gbs / peak, 4),
                                "algorithmic_bytes_per_launch": batch * (sfb + dfb) // launches_per_step,
                                "sum_of_stage_passes_gbs": round(gbs, 1), "sum_of_stage_passes_frac": round(gbs / peak, 4),
                                "note": "achieved / frac count the source read once and the result written once; "
                                        "sum_of_stage_passes counts every stage as its own pass over HBM (SURVEY 8d), which a fused "
                                        "or resident chain does not make"})
    if sampler:
        res["clocks"] = sampler.result()
    if e2e:
        res["e2e"] = e2e
    if world == 1 and do_cpu and not args.no_cpu:
        libs = cpu_libs()
        cores = os.cpu_count() or 1
        r = {}
        plan = (("stock", 5.0, cores), ("c", 5.0, cores), ("stock", 2.5, 1), ("c", 2.5, 1)) if name == args.workload \
            else (("stock", 3.0, cores), ("c", 3.0, cores))
        for key, secs, thr in plan:
            out = cpubench(libs, key, thr, secs, kind, a, b, w, h)
            if out is None:
                r = None
                break
            r[(key, thr)] = out["frames_per_s"]
        if r:
            res["cpu_baseline"] = {
                "value": round(r[("stock", cores)], 2), "unit": "frames/s", "cores": cores, "kind": libs["stock"][2],
                "sample": f"{plan[0][1]:.0f} s per variant on {cores} pthreads, same frame size, each thread cycling through 4 distinct frames; "
                          "the clock starts after every thread has filled its buffers",
                "path": "aclib stock ac_init(AC_ALL) => SSE2" + ("" if kind == "convert" else ", driven per frame by the reference libtcvideo"),
                "c_path_all_cores": round(r[("c", cores)], 2),
                "host_bytes_per_s": round(r[("stock", cores)] * (sfb + dfb)), "cpu_model": cpu_model(),
            }
            if ("stock", 1) in r:
                res["cpu_baseline"]["sse2_1_thread"] = round(r[("stock", 1)], 2)
                res["cpu_baseline"]["c_path_1_thread"] = round(r[("c", 1)], 2)
    return res


def run_reference(args, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    libs = cpu_libs()
    lib, accel, kd, is_o = libs["stock"]
    cores = os.cpu_count() or 1

    def one(nm, steps, warmup, per_step):
        kind, a, b, w, h, _ = WORKLOADS[nm]
        vals = []
        for i in range(warmup + steps):
            r = cpubench(libs, "stock", cores, per_step, kind, a, b, w, h)
            if r is None:
                return None
            if i >= warmup:
                vals.append(r["frames_per_s"])
        v = sum(vals) / len(vals)
        return {
            "impl": "reference", "metric": METRIC.get(nm, nm + " frames/s"), "value": round(v, 2), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 * per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": nm, "width": w, "height": h, "path": ("aclib stock path ac_init(AC_ALL) => SSE2 asm" if accel else "aclib C path")
                       + ("" if kind == "convert" else ", whole frames through the reference libtcvideo"),
                       "threads": cores, "step": f"{per_step:.1f} s of frame-parallel conversion on all host cores"},
            "cpu_baseline": {"value": round(v, 2), "unit": "frames/s", "cores": cores, "kind": kd, "cpu_model": cpu_model(),
                             "sample": f"{steps} x {per_step:.1f} s, {cores} pthreads, each cycling through 4 distinct frames; "
                                       "the clock starts after every thread has filled its buffers"},
            "e2e": {"value": round(v, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }

    # a step = a bounded slice of frame-parallel conversion; the whole run stays near two minutes whatever K and W are
    per_step = max(0.2, min(1.0, 120.0 / max(1, args.steps + args.warmup)))
    line = one(name, args.steps, args.warmup, per_step)
    if name == "yuv420p_rgb24_1080p" and not args.no_extra:
        line["extra"] = [x for x in (one(nm, 2, 0, 1.5) for nm in EXTRA_DEFAULT) if x]
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
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs carried in `extra`")
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"], help="process-group backend (gloo: CPU tests)")
    ap.add_argument("--dry-run", action="store_true",
                    help="exercise the multi-rank plumbing with no GPU work (tests/test_bench_dist.py); not a measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "acgpu":
        args.warmup = 3
    name = args.workload

    if args.impl == "reference":
        run_reference(args, name)
        return

    dc = Dist(args.backend)
    rank, local_rank, world = dc.rank, dc.local_rank, dc.world

    if args.dry_run:
        # no device: each rank pretends its steps took (rank+1) ms so the max-over-ranks path is observable
        batch = args.batch or WORKLOADS[name][5]
        dc.barrier()
        ms = dc.max_float(float(rank + 1) * args.steps)
        total_units = dc.sum_int(batch)
        dc.barrier()
        if rank == 0:
            print(json.dumps({"dry_run": True, "n_gpus": world, "steps": args.steps, "ms_max": ms,
                              "units_all_ranks": total_units, "value": aggregate_value(total_units, args.steps, ms)}), flush=True)
        dc.close()
        return

    ac = pkg.AcGpu()
    _AC[0] = ac
    if not ac.lib.acgpu_set_device(local_rank):
        raise SystemExit("libacgpu: " + ac.last_error())
    if ac.ac_init(pkg.AC_CUDA) != 1:
        raise SystemExit("libacgpu: ac_init(AC_CUDA) failed: " + ac.last_error() + " (no CPU fallback)")
    ac.lib.acgpu_force_tier(args.tier)

    line = measure(ac, dc, args, name, args.steps, args.warmup)
    extras = []
    if name == "yuv420p_rgb24_1080p" and not args.no_extra:
        for nm in EXTRA_DEFAULT:
            r = measure(ac, dc, args, nm, max(3, min(args.steps, 8)), 3, clocks=False)
            if r:
                for k in ("higher_is_better", "scaling", "vs_baseline", "dtype", "data", "warmup"):
                    r.pop(k, None)
                extras.append(r)
    # the same host-frame run cut across every GPU this ONE process can see (acgpu_imgconvert_frames_host_multi)
    if rank == 0 and world == 1 and not args.no_e2e and name == "yuv420p_rgb24_1080p" and ac.lib.acgpu_device_count() > 1:
        kind, a, b, w, h, batch = WORKLOADS[name]
        sfb, dfb = F.frame_bytes(a, w, h), F.frame_bytes(b, w, h)
        nd = ac.lib.acgpu_device_count()
        eb = 64 * nd
        hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
        hs.array[:] = 0x80
        for _ in range(2):
            ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, a, hd.ptr, b, w, h, eb, nd))
        t0 = time.perf_counter()
        for _ in range(5):
            ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, a, hd.ptr, b, w, h, eb, nd))
        line["e2e"]["one_process_all_devices"] = {"devices": nd, "frames_per_s": round(5 * eb / (time.perf_counter() - t0), 2)}
        hs.free(); hd.free()
    if rank == 0:
        if extras:
            line["extra"] = extras
        print(json.dumps(line), flush=True)
    dc.close()


if __name__ == "__main__":
    main()
