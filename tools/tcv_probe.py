#!/usr/bin/env python
"""Throughput probe for the libtcvideo plane operations on the device (profiling aid, not the contract bench).

    python tools/tcv_probe.py [--size 1920x1080] [--frames 128]
Each line: one launch per step over `frames` planes (about 1 GB of traffic), CUDA-event timed on the launch stream;
GB/s counts the bytes the C reference reads plus the bytes it writes (algorithmic bytes), fraction is of the measured
HBM copy peak in MEASURED_PEAKS.json.
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--only", default="", help="run only the lines whose label contains this text (for ncu captures)")
ap.add_argument("--bpp", type=int, default=0, help="1 or 3 (default: both)")
a = ap.parse_args()
w, h = map(int, a.size.split("x"))
pkg = e.load_package(); ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_CUDA) == 1
lib = ac.lib
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
st = lib.acgpu_stream_create(); e0, e1 = lib.acgpu_event_create(), lib.acgpu_event_create()


def blocky(w, h, bpp):
    base = (np.add.outer(np.arange(h) // 3 * 40, np.arange(w) // 5 * 30) % 256).astype(np.uint8)
    return np.repeat(base.reshape(-1), bpp)


def run(name, fn, n, in_b, out_b):
    if a.only and a.only not in name:
        return
    for _ in range(3):
        assert fn() == 1, lib.acgpu_last_error()
    lib.acgpu_event_record(e0, st)
    for _ in range(a.steps):
        fn()
    lib.acgpu_event_record(e1, st); ac.sync(st)
    ms = lib.acgpu_event_elapsed_ms(e0, e1) / a.steps
    gbs = n * (in_b + out_b) / ms / 1e6
    print("%-34s %9.0f frames/s %8.1f GB/s  %.3f" % (name, n / ms * 1e3, gbs, gbs / peak), flush=True)


for bpp in ((a.bpp,) if a.bpp else (1, 3)):
    n = max(a.frames // bpp, 1)
    fb = w * h * bpp
    src = ac.malloc(n * fb); dst = ac.malloc(n * (w + 64) * (h + 16) * bpp)
    rnd = np.random.default_rng(1).integers(0, 256, fb, dtype=np.uint8)
    def fill(img):
        for i in range(n):
            src.upload(img, offset=i * fb)
    fill(rnd)
    S, D = src.ptr, dst.ptr
    t = f"{w}x{h} Bpp {bpp}"
    cw, ch = w - 32, h - 16
    run(f"clip -16/-16/-8/-8    {t}", lambda: lib.acgpu_clip_batch(S, D, w, h, bpp, 16, 16, 8, 8, 16, fb, cw * ch * bpp, n, st), n, cw * ch * bpp, cw * ch * bpp)
    gw, gh = w + 64, h + 16
    run(f"clip grow +32/+32/+8/+8 {t}", lambda: lib.acgpu_clip_batch(S, D, w, h, bpp, -32, -32, -8, -8, 16, fb, gw * gh * bpp, n, st), n, fb, gw * gh * bpp)
    run(f"clip odd 3/5/1/1      {t}", lambda: lib.acgpu_clip_batch(S, D, w, h, bpp, 3, 5, 1, 1, 16, fb, (w - 8) * (h - 2) * bpp, n, st), n, (w - 8) * (h - 2) * bpp, (w - 8) * (h - 2) * bpp)
    run(f"reduce 2x2            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 2, 2, fb, fb // 4, n, st), n, fb // 4, fb // 4)
    run(f"reduce 3x3            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 3, 3, fb, fb // 9, n, st), n, fb // 9, fb // 9)
    run(f"reduce 4x4            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 4, 4, fb, fb // 16, n, st), n, fb // 16, fb // 16)
    run(f"reduce 5x5            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 5, 5, fb, (w // 5) * (h // 5) * bpp, n, st), n, (w // 5) * (h // 5) * bpp, (w // 5) * (h // 5) * bpp)
    run(f"reduce 8x8 (gather)   {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 8, 8, fb, (w // 8) * (h // 8) * bpp, n, st), n, (w // 8) * (h // 8) * bpp, (w // 8) * (h // 8) * bpp)
    run(f"reduce 6x6            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 6, 6, fb, (w // 6) * (h // 6) * bpp, n, st), n, (w // 6) * (h // 6) * bpp, (w // 6) * (h // 6) * bpp)
    run(f"reduce 7x7 (gather)   {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 7, 7, fb, (w // 7) * (h // 7) * bpp, n, st), n, (w // 7) * (h // 7) * bpp, (w // 7) * (h // 7) * bpp)
    run(f"reduce 1x2            {t}", lambda: lib.acgpu_reduce_batch(S, D, w, h, bpp, 1, 2, fb, fb // 2, n, st), n, fb // 2, fb // 2)
    run(f"flip_v                {t}", lambda: lib.acgpu_flip_v_batch(S, D, w, h, bpp, fb, fb, n, st), n, fb, fb)
    run(f"flip_v in place       {t}", lambda: lib.acgpu_flip_v_batch(S, S, w, h, bpp, fb, fb, n, st), n, fb, fb)
    run(f"flip_h                {t}", lambda: lib.acgpu_flip_h_batch(S, D, w, h, bpp, fb, fb, n, st), n, fb, fb)
    run(f"flip_h in place       {t}", lambda: lib.acgpu_flip_h_batch(S, S, w, h, bpp, fb, fb, n, st), n, fb, fb)
    run(f"gamma 2.2             {t}", lambda: lib.acgpu_gamma_correct_batch(S, D, w, h, bpp, 2.2, fb, fb, n, st), n, fb, fb)
    aa = lambda: lib.acgpu_antialias_batch(S, D, w, h, bpp, 0.333, 0.5, fb, fb, n, st)
    # how much of the picture takes the 9-tap path decides the cost: uniform random bytes smooth ~40 % of the pixels
    # at Bpp 1 and ~3 % at Bpp 3, flat blocks with edges ~15 %, a soft gradient almost none
    run(f"antialias random bytes {t}", aa, n, fb, fb)
    fill(blocky(w, h, bpp))
    run(f"antialias blocky       {t}", aa, n, fb, fb)
    yy, xx = np.mgrid[0:h, 0:w]
    fill(np.repeat(((xx // 3 + yy // 2) % 256).astype(np.uint8).reshape(-1), bpp))
    run(f"antialias gradient     {t}", aa, n, fb, fb)
    src.free(); dst.free()
