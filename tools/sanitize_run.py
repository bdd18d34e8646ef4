#!/usr/bin/env python
"""Small, fast pass over every kernel for compute-sanitizer (memcheck / racecheck): all 256 pairs through the
generic and the vectorised tier at two sizes, the bulk-store tier, the row-blend shapes.  No checker needed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
F = pkg.F
ac = pkg.AcGpu()
assert ac.ac_init(pkg.AC_CUDA) == 1
rng = np.random.default_rng(0)
n = 0
for (w, h) in [(64, 8), (176, 6)]:
    for sf in F.FORMATS_16:
        for df in F.FORMATS_16:
            frames = rng.integers(0, 256, size=(2, F.frame_bytes(sf, w, h)), dtype=np.uint8)
            tiers = [1, 2] + ([3] if df in (F.IMG_RGB24, F.IMG_BGR24) and (sf >> 12) == 1 and sf != F.IMG_Y8 else [])
            for t in tiers:
                ac.lib.acgpu_force_tier(t)
                try:
                    ac.convert_batch(frames, sf, df, w, h)
                    n += 1
                except pkg.AcGpuError:
                    pass            # this size/alignment is outside the forced tier's domain
ac.lib.acgpu_force_tier(0)
for (w, h, bpp) in [(64, 9, 1), (100, 8, 3), (1920, 6, 3)]:
    fb = w * h * bpp
    src = ac.malloc(2 * fb + w * bpp).upload(rng.integers(0, 256, size=2 * fb + w * bpp, dtype=np.uint8))
    dst = ac.malloc(2 * fb)
    for mode in (0, 1):
        ac._ok(ac.lib.acgpu_deinterlace_batch(src.ptr, dst.ptr, w, h, bpp, mode, fb, fb, 2, None))
    ac.sync()
for (w, h, bpp, rw, rh, sw, sh) in [(64, 32, 1, 0, -1, 8, 8), (64, 32, 3, 0, 2, 8, 8), (64, 32, 3, -2, 0, 8, 8), (128, 16, 1, 2, 0, 8, 8)]:
    nw, nh = w + rw * sw, h + rh * sh
    src = ac.malloc(2 * w * h * bpp + w * bpp).upload(rng.integers(0, 256, size=2 * w * h * bpp + w * bpp, dtype=np.uint8))
    dst = ac.malloc(2 * nw * nh * bpp)
    ac._ok(ac.lib.acgpu_resize_batch(src.ptr, dst.ptr, w, h, bpp, rw, rh, sw, sh, w * h * bpp, nw * nh * bpp, 2, None))
    ac.sync()
a = rng.integers(0, 256, size=5000, dtype=np.uint8)
for off, m in [(0, 4096), (1, 777), (3, 33)]:
    ac.ac_average(a[off:off + m], a[off + 100:off + 100 + m])
    ac.ac_rescale(a[off:off + m], a[off + 100:off + 100 + m], 30000, 35536)
# round 2: the tensor-map staged paths (automatic for wide 4:2:0 / 4:1:1 / 4:2:2 -> RGB24, every source with ACGPU_TMA_AUTO=7),
# the plane operations, the chains (fused and not)
for (w, h) in [(1280, 6), (1920, 4), (4160, 4)]:
    for sf in (F.IMG_YUV420P, F.IMG_YUV422P, F.IMG_YUV411P, F.IMG_YUV444P, F.IMG_YUY2, F.IMG_UYVY):
        for df in (F.IMG_RGB24, F.IMG_BGR24):
            if sf == F.IMG_YUV411P and w % 64:
                continue
            frames = rng.integers(0, 256, size=(3, F.frame_bytes(sf, w, h)), dtype=np.uint8)
            ac.convert_batch(frames, sf, df, w, h)
            n += 1
for (w, h, bpp) in [(64, 9, 1), (100, 8, 3), (1920, 6, 3), (1920, 5, 1)]:
    fb = w * h * bpp
    frames = rng.integers(0, 256, size=(2, fb), dtype=np.uint8)
    for op, args, outb in [("clip", (3, 5, 1, 1, 16), (w - 8) * (h - 2) * bpp), ("clip", (-4, 8, 0, -2, 0), (w - 4) * (h + 2) * bpp),
                           ("reduce", (2, 2), (w // 2) * (h // 2) * bpp), ("reduce", (3, 1), (w // 3) * h * bpp), ("reduce", (4, 2), (w // 4) * (h // 2) * bpp),
                           ("reduce", (5, 1), (w // 5) * h * bpp), ("flip_v", (), fb), ("flip_h", (), fb), ("gamma_correct", (0.7,), fb),
                           ("antialias", (0.333, 0.5), fb)]:
        ok, _ = ac.plane_op_batch(op, frames, outb, w, h, bpp, *args)
        assert ok == 1, (op, ac.last_error())
    for op in ("flip_v", "flip_h"):
        ok, _ = ac.plane_op_batch(op, frames, fb, w, h, bpp, inplace=True)
        assert ok == 1
import ctypes as C
for fmt, w, h, stages in [(F.IMG_YUV420P, 128, 16, [(pkg.CHAIN_CONVERT, F.IMG_RGB24), (pkg.CHAIN_CONVERT, F.IMG_YUV422P)]),
                          (F.IMG_YUV420P, 1920, 8, [(pkg.CHAIN_CONVERT, F.IMG_BGR24), (pkg.CHAIN_CONVERT, F.IMG_YUV444P)]),
                          (F.IMG_YUV420P, 128, 32, [(pkg.CHAIN_CLIP, 8, 8, 4, 4), (pkg.CHAIN_DEINTERLACE, 5), (pkg.CHAIN_RESIZE, -2, 1), (pkg.CHAIN_GAMMA, 0.8)]),
                          (F.IMG_RGB24, 96, 24, [(pkg.CHAIN_FLIP_V,), (pkg.CHAIN_RGBSWAP,), (pkg.CHAIN_DECOLOR,), (pkg.CHAIN_ANTIALIAS, 0.3, 0.5), (pkg.CHAIN_REDUCE, 2, 2)])]:
    ops = pkg.chain_ops(stages)
    of, ow, oh = C.c_int(0), C.c_int(0), C.c_int(0)
    assert ac.lib.acgpu_chain_output(fmt, w, h, ops, len(stages), C.byref(of), C.byref(ow), C.byref(oh)) == 1
    inb, outb = F.frame_bytes(fmt, w, h), F.frame_bytes(of.value, ow.value, oh.value)
    src = ac.malloc(3 * inb).upload(rng.integers(0, 256, size=3 * inb, dtype=np.uint8))
    dst = ac.malloc(3 * outb)
    ac._ok(ac.lib.acgpu_chain_batch(src.ptr, fmt, w, h, inb, dst.ptr, outb, ops, len(stages), 3, None))
    ac.sync()
    hs, hd = np.zeros(3 * inb, np.uint8), np.zeros(3 * outb, np.uint8)
    ac._ok(ac.lib.acgpu_chain_frames_host(hs.ctypes.data, fmt, w, h, hd.ctypes.data, ops, len(stages), 3))
print("sanitize_run: ok,", n, "conversions")
