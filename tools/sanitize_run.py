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
print("sanitize_run: ok,", n, "conversions")
