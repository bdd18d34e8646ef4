"""Times the UNMODIFIED reference libtcvideo (oracle/_ref/libtcv_ref.so, plain-C aclib inside) on this host, one thread,
1920x1080 planes -- the CPU figure that sits beside tools/tcv_probe.py's GPU numbers.  TEST INFRASTRUCTURE (it loads the
reference build under oracle/_ref).

    python tests/golden/tcv_cpu_timing.py > profiles/r1c_tcv_reference_cpu.txt
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import checkers as ck  # noqa: E402
import tcv_cases  # noqa: E402

ref = ck.TcvRef()
w, h = 1920, 1080
u8p = C.POINTER(C.c_uint8)


def ptr(a):
    return C.cast(a.ctypes.data, u8p)


def bench(label, fn, in_b, out_b, seconds=1.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn()
        n += 1
    dt = (time.perf_counter() - t0) / n
    print("%-40s %8.1f frames/s %7.2f GB/s" % (label, 1 / dt, (in_b + out_b) / dt / 1e9), flush=True)


print("# reference libtcvideo (plain-C aclib), 1 thread, %dx%d, %s" % (w, h, open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")))
for bpp in (1, 3):
    fb = w * h * bpp
    src = ck.splitmix_bytes(fb, 3)
    blocky = tcv_cases.blocky_image(w, h, bpp, 4)
    dst = np.zeros((w + 64) * (h + 16) * bpp, np.uint8)
    L, H = ref.lib, ref.handle
    t = "Bpp %d" % bpp
    bench(f"clip 16/16/8/8 {t}", lambda: L.tcv_clip(H, ptr(src), ptr(dst), w, h, bpp, 16, 16, 8, 8, 16), (w - 32) * (h - 16) * bpp, (w - 32) * (h - 16) * bpp)
    bench(f"reduce 2x2 {t}", lambda: L.tcv_reduce(H, ptr(src), ptr(dst), w, h, bpp, 2, 2), fb // 4, fb // 4)
    bench(f"flip_v {t}", lambda: L.tcv_flip_v(H, ptr(src), ptr(dst), w, h, bpp), fb, fb)
    bench(f"flip_h {t}", lambda: L.tcv_flip_h(H, ptr(src), ptr(dst), w, h, bpp), fb, fb)
    bench(f"gamma 2.2 {t}", lambda: L.tcv_gamma_correct(H, ptr(src), ptr(dst), w, h, bpp, 2.2), fb, fb)
    bench(f"antialias random bytes {t}", lambda: L.tcv_antialias(H, ptr(src), ptr(dst), w, h, bpp, 0.333, 0.5), fb, fb)
    bench(f"antialias blocky {t}", lambda: L.tcv_antialias(H, ptr(blocky), ptr(dst), w, h, bpp, 0.333, 0.5), fb, fb)
    s2 = src.copy()
    bench(f"deinterlace interpolate {t}", lambda: L.tcv_deinterlace(H, ptr(s2), ptr(dst), w, h, bpp, 2), fb, fb)
    bench(f"deinterlace linear blend {t}", lambda: L.tcv_deinterlace(H, ptr(s2), ptr(dst), w, h, bpp, 3), fb, fb)
    bench(f"resize 1080->720 rows {t}", lambda: L.tcv_resize(H, ptr(src), ptr(dst), w, h, bpp, 0, -45, 8, 8), fb, fb * 2 // 3)
    bench(f"resize 1920->1280 columns {t}", lambda: L.tcv_resize(H, ptr(src), ptr(dst), w, h, bpp, -80, 0, 8, 8), fb, fb * 2 // 3)
