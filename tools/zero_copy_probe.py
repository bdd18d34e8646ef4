#!/usr/bin/env python
"""Probe: can the conversion kernel write its output straight into page-locked HOST memory (zero-copy stores over PCIe)
faster than kernel + D2H copy?  1080p YUV420P -> RGB24, device-resident source (profiling aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e.load_package(); F = pkg.F; ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_CUDA) == 1
lib = ac.lib
w, h, nf = 1920, 1080, 256
sf, df = F.IMG_YUV420P, F.IMG_RGB24
sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
src = ac.malloc(nf * sfb); hd = ac.pinned(nf * dfb); dd = ac.malloc(nf * dfb)
hs = ac.pinned(nf * sfb); hs.array[:] = 9
st = lib.acgpu_stream_create()
def timed(fn, n=5):
    fn(); ac.sync(st)
    t = time.perf_counter()
    for _ in range(n): fn()
    ac.sync(st)
    return (time.perf_counter() - t) / n
t = timed(lambda: ac._ok(ac.imgconvert_batch(src.ptr, sf, sfb, hd.ptr, df, dfb, w, h, nf, st)))
print("kernel storing straight into pinned host memory: %.0f frames/s (%.1f GB/s over PCIe)" % (nf / t, nf * dfb / t / 1e9))
def kd():
    ac._ok(ac.imgconvert_batch(src.ptr, sf, sfb, dd.ptr, df, dfb, w, h, nf, st))
    lib.acgpu_memcpy_d2h(hd.ptr, dd.ptr, nf * dfb, st)
t = timed(kd)
print("kernel to device memory + one D2H copy:          %.0f frames/s (%.1f GB/s)" % (nf / t, nf * dfb / t / 1e9))
t = timed(lambda: ac._ok(lib.acgpu_imgconvert_frames_host(hs.ptr, sf, hd.ptr, df, w, h, nf)))
print("acgpu_imgconvert_frames_host (H2D + kernel + D2H, pipelined): %.0f frames/s" % (nf / t))
