#!/usr/bin/env python
"""Horizontal tcv_resize throughput probe (profiling aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e.load_package(); ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_CUDA) == 1
lib = ac.lib; w, h, nf = 1920, 1080, 256
st = lib.acgpu_stream_create(); e0, e1 = lib.acgpu_event_create(), lib.acgpu_event_create()
for bpp, nw in ((1, 1280), (3, 1280), (1, 2560)):
    n = nf // bpp
    src = ac.malloc(n * w * h * bpp + w * bpp); dst = ac.malloc(n * nw * h * bpp)
    rw = (nw - w) // 8
    for _ in range(3): lib.acgpu_resize_batch(src.ptr, dst.ptr, w, h, bpp, rw, 0, 8, 8, w*h*bpp, nw*h*bpp, n, st)
    lib.acgpu_event_record(e0, st)
    for _ in range(10): assert lib.acgpu_resize_batch(src.ptr, dst.ptr, w, h, bpp, rw, 0, 8, 8, w*h*bpp, nw*h*bpp, n, st) == 1
    lib.acgpu_event_record(e1, st); ac.sync(st)
    ms = lib.acgpu_event_elapsed_ms(e0, e1) / 10
    print("resize_h %d->%d Bpp %d: %.0f frames/s, %.0f GB/s unique" % (w, nw, bpp, n/ms*1e3, n*(w*h+nw*h)*bpp/ms/1e6))
    src.free(); dst.free()
