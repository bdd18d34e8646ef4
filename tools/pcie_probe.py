#!/usr/bin/env python
"""PCIe ceiling probe: pinned H2D alone, D2H alone, both at once (profiling aid for the e2e number)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
ac = pkg.AcGpu()
assert ac.ac_init(pkg.AC_CUDA) == 1
lib = ac.lib
N = 768 << 20
h1, h2 = ac.pinned(N), ac.pinned(N)
d1, d2 = ac.malloc(N), ac.malloc(N)
s1, s2 = lib.acgpu_stream_create(), lib.acgpu_stream_create()


def timed(fn, reps=5):
    fn()
    ac.sync(s1); ac.sync(s2)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ac.sync(s1); ac.sync(s2)
    return (time.perf_counter() - t0) / reps


t = timed(lambda: lib.acgpu_memcpy_h2d(d1.ptr, h1.ptr, N, s1))
print(f"H2D alone       {N / t / 1e9:6.1f} GB/s")
t = timed(lambda: lib.acgpu_memcpy_d2h(h2.ptr, d2.ptr, N, s2))
print(f"D2H alone       {N / t / 1e9:6.1f} GB/s")
t = timed(lambda: (lib.acgpu_memcpy_h2d(d1.ptr, h1.ptr, N // 2, s1), lib.acgpu_memcpy_d2h(h2.ptr, d2.ptr, N, s2)))
print(f"H2D(N/2)+D2H(N) {N / t / 1e9:6.1f} GB/s D2H, {N / 2 / t / 1e9:6.1f} GB/s H2D (concurrent)")
for mb in (4, 16, 64):
    c = mb << 20
    def chunks():
        for o in range(0, N, c):
            lib.acgpu_memcpy_d2h(h2.ptr + o, d2.ptr + o, c, s2)
    t = timed(chunks)
    print(f"D2H in {mb:3d} MB chunks {N / t / 1e9:6.1f} GB/s")
