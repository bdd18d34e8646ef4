import os, sys, time, threading
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); F = pkg.F
ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_ALL) == 1
w, h = 1920, 1080
def worker(n, out, i):
    src = np.random.default_rng(i).integers(0, 256, F.frame_bytes(F.IMG_YUV420P, w, h), dtype=np.uint8)
    dst = np.zeros(F.frame_bytes(F.IMG_RGB24, w, h), np.uint8)
    a = pkg.AcGpu()
    for _ in range(3): a.ac_imgconvert(src, F.IMG_YUV420P, dst, F.IMG_RGB24, w, h)
    t0 = time.perf_counter()
    for _ in range(n): a.ac_imgconvert(src, F.IMG_YUV420P, dst, F.IMG_RGB24, w, h)
    out[i] = n / (time.perf_counter() - t0)
for T in (1, 2, 4, 8, 16):
    out = [0] * T
    th = [threading.Thread(target=worker, args=(60, out, i)) for i in range(T)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"legacy ac_imgconvert host pageable 1080p 420P->RGB24: {T:2d} threads -> {T*63/dt:8.1f} frames/s total")
# the same unmodified ac_imgconvert call when the CALLER's frame buffers are page-locked (transcode
# allocating its vframe buffers with acgpu_host_alloc instead of tc_bufalloc, libtc/tcframes.c:214-220): no bounce copy
def worker_pinned(n, out, i):
    a = pkg.AcGpu()
    ps = pkg.PinnedBuffer(a, F.frame_bytes(F.IMG_YUV420P, w, h)); pd = pkg.PinnedBuffer(a, F.frame_bytes(F.IMG_RGB24, w, h))
    ps.array[:] = np.random.default_rng(i).integers(0, 256, ps.nbytes, dtype=np.uint8)
    for _ in range(3): a.ac_imgconvert(ps.array, F.IMG_YUV420P, pd.array, F.IMG_RGB24, w, h)
    t0 = time.perf_counter()
    for _ in range(n): a.ac_imgconvert(ps.array, F.IMG_YUV420P, pd.array, F.IMG_RGB24, w, h)
    out[i] = n / (time.perf_counter() - t0)
    ps.free(); pd.free()
for T in (1, 2, 4, 8, 16):
    out = [0] * T
    th = [threading.Thread(target=worker_pinned, args=(120, out, i)) for i in range(T)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"legacy ac_imgconvert host PINNED   1080p 420P->RGB24: {T:2d} threads -> {T*123/dt:8.1f} frames/s total")
# single-frame latency with device pointers
src = np.random.default_rng(0).integers(0, 256, F.frame_bytes(F.IMG_YUV420P, w, h), dtype=np.uint8)
ds = ac.malloc(src.size).upload(src); dd = ac.malloc(F.frame_bytes(F.IMG_RGB24, w, h))
import ctypes as C
so, do = F.plane_offsets(F.IMG_YUV420P, w, h), [0]
sp = (C.c_void_p * 3)(*[ds.ptr + o for o in so]); dp = (C.c_void_p * 3)(dd.ptr, None, None)
for _ in range(10): ac.lib.ac_imgconvert(sp, F.IMG_YUV420P, dp, F.IMG_RGB24, w, h)
t0 = time.perf_counter()
for _ in range(2000): ac.lib.ac_imgconvert(sp, F.IMG_YUV420P, dp, F.IMG_RGB24, w, h)
dt = (time.perf_counter() - t0) / 2000
print(f"legacy ac_imgconvert DEVICE pointers, 1 frame per call (sync): {dt*1e6:.1f} us/call = {1/dt:.0f} frames/s")
