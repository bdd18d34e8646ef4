"""-m gpu: the small device / memory / stream / event helpers of include/acgpu.h and the device-pointer forms of
ac_average / ac_rescale (acgpu_average, acgpu_rescale)."""
import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck

pkg = entry.load_package()
F = pkg.F
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


def test_version_device_and_counters(ac):
    L = ac.lib
    assert L.acgpu_version().startswith(b"libacgpu")
    n = L.acgpu_device_count()
    assert n >= 1
    cur = L.acgpu_get_device()
    assert 0 <= cur < n
    assert L.acgpu_set_device(cur) == 1 and L.acgpu_get_device() == cur
    assert L.acgpu_set_device(n) == 0 and L.acgpu_get_device() == cur          # refused, binding unchanged
    assert L.acgpu_device_sm_count() == 148                                     # B200
    # launch counter: what bench.py reports as gpu_launches
    w, h, nf = 128, 8, 2
    src = ac.malloc(nf * F.frame_bytes(F.IMG_YUV420P, w, h)).fill(0x40)
    dst = ac.malloc(nf * F.frame_bytes(F.IMG_RGB24, w, h))
    L.acgpu_launch_count(1)
    assert L.acgpu_launch_count(0) == 0
    for _ in range(3):
        ac._ok(ac.imgconvert_batch(src.ptr, F.IMG_YUV420P, F.frame_bytes(F.IMG_YUV420P, w, h), dst.ptr, F.IMG_RGB24,
                                   F.frame_bytes(F.IMG_RGB24, w, h), w, h, nf))
    ac.sync()
    assert L.acgpu_launch_count(0) == 3          # one launch per batch call
    assert L.acgpu_launch_count(1) == 3 and L.acgpu_launch_count(0) == 0
    src.free(); dst.free()


def test_memory_stream_and_event_helpers(ac):
    L = ac.lib
    n = 1 << 20
    a = ck.splitmix_bytes(n, 4)
    d1, d2 = ac.malloc(n).upload(a), ac.malloc(n)
    st = L.acgpu_stream_create()
    e0, e1 = L.acgpu_event_create(), L.acgpu_event_create()
    assert st and e0 and e1
    assert L.acgpu_event_record(e0, st) == 1
    assert L.acgpu_memset(d2.ptr, 0x5C, n, st) == 1
    assert L.acgpu_memcpy_d2d(d2.ptr + 4096, d1.ptr + 100, 50000, st) == 1      # unaligned source, inside the buffer
    assert L.acgpu_event_record(e1, st) == 1
    assert L.acgpu_event_sync(e1) == 1 and L.acgpu_stream_sync(st) == 1
    assert L.acgpu_event_elapsed_ms(e0, e1) > 0
    got = d2.download()
    want = np.full(n, 0x5C, np.uint8)
    want[4096:4096 + 50000] = a[100:50100]
    assert np.array_equal(got, want)
    p = ac.pinned(4096)
    p.array[:] = 7
    assert L.acgpu_memcpy_h2d(d2.ptr, p.ptr, 4096, st) == 1 and L.acgpu_stream_sync(st) == 1
    assert (d2.download(4096) == 7).all()
    p.free()
    L.acgpu_event_destroy(e0); L.acgpu_event_destroy(e1); L.acgpu_stream_destroy(st)
    d1.free(); d2.free()


def test_device_pointer_average_and_rescale(ac):
    """acgpu_average / acgpu_rescale: ac_average / ac_rescale on device memory, asynchronous on a caller stream."""
    L = ac.lib
    oracle = ck.Oracle()
    big = ck.splitmix_bytes(1 << 16, 8)
    buf = ac.malloc(big.size).upload(big)
    out = ac.malloc(1 << 15)
    st = L.acgpu_stream_create()
    for (o1, o2, n) in [(0, 32768, 16384), (1, 20003, 9999), (48, 4096, 1), (16, 32, 5760)]:
        a, b = big[o1:o1 + n], big[o2:o2 + n]
        out.fill(0x99)
        assert L.acgpu_average(buf.ptr + o1, buf.ptr + o2, out.ptr + 3, n, st) == 1
        L.acgpu_stream_sync(st)
        got = out.download()
        assert np.array_equal(got[3:3 + n], oracle.average(a, b)) and got[2] == 0x99 and got[3 + n] == 0x99
        for w1, w2 in [(49152, 16384), (1, 65535), (40000, 40000), (65535, 65535)]:
            assert L.acgpu_rescale(buf.ptr + o1, buf.ptr + o2, out.ptr, n, w1, w2, st) == 1
            L.acgpu_stream_sync(st)
            assert np.array_equal(out.download(n), oracle.rescale(a, b, w1, w2)), (o1, o2, n, w1, w2)
        # copy branches: the other source is never read (rescale.c:26-29) -- hand it an address far outside any allocation
        assert L.acgpu_rescale(buf.ptr + o1, 1 << 44, out.ptr, n, 65536, 0, st) == 1
        L.acgpu_stream_sync(st)
        assert np.array_equal(out.download(n), a)
        assert L.acgpu_rescale(1 << 44, buf.ptr + o2, out.ptr, n, 0, 70000, st) == 1
        L.acgpu_stream_sync(st)
        assert np.array_equal(out.download(n), b)
    L.acgpu_stream_destroy(st)
    buf.free(); out.free()


def test_ac_memcpy_on_device_memory_is_memmove(ac):
    """aclib/memcpy.c:16-25 is memmove (ac.h:80-82 promises the ascending-copy behaviour callers rely on for overlapping
    buffers); device-to-device copies keep that for overlapping ranges in both directions, and mixed host/device work."""
    L = ac.lib
    n = 100000
    a = ck.splitmix_bytes(n, 12)
    for (dst_off, src_off, size) in [(0, 1000, 50000), (1000, 0, 50000), (5, 4, 99000), (0, 60000, 30000)]:
        buf = ac.malloc(n).upload(a)
        assert L.ac_memcpy(buf.ptr + dst_off, buf.ptr + src_off, size) == buf.ptr + dst_off
        want = a.copy()
        want[dst_off:dst_off + size] = a[src_off:src_off + size]
        assert np.array_equal(buf.download(), want), (dst_off, src_off, size)
        buf.free()
    buf = ac.malloc(n)
    assert L.ac_memcpy(buf.ptr, a.ctypes.data, n) == buf.ptr                    # host -> device
    back = np.zeros(n, np.uint8)
    assert L.ac_memcpy(back.ctypes.data, buf.ptr, n) == back.ctypes.data        # device -> host
    assert np.array_equal(back, a)
    buf.free()
