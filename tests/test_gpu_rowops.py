"""-m gpu parity for ac_average / ac_rescale and the libtcvideo row shapes built on them (BASELINE config 3)."""
import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck
from test_oracle import average_vectors

pkg = entry.load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def oracle():
    return ck.Oracle()


@pytest.fixture(scope="module")
def tcv():
    """The reference libtcvideo itself when oracle/_ref/libtcv_ref.so travelled with the repo, else the restatement."""
    return ck.best_tcv_checker()


def test_average_known_answers_with_guard_bands(ac):
    """testsuite/test-average.c:177-645 (135 vectors, exact (a+b+1)/2, 8-byte 0x11 guard bands)."""
    spill = 8
    for a, b in average_vectors():
        n = a.size
        want = ((a.astype(np.int32) + b + 1) // 2).astype(np.uint8)
        buf = np.full(n + 2 * spill, 0x11, np.uint8)
        ac.lib.ac_average(a.ctypes.data, b.ctypes.data, buf.ctypes.data + spill, n)
        assert np.array_equal(buf[spill:spill + n], want)
        assert (buf[:spill] == 0x11).all() and (buf[spill + n:] == 0x11).all()


def test_average_and_rescale_exhaustive_byte_pairs(ac, oracle):
    a = np.repeat(np.arange(256, dtype=np.uint8), 256)
    b = np.tile(np.arange(256, dtype=np.uint8), 256)
    assert np.array_equal(ac.ac_average(a, b), oracle.average(a, b))
    weights = [(0, 65536), (1, 65535), (32767, 32769), (32768, 32768), (65535, 1), (65536, 0), (49152, 16384),
               (16384, 49152), (40000, 40000), (65535, 65535), (12345, 54321), (70000, 5), (0, 0), (65535, 0)]
    weights += [(w, 65536 - w) for w in range(7, 65536, 4099)]
    for w1, w2 in weights:
        assert np.array_equal(ac.ac_rescale(a, b, w1, w2), oracle.rescale(a, b, w1, w2)), (w1, w2)


def test_rescale_copy_branch_never_reads_src2(ac):
    """rescale.c:26-29: with weight1 >= 65536 src2 may point past the frame (tcvideo.c:469-470)."""
    a = np.arange(100, dtype=np.uint8)
    d = np.zeros(100, np.uint8)
    ac.lib.ac_rescale(a.ctypes.data, 8, d.ctypes.data, 100, 65536, 0)      # src2 = bogus pointer
    assert np.array_equal(d, a)


def test_unaligned_and_aliased_blends(ac, oracle):
    big = ck.splitmix_bytes(5000, 3)
    for off1, off2, n in [(1, 2, 777), (3, 16, 1920), (0, 0, 33), (5, 5, 1)]:
        a, b = big[off1:off1 + n], big[off2 + 2000:off2 + 2000 + n]
        assert np.array_equal(ac.ac_average(a, b), oracle.average(a, b))
        assert np.array_equal(ac.ac_rescale(a, b, 20000, 45536), oracle.rescale(a, b, 20000, 45536))
    # dest aliasing a source (libtcvideo/tcvideo.c:381,386)
    a, b = big[:1920].copy(), big[2000:3920].copy()
    want = oracle.average(a, b)
    ac.lib.ac_average(a.ctypes.data, b.ctypes.data, b.ctypes.data, a.size)
    assert np.array_equal(b, want)


@pytest.mark.parametrize("bpp", [1, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("size", [(1920, 1080), (720, 577), (64, 2), (50, 7)])
def test_deinterlace_shapes(ac, tcv, bpp, mode, size):
    w, h = size
    nf = 2
    fb = w * h * bpp
    frames = np.stack([ck.splitmix_bytes(fb, 40 + i) for i in range(nf)])
    src = ac.malloc(nf * fb).upload(frames.reshape(-1))
    dst = ac.malloc(nf * fb).fill(0x55)
    ac._ok(ac.lib.acgpu_deinterlace_batch(src.ptr, dst.ptr, w, h, bpp, mode, fb, fb, nf, None))
    ac.sync()
    got = dst.download().reshape(nf, fb)
    for i in range(nf):
        assert np.array_equal(got[i], tcv.deinterlace(frames[i], w, h, bpp, mode)), (size, bpp, mode, i)
    assert np.array_equal(src.download().reshape(nf, fb), frames)     # src is left intact
    src.free(); dst.free()


@pytest.mark.parametrize("bpp", [1, 3])
@pytest.mark.parametrize("case", [
    (1920, 1080, 0, -45, 8, 8),    # 1080 -> 720  
    (1280, 720, 0, 45, 8, 8),      # 720 -> 1080: a third of the rows take the copy branch
    (960, 540, 0, -15, 4, 4),      # 4:2:0 chroma plane, scale 4
    (1920, 1080, -80, 0, 8, 8),    # horizontal 1920 -> 1280
    (720, 576, 4, 0, 8, 8),        # horizontal grow
    (64, 32, 0, 3, 8, 2),
    (1920, 1080, -160, 0, 8, 8),   # horizontal 1920 -> 640: ratio 3, taps too far apart for the window kernel
    (640, 480, 80, 0, 8, 8),       # horizontal 640 -> 1280
    (1920, 1080, -1, 0, 8, 8),     # 1920 -> 1912: nearly 1:1, every phase of the weight table
    (1024, 64, -60, 0, 8, 4),      # 1024 -> 544 (ratio 1.88), scale_h 4
])
def test_resize_shapes(ac, tcv, bpp, case):
    w, h, rw, rh, sw, sh = case
    nw, nh = w + rw * sw, h + rh * sh
    nf = 2
    sfb, dfb = w * h * bpp, nw * nh * bpp
    frames = np.stack([ck.splitmix_bytes(sfb, 60 + i) for i in range(nf)])
    # one spare row after the last frame: the copy branch must not read it, the blend branch never needs it
    src = ac.malloc(nf * sfb + w * bpp).fill(0xEE)
    src.upload(frames.reshape(-1))
    dst = ac.malloc(nf * dfb).fill(0x55)
    want = [tcv.resize(frames[i], w, h, bpp, rw, rh, sw, sh) for i in range(nf)]
    for tier in (0, 1):       # 1: the byte-gather horizontal kernel instead of the window kernel
        dst.fill(0x55)
        ac.lib.acgpu_force_tier(tier)
        try:
            ac._ok(ac.lib.acgpu_resize_batch(src.ptr, dst.ptr, w, h, bpp, rw, rh, sw, sh, sfb, dfb, nf, None))
        finally:
            ac.lib.acgpu_force_tier(0)
        ac.sync()
        got = dst.download().reshape(nf, dfb)
        for i in range(nf):
            assert np.array_equal(got[i], want[i]), (case, bpp, tier, i)
    src.free(); dst.free()


def test_resize_rejects_what_tcv_resize_rejects(ac):
    buf = ac.malloc(1 << 16)
    bad = [(64, 32, 2, 0, 0, 8, 8), (64, 32, 1, 0, -1, 3, 8), (60, 32, 1, 0, -1, 8, 8), (64, 32, 1, -8, 0, 8, 8), (64, 32, 1, 1, 1, 8, 8)]
    for (w, h, bpp, rw, rh, sw, sh) in bad:
        assert ac.lib.acgpu_resize_batch(buf.ptr, buf.ptr, w, h, bpp, rw, rh, sw, sh, 0, 0, 1, None) == 0
    buf.free()


@pytest.mark.parametrize("mode,first", [(2, 1), (3, 0)])
def test_deinterlace_drop_field(ac, mode, first):
    """tcvideo.c:326-338 -- drop top field keeps lines 1,3,5..., drop bottom keeps 0,2,4...; height/2 rows written."""
    for (w, h, bpp) in [(1920, 1080, 1), (720, 577, 3), (50, 7, 3)]:
        nf, bpl = 2, w * bpp
        frames = np.stack([ck.splitmix_bytes(bpl * h, 80 + i) for i in range(nf)])
        src = ac.malloc(nf * bpl * h).upload(frames.reshape(-1))
        dst = ac.malloc(nf * bpl * h).fill(0x55)
        ac._ok(ac.lib.acgpu_deinterlace_batch(src.ptr, dst.ptr, w, h, bpp, mode, bpl * h, bpl * h, nf, None))
        ac.sync()
        got = dst.download().reshape(nf, h, bpl)
        for i in range(nf):
            want = frames[i].reshape(h, bpl)[first::2][: h // 2]
            assert np.array_equal(got[i, : h // 2], want)
            assert (got[i, h // 2:] == 0x55).all()
        src.free(); dst.free()


def test_table_cache_eviction_never_frees_a_table_in_use(ac, tcv):
    """libacgpu caches the small device tables it builds (resize weights / window selectors, gamma and antialias
    tables, row-operation lists) per thread, 64 entries, least recently used.  A call that needs two tables must not
    lose the first one when the second one's miss evicts old entries: run well past the cache size with calls that
    need one table and calls that need two, checking every result."""
    w, h, bpp = 256, 8, 1
    src = ck.splitmix_bytes(w * h * bpp, 123)
    f = src[None, :]
    for i in range(150):
        g = 0.5 + i * 0.01                                    # one new table per call
        ok, got = ac.plane_op_batch("gamma_correct", f, w * h * bpp, w, h, bpp, g)
        assert ok == 1 and np.array_equal(got[0], tcv.gamma(src, w, h, bpp, g)[1]), ("gamma", i)
        rw = -(i % 11) - 1                                    # 256 -> 248 ... 168: two new tables per distinct width
        width = w - 8 * (i % 3)                               # 256, 248, 240 wide sources: 33 distinct table pairs
        s2 = src[: width * h]
        nw = width + rw * 8
        ok, got = ac.plane_op_batch("resize", s2[None, :], nw * h * bpp, width, h, bpp, rw, 0, 8, 8)
        assert ok == 1 and np.array_equal(got[0], tcv.resize(s2, width, h, bpp, rw, 0, 8, 8)), ("resize", i, width, nw)
