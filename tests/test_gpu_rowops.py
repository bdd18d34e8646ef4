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


def _model_rowop(src, op, n):
    """numpy statement of one acgpu_rowop (aclib/rescale.c:23-46, average.c:33-39) on a flat source plane."""
    a = src[op.src1_off: op.src1_off + n].astype(np.uint64)
    if op.op == pkg.ACGPU_ROW_COPY:
        return a.astype(np.uint8)
    b = src[op.src2_off: op.src2_off + n].astype(np.uint64)
    if op.op == pkg.ACGPU_ROW_AVERAGE:
        return ((a + b + 1) >> 1).astype(np.uint8)
    if op.op == pkg.ACGPU_ROW_AVERAGE3:
        c = src[op.src3_off: op.src3_off + n].astype(np.uint64)
        return ((c + ((a + b + 1) >> 1) + 1) >> 1).astype(np.uint8)
    if op.weight1 >= 0x10000:
        return a.astype(np.uint8)
    if op.weight2 >= 0x10000:
        return b.astype(np.uint8)
    return (((a * op.weight1 + b * op.weight2 + 32768) & 0xFFFFFFFF) >> 16).astype(np.uint8)


@pytest.mark.parametrize("aligned", [True, False])
def test_rowops_run_with_caller_built_tables(ac, aligned):
    """acgpu_rowops_run directly: random tables of all four operations, weights on both sides of 65536 (the copy
    branches must not read the other row: its offset points far outside the plane), rows at aligned offsets (tiled
    cp.async kernel) or byte offsets (byte kernel), several frames with a pitch gap."""
    rng = np.random.default_rng(321 + aligned)
    for trial in range(12):
        row = int(rng.integers(1, 40)) * 16 if aligned else int(rng.integers(1, 700))
        rows_src, nops, nf = int(rng.integers(3, 40)), int(rng.integers(1, 60)), int(rng.integers(1, 4))
        step = row if aligned else row + int(rng.integers(0, 5))
        sfb = rows_src * step + row
        dfb = nops * step + row
        spitch = sfb + (32 if aligned else 7)
        dpitch = dfb + (48 if aligned else 11)
        frames = [ck.splitmix_bytes(sfb, 1000 * trial + i) for i in range(nf)]
        ops = (pkg.RowOp * nops)()
        far = 1 << 40                                     # never dereferenced
        for k in range(nops):
            o = ops[k]
            o.op = int(rng.integers(0, 4))
            o.src1_off, o.src2_off, o.src3_off = (int(rng.integers(0, rows_src)) * step for _ in range(3))
            o.dest_off = k * step
            w1 = int(rng.choice([0, 1, 32768, 49152, 65535, 65536, 70000, int(rng.integers(0, 65536))]))
            o.weight1, o.weight2 = w1, int(rng.choice([65536 - min(w1, 65536), 65536, int(rng.integers(0, 65536))]))
            if o.op == pkg.ACGPU_ROW_RESCALE and o.weight1 >= 0x10000:
                o.src2_off = far
            elif o.op == pkg.ACGPU_ROW_RESCALE and o.weight2 >= 0x10000:
                o.src1_off = far
            elif o.op == pkg.ACGPU_ROW_COPY:
                o.src2_off = o.src3_off = far
        hs = np.full(nf * spitch, 0x33, np.uint8)
        for i in range(nf):
            hs[i * spitch: i * spitch + sfb] = frames[i]
        dsrc = ac.malloc(hs.size).upload(hs)
        ddst = ac.malloc(nf * dpitch).fill(0x77)
        ac._ok(ac.lib.acgpu_rowops_run(dsrc.ptr, spitch, ddst.ptr, dpitch, ops, nops, row, nf, None))
        ac.sync()
        got = ddst.download().reshape(nf, dpitch)
        for i in range(nf):
            want = np.full(dpitch, 0x77, np.uint8)
            for k in range(nops):
                want[ops[k].dest_off: ops[k].dest_off + row] = _model_rowop(frames[i], ops[k], row)
            assert np.array_equal(got[i], want), (aligned, trial, i)
        dsrc.free(); ddst.free()
