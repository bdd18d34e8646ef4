"""Parity tests proper (-m gpu): libacgpu's CUDA path, called through its C ABI, against the checker.

The checker is the unmodified reference (oracle/_ref, ac_init(AC_NONE)) when its build travelled with
the repo, else the restatement (oracle/liboracle.so); both are pinned to each other by test_oracle.py.
Everything here is integer/byte work, so the bar is bit-exact, including bytes the C path leaves
untouched (dest is pre-filled with 0x55) and 64-byte guard bands.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck
from checkers import F

pkg = entry.load_package()
pytestmark = pytest.mark.gpu

ALL_PAIRS = [(s, d) for s in F.FORMATS_16 for d in F.FORMATS_16]
CFG2_SRC = [F.IMG_YUV420P, F.IMG_YUV422P, F.IMG_YUV444P, F.IMG_YUY2, F.IMG_UYVY, F.IMG_Y8]
CFG2_DST = [F.IMG_RGB24, F.IMG_BGR24, F.IMG_RGBA32]


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_cpuinfo() & pkg.AC_CUDA, "no usable B200"
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def chk():
    return ck.best_checker()


def size_variants(srcfmt, dstfmt, w, h):
    uw, uh = F.size_unit(srcfmt, dstfmt)
    return [(w, h), (w - uw, h), (w, h - uh), (w - uw, h - uh)]


def assert_same(got, want, what):
    if not np.array_equal(got, want):
        bad = np.flatnonzero(got != want)
        raise AssertionError(f"{what}: {bad.size} bytes differ, first at {bad[0]} (got {got[bad[0]]}, want {want[bad[0]]})")


# ---- every pair, legacy host-pointer API, the reference test's four size variants -----------------
@pytest.mark.parametrize("srcfmt,dstfmt", ALL_PAIRS, ids=lambda f: F.NAMES[f])
def test_all_pairs_legacy_api_small(ac, chk, srcfmt, dstfmt):
    for (w, h) in size_variants(srcfmt, dstfmt, 64, 16):
        src = ck.random_frame(srcfmt, w, h, seed=1)
        ok, got = ac.convert(src, srcfmt, dstfmt, w, h)
        ok2, want = chk.convert(src, srcfmt, dstfmt, w, h)
        assert ok == 1 and ok2 == 1, ac.last_error()
        assert_same(got, want, f"{F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} @ {w}x{h}")


# ---- every pair, batched device-resident API on an aligned size (vectorised tiers) + forced tier 1 --
@pytest.mark.parametrize("srcfmt,dstfmt", ALL_PAIRS, ids=lambda f: F.NAMES[f])
def test_all_pairs_batched_aligned(ac, chk, srcfmt, dstfmt):
    w, h, nf = 128, 8, 3
    frames = np.stack([ck.random_frame(srcfmt, w, h, seed=20 + i) for i in range(nf)])
    dfb = F.frame_bytes(dstfmt, w, h)
    want = np.stack([chk.convert(frames[i], srcfmt, dstfmt, w, h, pad=0)[1] for i in range(nf)])
    tiers = [0, 1]
    if dstfmt in (F.IMG_RGB24, F.IMG_BGR24) and (srcfmt >> 12) == 1 and srcfmt != F.IMG_Y8:
        tiers.append(3)          # bulk-store variant exists for YUV -> 24-bit RGB
    for tier in tiers:
        ac.lib.acgpu_force_tier(tier)
        try:
            got = ac.convert_batch(frames, srcfmt, dstfmt, w, h, dst_pitch=dfb + 256)
        finally:
            ac.lib.acgpu_force_tier(0)
        assert_same(got[:, :dfb], want, f"batched {F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} tier {tier}")
        assert (got[:, dfb:] == 0x55).all(), "wrote into the inter-frame gap"
        if tier == 0:
            # aligned planes + sizes on the 16-pixel grid: every pair must be served by a vectorised tier
            assert ac.lib.acgpu_last_kernel_tier() >= 2, "fell back to the generic tier"


def test_unknown_pairs_and_degenerate_sizes(ac):
    src = np.zeros(4096, np.uint8)
    assert ac.convert(src, 0x1234, F.IMG_RGB24, 8, 8)[0] == 0          # imgconvert.c:63
    assert ac.convert(src, F.IMG_RGB24, 0, 8, 8)[0] == 0
    ok, d = ac.convert(src, F.IMG_YUV420P, F.IMG_RGB24, 0, 8)            # loops do not iterate
    assert ok == 1 and (d == 0x55).all()
    ok, d = ac.convert(src, F.IMG_YUV420P, F.IMG_RGB24, 8, 0)
    assert ok == 1 and (d == 0x55).all()


def test_golden_digests_through_libacgpu(ac):
    with open(os.path.join(os.path.dirname(__file__), "golden", "imgconvert_digests.json")) as f:
        gold = json.load(f)
    for key, want in gold["digests"].items():
        sname, dname, size = key.split(":")
        w, h = map(int, size.split("x"))
        sf, df = F.BY_NAME[sname], F.BY_NAME[dname]
        src = ck.random_frame(sf, w, h, seed=gold["seed"])
        ok, got = ac.convert(src, sf, df, w, h, prefill=0x55)
        assert ok == 1 and hashlib.sha256(got.tobytes()).hexdigest()[:16] == want, key


# ---- the reference test's own image and sizes (testsuite/test-imgconvert.c:180-224,361-363) ------------
@pytest.mark.parametrize("srcfmt,dstfmt", [
    (F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_RGB24, F.IMG_YUV422P), (F.IMG_YUY2, F.IMG_YUV420P),
    (F.IMG_YUV444P, F.IMG_BGRA32), (F.IMG_UYVY, F.IMG_YUV411P), (F.IMG_YUV411P, F.IMG_YVYU),
    (F.IMG_ARGB32, F.IMG_YV12), (F.IMG_GRAY8, F.IMG_YUV420P), (F.IMG_Y8, F.IMG_ABGR32),
    (F.IMG_YUV444P, F.IMG_YUV420P), (F.IMG_RGBA32, F.IMG_BGR24), (F.IMG_BGR24, F.IMG_GRAY8),
], ids=lambda f: F.NAMES[f])
def test_reference_test_image_768x512(ac, chk, srcfmt, dstfmt):
    img = ck.glibc_random_bytes(768 * 512 * 4, 0)
    for (w, h) in size_variants(srcfmt, dstfmt, 768, 512):
        src = img[: F.frame_bytes(srcfmt, w, h)]
        ok, got = ac.convert(src, srcfmt, dstfmt, w, h, prefill=0)
        _, want = chk.convert(src, srcfmt, dstfmt, w, h, prefill=0)
        assert ok == 1
        assert_same(got, want, f"{F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} @ {w}x{h}")


# ---- BASELINE config 2: the 36-pair matrix at 1920x1080, both directions ---------------------------------
@pytest.mark.parametrize("a,b", [(s, d) for s in CFG2_SRC for d in CFG2_DST], ids=lambda f: F.NAMES[f])
def test_config2_matrix_1080p(ac, chk, a, b):
    w, h = 1920, 1080
    for sf, df in ((a, b), (b, a)):
        src = ck.random_frame(sf, w, h, seed=5)
        got = ac.convert_batch(src[None, :], sf, df, w, h)[0]
        _, want = chk.convert(src, sf, df, w, h, pad=0)
        assert_same(got, want, f"1080p {F.NAMES[sf]}->{F.NAMES[df]}")


# ---- every pair at full size: the grid-stride paths (block clamp, several trips per block, frame segments) that the
# small all-pairs tests never reach (testsuite/test-imgconvert.c runs its matrix at one size only, 768x512) -------------
# PAL is BASELINE config 1's size (4:2:0 chroma rows of 360 bytes: not 16-byte aligned), UHD config 4's.
@pytest.mark.parametrize("size", [(1920, 1080), (1280, 720), (720, 576), (3840, 2160)], ids=lambda s: f"{s[0]}x{s[1]}")
@pytest.mark.parametrize("srcfmt", F.FORMATS_16, ids=lambda f: F.NAMES[f])
def test_all_pairs_full_size(ac, chk, srcfmt, size):
    w, h = size
    nf = 1 if w * h > 4_000_000 else 2
    frames = np.stack([ck.random_frame(srcfmt, w, h, seed=70 + i) for i in range(nf)])
    for dstfmt in F.FORMATS_16:
        dfb = F.frame_bytes(dstfmt, w, h)
        got = ac.convert_batch(frames, srcfmt, dstfmt, w, h, dst_pitch=dfb + 256)
        assert ac.lib.acgpu_last_kernel_tier() >= 2, "fell back to the generic tier"
        for i in range(nf):
            _, want = chk.convert(frames[i], srcfmt, dstfmt, w, h, pad=0)
            assert_same(got[i, :dfb], want, f"{w}x{h} {F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} frame {i}")
        assert (got[:, dfb:] == 0x55).all(), "wrote into the inter-frame gap"


def test_config1_pal_and_config4_uhd_round_trip(ac, chk):
    # config 1 size (PAL; chroma pitch 360 is not 16-byte aligned) and config 4 (UHD 420P->RGB24->422P)
    w, h = 720, 576
    src = ck.random_frame(F.IMG_YUV420P, w, h, seed=9)
    got = ac.convert_batch(src[None, :], F.IMG_YUV420P, F.IMG_RGB24, w, h)[0]
    assert_same(got, chk.convert(src, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)[1], "PAL 420P->RGB24")
    w, h = 3840, 2160
    src = ck.random_frame(F.IMG_YUV420P, w, h, seed=9)
    rgb = ac.convert_batch(src[None, :], F.IMG_YUV420P, F.IMG_RGB24, w, h)[0]
    assert_same(rgb, chk.convert(src, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)[1], "UHD 420P->RGB24")
    yuv = ac.convert_batch(rgb[None, :], F.IMG_RGB24, F.IMG_YUV422P, w, h)[0]
    assert_same(yuv, chk.convert(rgb, F.IMG_RGB24, F.IMG_YUV422P, w, h, pad=0)[1], "UHD RGB24->422P")


def test_config5_batch_64x720p_yuy2_to_420p(ac, chk):
    w, h, nf = 1280, 720, 64
    base = ck.random_frame(F.IMG_YUY2, w, h, seed=11)
    frames = np.stack([np.roll(base, 977 * i) for i in range(nf)])
    got = ac.convert_batch(frames, F.IMG_YUY2, F.IMG_YUV420P, w, h)
    for i in (0, 1, 31, 63):
        assert_same(got[i], chk.convert(frames[i], F.IMG_YUY2, F.IMG_YUV420P, w, h, pad=0)[1], f"720p frame {i}")
    # every frame is a rolled copy of frame 0's bytes, so checksums of all outputs must be distinct and stable
    sums = {hashlib.sha256(got[i].tobytes()).hexdigest() for i in range(nf)}
    assert len(sums) == nf


# ---- exhaustive cubes (SURVEY.md 8d iii) -------------------------------------------------------------------
def test_exhaustive_yuv_cube_to_rgb(ac, chk):
    """All 2^24 (Y,U,V) triples as a 4096x4096 YUV444P image -> RGB24 and -> BGRA32."""
    n = 1 << 24
    idx = np.arange(n, dtype=np.uint32)
    src = np.concatenate([(idx & 0xFF).astype(np.uint8), ((idx >> 8) & 0xFF).astype(np.uint8), (idx >> 16).astype(np.uint8)])
    w = h = 4096
    got = ac.convert_batch(src[None, :], F.IMG_YUV444P, F.IMG_RGB24, w, h)[0]
    assert_same(got, chk.convert(src, F.IMG_YUV444P, F.IMG_RGB24, w, h, pad=0)[1], "YUV cube -> RGB24")
    got = ac.convert_batch(src[None, :], F.IMG_YUV444P, F.IMG_BGRA32, w, h)[0]
    assert_same(got, chk.convert(src, F.IMG_YUV444P, F.IMG_BGRA32, w, h, pad=0)[1], "YUV cube -> BGRA32")


def test_exhaustive_rgb_cube_to_yuv(ac, chk):
    n = 1 << 24
    idx = np.arange(n, dtype=np.uint32)
    src = np.stack([(idx & 0xFF).astype(np.uint8), ((idx >> 8) & 0xFF).astype(np.uint8), (idx >> 16).astype(np.uint8)], axis=1).reshape(-1)
    w = h = 4096
    for df in (F.IMG_YUV444P, F.IMG_YUV420P, F.IMG_YUY2, F.IMG_GRAY8):
        got = ac.convert_batch(src[None, :], F.IMG_RGB24, df, w, h)[0]
        assert_same(got, chk.convert(src, F.IMG_RGB24, df, w, h, pad=0)[1], f"RGB cube -> {F.NAMES[df]}")


# ---- size-independent properties at full size ----------------------------------------------------------------
def test_properties_full_size(ac):
    w, h = 1920, 1080
    src = ck.random_frame(F.IMG_RGBA32, w, h, seed=2)
    # pure permutes are involutions / invertible: RGBA -> ABGR -> RGBA, RGBA -> BGRA -> RGBA
    for mid in (F.IMG_ABGR32, F.IMG_BGRA32, F.IMG_ARGB32):
        a = ac.convert_batch(src[None, :], F.IMG_RGBA32, mid, w, h)[0]
        b = ac.convert_batch(a[None, :], mid, F.IMG_RGBA32, w, h)[0]
        assert np.array_equal(b, src)
    # packed YUV byte orders round-trip, planar up-sampling then box down-sampling is the identity
    p = ck.random_frame(F.IMG_YUY2, w, h, seed=3)
    for mid in (F.IMG_UYVY, F.IMG_YVYU):
        a = ac.convert_batch(p[None, :], F.IMG_YUY2, mid, w, h)[0]
        assert np.array_equal(ac.convert_batch(a[None, :], mid, F.IMG_YUY2, w, h)[0], p)
    y = ck.random_frame(F.IMG_YUV420P, w, h, seed=4)
    up = ac.convert_batch(y[None, :], F.IMG_YUV420P, F.IMG_YUV444P, w, h)[0]
    assert np.array_equal(ac.convert_batch(up[None, :], F.IMG_YUV444P, F.IMG_YUV420P, w, h)[0], y)
    # YV12 is YUV420P with the chroma planes exchanged
    a = ac.convert_batch(y[None, :], F.IMG_YUV420P, F.IMG_RGB24, w, h)[0]
    p0, c = w * h, (w // 2) * (h // 2)
    yv = np.concatenate([y[:p0], y[p0 + c:], y[p0:p0 + c]])
    assert np.array_equal(ac.convert_batch(yv[None, :], F.IMG_YV12, F.IMG_RGB24, w, h)[0], a)
    # alpha stays whatever the destination held (img_yuv_rgb.c:62-64)
    out = ac.convert_batch(y[None, :], F.IMG_YUV420P, F.IMG_RGBA32, w, h, prefill=0xA7)[0]
    assert (out[3::4] == 0xA7).all() and np.array_equal(out.reshape(-1, 4)[:, :3].reshape(-1), a)


def test_in_place_packed_and_rgba_permutes(ac, chk):
    """img_yuv_packed.c:22,29,40 and img_rgb_packed.c:22,44: these work with src == dest."""
    w, h = 128, 16
    for sf, df in [(F.IMG_YUY2, F.IMG_UYVY), (F.IMG_YUY2, F.IMG_YVYU), (F.IMG_UYVY, F.IMG_YVYU), (F.IMG_YVYU, F.IMG_UYVY),
                   (F.IMG_RGBA32, F.IMG_ABGR32), (F.IMG_RGBA32, F.IMG_BGRA32), (F.IMG_ARGB32, F.IMG_RGBA32), (F.IMG_RGBA32, F.IMG_ARGB32),
                   (F.IMG_RGB24, F.IMG_BGR24), (F.IMG_BGR24, F.IMG_RGB24)]:    # src/video_trans.c:363-371 (-k) swaps in place
        src = ck.random_frame(sf, w, h, seed=6)
        buf = ac.malloc(src.size).upload(src)
        for tier in (0, 1):
            buf.upload(src)
            ac.lib.acgpu_force_tier(tier)
            try:
                ac._ok(ac.imgconvert_batch(buf.ptr, sf, src.size, buf.ptr, df, src.size, w, h, 1))
            finally:
                ac.lib.acgpu_force_tier(0)
            ac.sync()
            assert_same(buf.download(), chk.convert(src, sf, df, w, h, pad=0)[1], f"in-place {F.NAMES[sf]}->{F.NAMES[df]} tier {tier}")
        buf.free()


def test_uyvy_source_is_not_modified(ac):
    """Documented difference: the reference rewrites UYVY/YVYU sources (img_yuv_mixed.c:24,30-32); libacgpu does not."""
    w, h = 64, 16
    src = ck.random_frame(F.IMG_UYVY, w, h, seed=8)
    s = src.copy()
    d = np.zeros(F.frame_bytes(F.IMG_YUV420P, w, h), np.uint8)
    assert ac.ac_imgconvert(s, F.IMG_UYVY, d, F.IMG_YUV420P, w, h) == 1
    assert np.array_equal(s, src)


def test_odd_heights_and_widths_where_the_c_path_is_well_defined(ac, chk):
    cases = [(F.IMG_YUV422P, F.IMG_RGB24, 64, 15), (F.IMG_YUY2, F.IMG_YUV422P, 62, 7), (F.IMG_YUV444P, F.IMG_RGB24, 37, 11),
             (F.IMG_RGB24, F.IMG_YUV444P, 33, 9), (F.IMG_RGB24, F.IMG_GRAY8, 31, 5), (F.IMG_RGBA32, F.IMG_RGB24, 17, 3),
             (F.IMG_YUV444P, F.IMG_YUV422P, 64, 5), (F.IMG_YUV422P, F.IMG_YUV444P, 66, 3), (F.IMG_Y8, F.IMG_RGBA32, 13, 13),
             (F.IMG_YUV411P, F.IMG_YUV444P, 68, 5), (F.IMG_YUV422P, F.IMG_YUY2, 66, 5), (F.IMG_GRAY8, F.IMG_UYVY, 10, 3)]
    for sf, df, w, h in cases:
        src = ck.random_frame(sf, w, h, seed=12)
        ok, got = ac.convert(src, sf, df, w, h)
        _, want = chk.convert(src, sf, df, w, h)
        assert ok == 1
        assert_same(got, want, f"{F.NAMES[sf]}->{F.NAMES[df]} @ {w}x{h}")


def test_device_guard_bands_all_pairs(ac, chk):
    """Kernels must not write outside [dest, dest + frame_bytes): canaries before and after every device frame.
    (compute-sanitizer is closed on this GPU pool, so overruns are caught this way.)"""
    w, h, nf, guard = 128, 8, 2, 512
    for srcfmt, dstfmt in ALL_PAIRS:
        sfb, dfb = F.frame_bytes(srcfmt, w, h), F.frame_bytes(dstfmt, w, h)
        dpitch = dfb + guard
        frames = np.stack([ck.random_frame(srcfmt, w, h, seed=70 + i) for i in range(nf)])
        ds = ac.malloc(nf * sfb).upload(frames.reshape(-1))
        dd = ac.malloc(guard + nf * dpitch).fill(0xC3)
        for tier in (0, 1):
            ac.lib.acgpu_force_tier(tier)
            try:
                ac._ok(ac.imgconvert_batch(ds.ptr, srcfmt, sfb, dd.ptr + guard, dstfmt, dpitch, w, h, nf))
            finally:
                ac.lib.acgpu_force_tier(0)
            ac.sync()
            out = dd.download()
            assert (out[:guard] == 0xC3).all(), f"underrun {F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} tier {tier}"
            body = out[guard:].reshape(nf, dpitch)
            assert (body[:, dfb:] == 0xC3).all(), f"overrun {F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} tier {tier}"
        ds.free(); dd.free()


def test_concurrent_callers_like_transcode_frame_threads(ac, chk):
    """src/frame_threads.c:174-228: N frame threads hit ac_imgconvert / ac_average concurrently and re-entrantly.
    Every thread gets its own stream and staging buffers inside libacgpu; results must not mix."""
    import threading
    w, h, rounds = 320, 240, 6
    pairs = [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_RGB24, F.IMG_YUV420P), (F.IMG_YUY2, F.IMG_YUV422P), (F.IMG_YUV422P, F.IMG_BGRA32),
             (F.IMG_UYVY, F.IMG_RGB24), (F.IMG_RGBA32, F.IMG_YUY2), (F.IMG_YUV444P, F.IMG_YUV420P), (F.IMG_GRAY8, F.IMG_RGB24)]
    want = {}
    for t, (sf, df) in enumerate(pairs):
        for r in range(rounds):
            src = ck.random_frame(sf, w, h, seed=1000 + 10 * t + r)
            want[(t, r)] = (src, chk.convert(src, sf, df, w, h)[1])
    errors = []

    def worker(t):
        sf, df = pairs[t]
        mine = pkg.AcGpu()          # same library, this thread's own thread-local context
        try:
            for r in range(rounds):
                src, exp = want[(t, r)]
                ok, got = mine.convert(src, sf, df, w, h)
                if ok != 1 or not np.array_equal(got, exp):
                    errors.append((t, r, "convert"))
                a, b = src[:4096].copy(), src[4096:8192].copy()
                if not np.array_equal(mine.ac_average(a, b), ((a.astype(np.int32) + b + 1) // 2).astype(np.uint8)):
                    errors.append((t, r, "average"))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(pairs))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


def test_rows_wider_than_one_block(ac, chk):
    """8K-class widths: rows are cut into column segments of 256 units (grid.z) in the 4:2:0 kernels."""
    w, h = 7680 + 32, 6          # 482 units per row: one full segment + a partial one with a ragged last warp
    for sf, df in [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_YUV420P, F.IMG_BGRA32), (F.IMG_RGB24, F.IMG_YUV420P),
                   (F.IMG_YUY2, F.IMG_YUV420P), (F.IMG_YUV420P, F.IMG_UYVY), (F.IMG_YUV444P, F.IMG_YUV420P),
                   (F.IMG_YUV420P, F.IMG_YUV422P)]:
        src = ck.random_frame(sf, w, h, seed=77)
        got = ac.convert_batch(np.stack([src, src[::-1].copy()]), sf, df, w, h)
        assert ac.lib.acgpu_last_kernel_tier() in (2, 3)      # 3: the tensor-map staged form of YUV420P -> RGB24 (segments too)
        assert_same(got[0], chk.convert(src, sf, df, w, h, pad=0)[1], f"wide {F.NAMES[sf]}->{F.NAMES[df]}")
        assert_same(got[1], chk.convert(src[::-1].copy(), sf, df, w, h, pad=0)[1], f"wide {F.NAMES[sf]}->{F.NAMES[df]} #2")


def test_full_8k_frames(ac, chk):
    """7680x4320: 33 M pixels per frame (132 MB as 32-bit RGB) -- the largest size of the sweeps, compared byte for byte."""
    w, h = 7680, 4320
    for sf, df in [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_RGB24, F.IMG_YUV422P), (F.IMG_YUY2, F.IMG_YUV420P), (F.IMG_YUV444P, F.IMG_BGRA32)]:
        src = ck.random_frame(sf, w, h, seed=88)
        got = ac.convert_batch(src[None, :], sf, df, w, h)[0]
        assert ac.lib.acgpu_last_kernel_tier() >= 2
        assert_same(got, chk.convert(src, sf, df, w, h, pad=0)[1], f"8K {F.NAMES[sf]}->{F.NAMES[df]}")


@pytest.mark.parametrize("mode", ["1", "2"])
def test_bulk_async_staged_variants_are_bit_exact(mode):
    """Tier 3b/3c (ACGPU_TMA=1/2: planar sources staged by cp.async.bulk + mbarrier, optionally bulk stores too) are
    measured slower than tier 2 and not the default, but they stay selectable, so they stay tested.  The mode is read
    once per process, hence the subprocess."""
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import __graft_entry__ as e, checkers as ck
F = ck.F
pkg = e.load_package(); ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_CUDA) == 1
chk = ck.best_checker()
ac.lib.acgpu_force_tier(3)
for (w, h, nf) in ((1920, 1080, 3), (1280, 720, 2), (64, 6, 5), (4128, 4, 2)):
    for df in (F.IMG_RGB24, F.IMG_BGR24):
        frames = np.stack([ck.random_frame(F.IMG_YUV420P, w, h, seed=300 + i) for i in range(nf)])
        got = ac.convert_batch(frames, F.IMG_YUV420P, df, w, h)
        assert ac.lib.acgpu_last_kernel_tier() == 3
        for i in range(nf):
            assert np.array_equal(got[i], chk.convert(frames[i], F.IMG_YUV420P, df, w, h, pad=0)[1]), (w, h, df, i)
print("ok")
""" % (entry.ROOT, os.path.join(entry.ROOT, "tests"))
    env = dict(os.environ, ACGPU_TMA=mode)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_ac_init_selects_and_deselects_the_cuda_path(ac):
    """aclib/accore.c:29-40 semantics with the new bit: the accel word is masked with ac_cpuinfo(); without AC_CUDA
    there is nothing to run (libacgpu has no CPU implementation), and ac_init may be called again to switch back."""
    src = np.zeros(F.frame_bytes(F.IMG_YUV420P, 64, 16), np.uint8)
    assert ac.ac_cpuinfo() == pkg.AC_CUDA
    try:
        assert ac.ac_init(pkg.AC_NONE) == 0
        assert "no CPU fallback" in ac.last_error()
        assert ac.convert(src, F.IMG_YUV420P, F.IMG_RGB24, 64, 16)[0] == 0       # like the reference before ac_init
        assert ac.ac_init(pkg.AC_SSE2) == 0                                       # an x86 bit alone selects nothing here
    finally:
        assert ac.ac_init(pkg.AC_ALL) == 1
    assert ac.convert(src, F.IMG_YUV420P, F.IMG_RGB24, 64, 16)[0] == 1
    assert ac.ac_init(pkg.AC_CUDA | pkg.AC_SSE2) == 1


def test_colour_bars_round_trip_exactly_on_the_gpu(ac):
    """testsuite/newtest.pl:544-566 (test_raw_raw_csp): YUV420P bars <-> RGB24 bars must be exact both ways."""
    for (w, h) in [(704, 576), (720, 480), (1920, 1080), (64, 64)]:
        yuv, rgb = ck.colour_bars_yuv420p(w, h), ck.colour_bars_rgb24(w, h)
        assert np.array_equal(ac.convert_batch(yuv[None, :], F.IMG_YUV420P, F.IMG_RGB24, w, h)[0], rgb)
        assert np.array_equal(ac.convert_batch(rgb[None, :], F.IMG_RGB24, F.IMG_YUV420P, w, h)[0], yuv)
        ok, got = ac.convert(yuv, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)
        assert ok == 1 and np.array_equal(got, rgb)


def test_flat_420_mode_sizes(ac, chk):
    """Widths whose 16-pixel units do not fill whole warps use the flat 4:2:0 walk (warps straddle row pairs and the
    transposed store writes two segments): PAL 720, 704, 640, 800; plus ragged last warps."""
    for (w, h, nf) in [(720, 576, 2), (704, 480, 1), (640, 34, 3), (800, 6, 2), (720, 4, 1), (528, 4, 2), (1296, 12, 1), (2000, 16, 2)]:
        for df in (F.IMG_RGB24, F.IMG_BGR24, F.IMG_RGBA32, F.IMG_ABGR32):
            frames = np.stack([ck.random_frame(F.IMG_YUV420P, w, h, seed=400 + i) for i in range(nf)])
            got = ac.convert_batch(frames, F.IMG_YUV420P, df, w, h, prefill=0x3C)
            # 24-bit destinations take the flat tensor-map staged form (tier 3) from a warp of units per row on
            assert ac.lib.acgpu_last_kernel_tier() == (3 if df in (F.IMG_RGB24, F.IMG_BGR24) and w >= 512 else 2)
            for i in range(nf):
                want = chk.convert(frames[i], F.IMG_YUV420P, df, w, h, prefill=0x3C, pad=0)[1]
                assert_same(got[i], want, f"flat420 {w}x{h} -> {F.NAMES[df]} frame {i}")


RGB_ALL = (F.IMG_RGB24, F.IMG_BGR24, F.IMG_RGBA32, F.IMG_ABGR32, F.IMG_ARGB32, F.IMG_BGRA32)


def test_ragged_420_widths_use_the_vectorised_tier(ac, chk):
    """4:2:0 frames whose width is not a multiple of 16 (854x480, 1080-wide portrait, the reference test's 766):
    rows are no longer 16-byte aligned, so YUV420P <-> RGB walks flat 16-pixel units and reaches the chroma rows with
    byte-aligned accesses (S420R / D420R); units that run over the end of a row fetch/store sample by sample."""
    # the last two: widths on the 16-pixel grid whose V plane is still not 16-byte aligned (width*height % 64 != 0)
    sizes = [(854, 480, 1), (1080, 40, 2), (766, 32, 1), (40, 8, 3), (50, 16, 2), (18, 16, 2), (24, 2, 1), (426, 240, 1),
             (48, 10, 1), (1936, 10, 1)]
    for (w, h, nf) in sizes:
        for rf in RGB_ALL:
            for sf, df in ((F.IMG_YUV420P, rf), (rf, F.IMG_YUV420P), (F.IMG_YV12, rf), (rf, F.IMG_YV12)):
                if (sf == F.IMG_YV12 or df == F.IMG_YV12) and rf not in (F.IMG_RGB24, F.IMG_ARGB32):
                    continue
                frames = np.stack([ck.random_frame(sf, w, h, seed=900 + i) for i in range(nf)])
                got = ac.convert_batch(frames, sf, df, w, h, prefill=0xA5)
                assert ac.lib.acgpu_last_kernel_tier() == 2, f"{F.NAMES[sf]}->{F.NAMES[df]} @ {w}x{h} fell back"
                for i in range(nf):
                    want = chk.convert(frames[i], sf, df, w, h, prefill=0xA5, pad=0)[1]
                    assert_same(got[i], want, f"ragged {F.NAMES[sf]}->{F.NAMES[df]} @ {w}x{h} frame {i}")


def test_ragged_420_widths_yuv_family(ac, chk):
    """The same ragged widths for 4:2:0 <-> 4:2:2 / 4:4:4 / 4:1:1 / packed / Y8 / GRAY8 / 4:2:0 (Ragged420From / To):
    luma and the other format stay on aligned flat units, the 4:2:0 chroma rows are reached bytewise; a unit that wraps
    from an odd row into an even one contributes only its tail to the vertical chroma mean."""
    others = [F.IMG_YUV422P, F.IMG_YUV444P, F.IMG_YUV411P, F.IMG_YUY2, F.IMG_UYVY, F.IMG_YVYU, F.IMG_Y8, F.IMG_GRAY8,
              F.IMG_YUV420P, F.IMG_YV12]
    sizes = [(854, 480, 1), (1080, 40, 2), (766, 32, 1), (40, 8, 3), (50, 16, 2), (18, 16, 2), (24, 2, 1), (426, 240, 1), (20, 4, 2),
             (48, 10, 1), (1936, 10, 1)]
    for (w, h, nf) in sizes:
        for o in others:
            for sf, df in ((F.IMG_YUV420P, o), (o, F.IMG_YUV420P)):
                frames = np.stack([ck.random_frame(sf, w, h, seed=950 + i) for i in range(nf)])
                got = ac.convert_batch(frames, sf, df, w, h, prefill=0x5A)
                # the OTHER format's planes still have to be 16-byte aligned (tiny frames: 4:2:2 V of 24x2 is not)
                aligned = all(off % 16 == 0 for off in F.plane_offsets(o, w, h)) or o in (F.IMG_YUV420P, F.IMG_YV12)
                aligned = aligned and (nf == 1 or (F.frame_bytes(sf, w, h) % 16 == 0 and F.frame_bytes(df, w, h) % 16 == 0))
                if aligned and not (o == F.IMG_YUV411P and w % 4):
                    assert ac.lib.acgpu_last_kernel_tier() == 2, f"{F.NAMES[sf]}->{F.NAMES[df]} @ {w}x{h} fell back"
                for i in range(nf):
                    want = chk.convert(frames[i], sf, df, w, h, prefill=0x5A, pad=0)[1]
                    assert_same(got[i], want, f"ragged {F.NAMES[sf]}->{F.NAMES[df]} @ {w}x{h} frame {i}")


def test_legacy_call_with_every_mix_of_pointer_kinds(ac, chk):
    """ac_imgconvert(src planes, ..., dest planes, ...) with the source and the destination independently in pageable
    host memory, page-locked host memory (acgpu_host_alloc) or device memory: libacgpu classifies the pointers per call
    (host planes are staged, device planes are used in place), the bytes are the same in all nine combinations."""
    import ctypes as C

    def place(kind, data):
        if kind == "pageable":
            arr = np.array(data, dtype=np.uint8, copy=True)
            return arr, arr.ctypes.data, (lambda: arr.copy()), (lambda: None)
        if kind == "pinned":
            pb = ac.pinned(data.size)
            pb.array[:] = data
            return pb, pb.ptr, (lambda: pb.array.copy()), pb.free
        db = ac.malloc(data.size).upload(data)
        return db, db.ptr, db.download, db.free

    for (w, h) in [(64, 16), (50, 16), (30, 6)]:
        for sf, df in [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_BGRA32, F.IMG_YUV422P), (F.IMG_UYVY, F.IMG_YUV444P), (F.IMG_YUV420P, F.IMG_ARGB32)]:
            src = ck.random_frame(sf, w, h, seed=31)
            want = chk.convert(src, sf, df, w, h, prefill=0x6B, pad=0)[1]
            so, do = F.plane_offsets(sf, w, h), F.plane_offsets(df, w, h)
            for sk in ("pageable", "pinned", "device"):
                for dk in ("pageable", "pinned", "device"):
                    _s, sptr, _sread, sfree = place(sk, src)
                    _d, dptr, dread, dfree = place(dk, np.full(want.size, 0x6B, np.uint8))
                    sp = (C.c_void_p * 3)(*[sptr + o for o in so] + [None] * (3 - len(so)))
                    dp = (C.c_void_p * 3)(*[dptr + o for o in do] + [None] * (3 - len(do)))
                    assert ac.lib.ac_imgconvert(sp, sf, dp, df, w, h) == 1, (sk, dk, ac.last_error())
                    ac.sync()
                    assert_same(dread(), want, f"{F.NAMES[sf]}->{F.NAMES[df]} {w}x{h} src {sk} dest {dk}")
                    sfree(); dfree()
