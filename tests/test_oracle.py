"""Pins the oracle (oracle/ac_oracle.c) before anything else trusts it.  CPU only.

Three anchors:
  1. the unmodified reference compiled here (oracle/_ref/libac_ref_c.so, ac_init(AC_NONE)): every one
     of the 256 reachable pairs, at the four size variants the reference's own test uses
     (testsuite/test-imgconvert.c:180-224), byte-for-byte including untouched bytes and guard bands;
  2. the reference's known-answer vectors: test-average.c:177-645 (135 entries, 0x11 guard bands) and
     the newtest.pl colour bars that must round-trip exactly (newtest.pl:544-566,1462-1537);
  3. committed digests generated from oracle/_ref (tests/golden/make_golden.py) so the pin still holds
     on a machine where the reference was never built.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import checkers as ck
from checkers import F

ORACLE = ck.Oracle()
HAVE_REF = ck.have_ref("c")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "imgconvert_digests.json")

BASE_W, BASE_H = 64, 16


def size_variants(srcfmt, dstfmt, w=BASE_W, h=BASE_H):
    uw, uh = F.size_unit(srcfmt, dstfmt)
    return [(w, h), (w - uw, h), (w, h - uh), (w - uw, h - uh)]


ALL_PAIRS = [(s, d) for s in F.FORMATS_16 for d in F.FORMATS_16]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (no /root/reference on this machine)")
@pytest.mark.parametrize("srcfmt,dstfmt", ALL_PAIRS, ids=lambda f: F.NAMES[f])
def test_oracle_matches_reference_all_pairs(srcfmt, dstfmt):
    ref = ck.RefLib("c")
    for (w, h) in size_variants(srcfmt, dstfmt):
        src = ck.random_frame(srcfmt, w, h, seed=1)
        ok_r, want = ref.convert(src, srcfmt, dstfmt, w, h, prefill=0x55)
        ok_o, got = ORACLE.convert(src, srcfmt, dstfmt, w, h, prefill=0x55)
        assert ok_r == 1 and ok_o == 1
        assert np.array_equal(want, got), f"{F.NAMES[srcfmt]}->{F.NAMES[dstfmt]} @ {w}x{h}"


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("srcfmt,dstfmt", [
    (F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_RGB24, F.IMG_YUV422P), (F.IMG_YUY2, F.IMG_YUV420P),
    (F.IMG_YUV444P, F.IMG_BGRA32), (F.IMG_UYVY, F.IMG_YUV411P), (F.IMG_YUV411P, F.IMG_YVYU),
    (F.IMG_ARGB32, F.IMG_YV12), (F.IMG_GRAY8, F.IMG_YUV420P), (F.IMG_Y8, F.IMG_ABGR32),
])
def test_oracle_matches_reference_768x512_glibc_random(srcfmt, dstfmt):
    """Same image the reference test converts: srandom(0), low byte of random(), 768x512."""
    w, h = 768, 512
    src = ck.glibc_random_bytes(w * h * 4, 0)[: F.frame_bytes(srcfmt, w, h)]
    ref = ck.RefLib("c")
    _, want = ref.convert(src, srcfmt, dstfmt, w, h, prefill=0)
    _, got = ORACLE.convert(src, srcfmt, dstfmt, w, h, prefill=0)
    assert np.array_equal(want, got)


def test_oracle_rejects_unknown_formats():
    src = np.zeros(64, dtype=np.uint8)
    assert ORACLE.convert(src, 0x1234, F.IMG_RGB24, 4, 4)[0] == 0
    assert ORACLE.convert(src, F.IMG_RGB24, 0, 4, 4)[0] == 0


def test_colour_bars_round_trip_exactly():
    """newtest.pl test_raw_raw_csp: YUV420P bars <-> RGB24 bars must be exact in both directions."""
    for (w, h) in [(704, 576), (720, 480), (1920, 1080), (64, 64)]:
        yuv = ck.colour_bars_yuv420p(w, h)
        rgb = ck.colour_bars_rgb24(w, h)
        ok, got = ORACLE.convert(yuv, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)
        assert ok == 1 and np.array_equal(got, rgb)
        ok, got = ORACLE.convert(rgb, F.IMG_RGB24, F.IMG_YUV420P, w, h, pad=0)
        assert ok == 1 and np.array_equal(got, yuv)


def test_known_colour_points():
    """The seven exact RGB<->YUV pairs of newtest.pl:1449-1458."""
    pts = [((16, 128, 128), (0, 0, 0)), ((89, 128, 128), (85, 85, 85)), ((162, 128, 128), (170, 170, 170)),
           ((235, 128, 128), (255, 255, 255)), ((81, 91, 239), (253, 0, 1)), ((145, 54, 35), (2, 255, 1)),
           ((41, 240, 111), (2, 0, 255))]
    for yuv, rgb in pts:
        s = np.array(yuv, dtype=np.uint8)
        _, got = ORACLE.convert(s, F.IMG_YUV444P, F.IMG_RGB24, 1, 1, pad=0)
        assert tuple(got) == rgb
        _, got = ORACLE.convert(np.array(rgb, dtype=np.uint8), F.IMG_RGB24, F.IMG_YUV444P, 1, 1, pad=0)
        assert tuple(got) == yuv


def average_vectors():
    """test-average.c:177-645 regenerated: for sizes 1/8/32/64 every pair of constant fills drawn from
    {00,01,02,03}^2 and {7F,80,FE,FF}^2; plus the k / 2k+1 ramps for sizes 7,8,15,31,32,63,127."""
    out = []
    for size in (1, 8, 32, 64):
        for group in ((0x00, 0x01, 0x02, 0x03), (0x7F, 0x80, 0xFE, 0xFF)):
            for a in group:
                for b in group:
                    out.append((np.full(size, a, np.uint8), np.full(size, b, np.uint8)))
    for size in (7, 8, 15, 31, 32, 63, 127):
        k = np.arange(size)
        out.append(((k & 0xFF).astype(np.uint8), ((2 * k + 1) & 0xFF).astype(np.uint8)))
    assert len(out) == 135
    return out


def test_average_known_answers_with_guard_bands():
    for a, b in average_vectors():
        want = ((a.astype(np.int32) + b + 1) // 2).astype(np.uint8)
        got = ORACLE.average(a, b)
        assert np.array_equal(got, want)
    # guard bands: the oracle must not write outside [0, n)
    n, spill = 127, 8
    a, b = average_vectors()[-1]
    buf = np.full(n + 2 * spill, 0x11, np.uint8)
    ORACLE._average(ck._ptr(a), ck._ptr(b), ck._ptr(buf, spill), n)
    assert (buf[:spill] == 0x11).all() and (buf[-spill:] == 0x11).all()


def test_rescale_edge_semantics():
    """SURVEY.md Appendix A8 probes of the C path (rescale.c:23-46)."""
    a = np.array([255, 255, 0, 200], np.uint8)
    b = np.array([255, 0, 255, 100], np.uint8)
    cases = {
        (32768, 32768): [255, 128, 128, 150],
        (65535, 65535): [254, 255, 255, 44],      # sum wraps mod 256, no saturation
        (40000, 40000): [55, 156, 156, 183],
        (65536, 0): [255, 255, 0, 200],           # copy of src1, src2 never read
        (70000, 5): [255, 255, 0, 200],
        (0, 65536): [255, 0, 255, 100],
        (1, 65535): [255, 0, 255, 100],
    }
    for (w1, w2), want in cases.items():
        assert list(ORACLE.rescale(a, b, w1, w2)) == want, (w1, w2)
    assert list(ORACLE.average(a, b)) == [255, 128, 128, 150]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_rescale_average_match_reference_exhaustive_pairs():
    ref = ck.RefLib("c")
    a = np.repeat(np.arange(256, dtype=np.uint8), 256)
    b = np.tile(np.arange(256, dtype=np.uint8), 256)
    assert np.array_equal(ref.average(a, b), ORACLE.average(a, b))
    for w1, w2 in [(0, 65536), (1, 65535), (32767, 32769), (32768, 32768), (65535, 1), (65536, 0),
                   (49152, 16384), (16384, 49152), (40000, 40000), (65535, 65535), (12345, 54321), (70000, 5)]:
        assert np.array_equal(ref.rescale(a, b, w1, w2), ORACLE.rescale(a, b, w1, w2)), (w1, w2)


def test_resize_table_matches_survey_examples():
    """tcvideo.c:1138-1165: 1080->720 alternates (49152,16384)/(16384,49152) with source=floor(1.5 i)."""
    s, w1, w2 = ORACLE.resize_table(1080, 720)
    assert len(s) == 90
    assert list(s[:4]) == [0, 1, 3, 4]
    assert list(w1[:4]) == [49152, 16384, 49152, 16384] and list(w2[:4]) == [16384, 49152, 16384, 49152]
    s, w1, w2 = ORACLE.resize_table(720, 1080)
    assert w1[0] == 65536 and w2[0] == 0 and w1[1] == 32768


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def test_oracle_matches_committed_golden_digests():
    """Digests of the REFERENCE's outputs (generated by tests/golden/make_golden.py from oracle/_ref)."""
    with open(GOLDEN) as f:
        gold = json.load(f)
    assert gold["generator"].startswith("oracle/_ref/libac_ref_c.so")
    n = 0
    for key, want in gold["digests"].items():
        sname, dname, size = key.split(":")
        w, h = map(int, size.split("x"))
        sf, df = F.BY_NAME[sname], F.BY_NAME[dname]
        src = ck.random_frame(sf, w, h, seed=gold["seed"])
        ok, got = ORACLE.convert(src, sf, df, w, h, prefill=0x55)
        assert ok == 1 and digest(got) == want, key
        n += 1
    assert n == 256 * 2


# ---- libtcvideo-shaped plane operations -----------------------------------------------
import tcv_cases  # noqa: E402

TCV_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tcv_digests.json")


def test_oracle_tcv_ops_match_committed_reference_digests():
    """Digests of the REFERENCE libtcvideo's outputs (tests/golden/make_golden_tcv.py, from oracle/_ref/libtcv_ref.so)."""
    with open(TCV_GOLDEN) as f:
        gold = json.load(f)
    assert gold["generator"].startswith("oracle/_ref/libtcv_ref.so")
    all_cases = tcv_cases.cases()
    assert len(all_cases) == len(gold["digests"])
    for case in all_cases:
        want_ok, want = gold["digests"][case[0]]
        ok, d = tcv_cases.run_case(ORACLE, case)
        assert ok == want_ok, case[0]
        if ok:
            assert digest(d) == want, case[0]


@pytest.mark.skipif(not ck.have_tcv_ref(), reason="oracle/_ref/libtcv_ref.so not built")
def test_oracle_tcv_ops_match_reference_libtcvideo_directly():
    ref = ck.TcvRef()
    for case in tcv_cases.cases():
        ok_r, d_r = tcv_cases.run_case(ref, case)
        ok_o, d_o = tcv_cases.run_case(ORACLE, case)
        assert ok_r == ok_o, case[0]
        if ok_r:
            assert np.array_equal(d_r, d_o), case[0]


@pytest.mark.skipif(not ck.have_tcv_ref(), reason="oracle/_ref/libtcv_ref.so not built")
def test_reference_resize_table_via_1080_to_720_rows():
    """config 3 (iii): tcv_resize(…, 0, -45, 8, 8) on a 1080-row plane, reference vs restatement."""
    ref = ck.TcvRef()
    w, h = 64, 1080
    src = ck.splitmix_bytes(w * h, 3)
    assert np.array_equal(ref.resize(src, w, h, 1, 0, -45, 8, 8), ORACLE.resize(src, w, h, 1, 0, -45, 8, 8))
    assert np.array_equal(ref.resize(src[: w * 720], w, 720, 1, 0, 45, 8, 8), ORACLE.resize(src[: w * 720], w, 720, 1, 0, 45, 8, 8))


def test_antialias_cases_actually_smooth_something():
    """Guards the case list: the blocky images must exercise the weighted 3x3 branch (tcvideo.c:950-966)."""
    changed = 0
    for case in tcv_cases.cases():
        if case[1] == "antialias" and case[4] == "blocky" and case[3] == (0.333, 0.5):
            _, _, (w, h, bpp), _, kind, seed = case
            ok, d = tcv_cases.run_case(ORACLE, case)
            changed += int((d != tcv_cases.image(kind, w, h, bpp, seed)).sum())
    assert changed > 500
