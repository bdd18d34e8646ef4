"""-m gpu parity for the "next" rows of SURVEY.md 8f that are built: device-resident tcv_convert
(libtcvideo/tcvideo.c:1001-1067) and the fused -K grayscale of src/video_trans.c:381-388."""
import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck
from checkers import F

pkg = entry.load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def chk():
    return ck.best_checker()


@pytest.mark.parametrize("size", [(1920, 1080), (128, 16), (50, 7), (766, 512)])
def test_decolor_rgb24_equals_the_two_reference_conversions(ac, chk, size):
    w, h = size
    nf = 2
    fb = w * h * 3
    frames = np.stack([ck.splitmix_bytes(fb, 90 + i) for i in range(nf)])
    want = []
    for i in range(nf):
        _, g = chk.convert(frames[i], F.IMG_RGB24, F.IMG_GRAY8, w, h, pad=0)
        _, r = chk.convert(g, F.IMG_GRAY8, F.IMG_RGB24, w, h, pad=0)
        want.append(r)
    pitch = fb + 256 - fb % 256 if fb % 256 else fb + 256
    buf = ac.malloc(nf * pitch).fill(0x55)
    for i in range(nf):
        buf.upload(frames[i], offset=i * pitch)
    ac._ok(ac.lib.acgpu_decolor_rgb24_batch(buf.ptr, w, h, pitch, nf, None))
    ac.sync()
    out = buf.download().reshape(nf, pitch)
    for i in range(nf):
        assert np.array_equal(out[i, :fb], want[i]), (size, i)
    assert (out[:, fb:] == 0x55).all()
    buf.free()


def test_convert_batch_matches_tcv_convert_semantics(ac, chk):
    w, h, nf = 128, 16, 3
    # different buffers
    for sf, df in [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_YV12, F.IMG_BGRA32), (F.IMG_RGB24, F.IMG_YUV422P), (F.IMG_UYVY, F.IMG_YUV420P)]:
        sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
        frames = np.stack([ck.random_frame(sf, w, h, seed=30 + i) for i in range(nf)])
        ds = ac.malloc(nf * sfb).upload(frames.reshape(-1))
        dd = ac.malloc(nf * dfb).fill(0x55)
        ac._ok(ac.lib.acgpu_convert_batch(ds.ptr, dd.ptr, w, h, sf, df, sfb, dfb, nf, None))
        ac.sync()
        got = dd.download().reshape(nf, dfb)
        for i in range(nf):
            assert np.array_equal(got[i], chk.convert(frames[i], sf, df, w, h, pad=0)[1]), (sf, df, i)
        ds.free(); dd.free()
    # in place (src == dest): converted through a temporary like tcvideo.c:1044-1064
    sf, df = F.IMG_RGB24, F.IMG_YUV420P
    sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
    frames = np.stack([ck.random_frame(sf, w, h, seed=40 + i) for i in range(nf)])
    buf = ac.malloc(nf * sfb).upload(frames.reshape(-1))
    ac._ok(ac.lib.acgpu_convert_batch(buf.ptr, buf.ptr, w, h, sf, df, sfb, sfb, nf, None))
    ac.sync()
    got = buf.download().reshape(nf, sfb)
    for i in range(nf):
        assert np.array_equal(got[i, :dfb], chk.convert(frames[i], sf, df, w, h, pad=0)[1])
        assert np.array_equal(got[i, dfb:], frames[i, dfb:])         # bytes beyond the new frame are untouched
    # same format: plain copy
    dd = ac.malloc(nf * sfb).fill(0)
    ac._ok(ac.lib.acgpu_convert_batch(buf.ptr, dd.ptr, w, h, sf, sf, sfb, sfb, nf, None))
    ac.sync()
    assert np.array_equal(dd.download(), buf.download())
    assert ac.lib.acgpu_convert_batch(buf.ptr, dd.ptr, w, h, 0, sf, sfb, sfb, nf, None) == 0
    buf.free(); dd.free()


@pytest.mark.parametrize("case", [
    (F.IMG_YUV420P, F.IMG_RGB24, 1920, 1080, 9),      # several pipeline chunks (16 MB each)
    (F.IMG_YUV420P, F.IMG_RGBA32, 320, 240, 5),       # alpha must survive the host round trip (dest is pre-loaded)
    (F.IMG_YUY2, F.IMG_YUV420P, 1280, 720, 7),
    (F.IMG_RGB24, F.IMG_YUV422P, 766, 512, 3),        # frame size not a multiple of 256: planes unaligned on device
    (F.IMG_YV12, F.IMG_BGR24, 64, 16, 33),
])
def test_frames_host_pipeline(ac, chk, case):
    """acgpu_imgconvert_frames_host: the end-to-end call bench.py times (pinned buffers, 3-slot pipeline)."""
    sf, df, w, h, nf = case
    sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
    hs, hd = ac.pinned(nf * sfb), ac.pinned(nf * dfb)
    frames = [ck.random_frame(sf, w, h, seed=200 + i) for i in range(nf)]
    for i in range(nf):
        hs.array[i * sfb:(i + 1) * sfb] = frames[i]
    hd.array[:] = 0x5A
    ac._ok(ac.lib.acgpu_imgconvert_frames_host(hs.ptr, sf, hd.ptr, df, w, h, nf))
    for i in range(nf):
        want = chk.convert(frames[i], sf, df, w, h, prefill=0x5A, pad=0)[1]
        assert np.array_equal(hd.array[i * dfb:(i + 1) * dfb], want), (case, i)
    # pageable buffers work too (slower: the driver stages them)
    src = np.concatenate(frames)
    dst = np.full(nf * dfb, 0x5A, np.uint8)
    ac._ok(ac.lib.acgpu_imgconvert_frames_host(src.ctypes.data, sf, dst.ctypes.data, df, w, h, nf))
    assert np.array_equal(dst, np.asarray(hd.array))
    hs.free(); hd.free()


def test_in_place_converts_on_two_streams_share_the_temporary_safely(ac, chk):
    """src == dest goes through one per-thread temporary (tcvideo.c:1044-1064 allocates its own).  Two such calls on
    DIFFERENT caller streams, back to back without a sync, must not overlap on that temporary: the second waits for
    the first (event), both results are right."""
    w, h, nf = 1920, 1080, 24
    s1, s2 = ac.lib.acgpu_stream_create(), ac.lib.acgpu_stream_create()
    jobs = []
    for k, (sf, df) in enumerate([(F.IMG_RGB24, F.IMG_YUV420P), (F.IMG_RGB24, F.IMG_YUV422P)]):
        sfb = F.frame_bytes(sf, w, h)
        one = ck.random_frame(sf, w, h, seed=500 + k)
        buf = ac.malloc(nf * sfb)
        for i in range(nf):
            buf.upload(one, offset=i * sfb)
        jobs.append((sf, df, sfb, one, buf))
    for rep in range(3):
        for (sf, df, sfb, one, buf), st in zip(jobs, (s1, s2)):
            for i in range(nf):
                buf.upload(one, offset=i * sfb)
        ac.sync()
        for (sf, df, sfb, one, buf), st in zip(jobs, (s1, s2)):
            ac._ok(ac.lib.acgpu_convert_batch(buf.ptr, buf.ptr, w, h, sf, df, sfb, sfb, nf, st))
        ac.sync(s1); ac.sync(s2)
        for (sf, df, sfb, one, buf) in jobs:
            dfb = F.frame_bytes(df, w, h)
            want = chk.convert(one, sf, df, w, h, pad=0)[1]
            got = buf.download().reshape(nf, sfb)
            for i in range(nf):
                assert np.array_equal(got[i, :dfb], want), (rep, F.NAMES[df], i)
    for (_, _, _, _, buf) in jobs:
        buf.free()
    ac.lib.acgpu_stream_destroy(s1); ac.lib.acgpu_stream_destroy(s2)


def test_frames_host_multi_splits_a_run_over_the_visible_devices(ac, chk):
    """acgpu_imgconvert_frames_host_multi: contiguous blocks of a host-frame run go to one device each (own thread,
    streams and staging), results are byte-identical to the single-device call.  With one visible GPU it degenerates to
    one block; asking for more devices than exist is refused."""
    ndev = ac.lib.acgpu_device_count()
    sf, df, w, h, nf = F.IMG_YUV420P, F.IMG_BGRA32, 640, 360, 23
    sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
    hs, hd = ac.pinned(nf * sfb), ac.pinned(nf * dfb)
    frames = [ck.random_frame(sf, w, h, seed=900 + i) for i in range(nf)]
    for i in range(nf):
        hs.array[i * sfb:(i + 1) * sfb] = frames[i]
    for use in sorted({1, ndev, 0}):
        hd.array[:] = 0x5A
        ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, sf, hd.ptr, df, w, h, nf, use))
        for i in range(nf):
            want = chk.convert(frames[i], sf, df, w, h, prefill=0x5A, pad=0)[1]
            assert np.array_equal(hd.array[i * dfb:(i + 1) * dfb], want), (use, i)
    assert ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, sf, hd.ptr, df, w, h, nf, ndev + 1) == 0
    assert b"visible" in ac.lib.acgpu_last_error()
    hs.free(); hd.free()


def test_decolor_on_host_frames(ac, chk):
    """-K on host RGB24 frames (src/video_trans.c:381-388 works on the host vframe buffer): staged once per call."""
    for (w, h, nf) in [(320, 48, 3), (37, 5, 2)]:       # the second size takes the two-step fallback (temporary gray plane)
        _decolor_host_case(ac, chk, w, h, nf)


def _decolor_host_case(ac, chk, w, h, nf):
    fb = w * h * 3
    pitch = fb + 7
    frames = [ck.random_frame(F.IMG_RGB24, w, h, seed=70 + i) for i in range(nf)]
    host = np.full(nf * pitch, 0x42, np.uint8)
    for i in range(nf):
        host[i * pitch: i * pitch + fb] = frames[i]
    ac._ok(ac.lib.acgpu_decolor_rgb24_batch(host.ctypes.data, w, h, pitch, nf, None))
    for i in range(nf):
        gray = chk.convert(frames[i], F.IMG_RGB24, F.IMG_GRAY8, w, h, pad=0)[1]
        want = chk.convert(gray, F.IMG_GRAY8, F.IMG_RGB24, w, h, pad=0)[1]
        assert np.array_equal(host[i * pitch: i * pitch + fb], want), i
        assert (host[i * pitch + fb: (i + 1) * pitch] == 0x42).all()


def test_convert_batch_on_host_frames(ac, chk):
    """acgpu_convert_batch with host frames (tcv_convert's callers hold host buffers): staged once per call; bytes the C
    path leaves alone (alpha of YUV -> 32-bit RGB) survive because the host destination is uploaded first; src == dest
    converts in place like tcvideo.c:1044-1064."""
    w, h, nf = 96, 20, 3
    for sf, df in [(F.IMG_YUV420P, F.IMG_ARGB32), (F.IMG_RGB24, F.IMG_YUV422P), (F.IMG_YUY2, F.IMG_YUV420P)]:
        sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
        frames = np.stack([ck.random_frame(sf, w, h, seed=55 + i) for i in range(nf)])
        hs = frames.reshape(-1).copy()
        hd = np.full(nf * dfb, 0x9D, np.uint8)
        ac._ok(ac.lib.acgpu_convert_batch(hs.ctypes.data, hd.ctypes.data, w, h, sf, df, sfb, dfb, nf, None))
        for i in range(nf):
            assert np.array_equal(hd[i * dfb:(i + 1) * dfb], chk.convert(frames[i], sf, df, w, h, prefill=0x9D, pad=0)[1]), (sf, df, i)
        assert np.array_equal(hs, frames.reshape(-1))
    sf, df = F.IMG_RGB24, F.IMG_YUV420P                       # in place on the host
    sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
    frames = np.stack([ck.random_frame(sf, w, h, seed=65 + i) for i in range(nf)])
    buf = frames.reshape(-1).copy()
    ac._ok(ac.lib.acgpu_convert_batch(buf.ctypes.data, buf.ctypes.data, w, h, sf, df, sfb, sfb, nf, None))
    for i in range(nf):
        assert np.array_equal(buf[i * sfb: i * sfb + dfb], chk.convert(frames[i], sf, df, w, h, pad=0)[1])
        assert np.array_equal(buf[i * sfb + dfb:(i + 1) * sfb], frames[i, dfb:])
