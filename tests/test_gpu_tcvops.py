"""-m gpu parity for the element-wise libtcvideo plane operations on the device:
acgpu_clip_batch / reduce / flip_v / flip_h / gamma_correct / antialias (+ deinterlace and resize replayed from the
same case list) against the reference libtcvideo itself (oracle/_ref/libtcv_ref.so when it travelled with the repo,
else the restatement) and against the committed digests of the reference's outputs."""
import hashlib
import json
import os

import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck
import tcv_cases

pkg = entry.load_package()
pytestmark = pytest.mark.gpu

OPS = {"gamma": "gamma_correct"}


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def tcv():
    return ck.best_tcv_checker()


def out_bytes(op, w, h, bpp, args):
    if op == "deinterlace":
        return w * (h // 2 if args[0] >= 2 else h) * bpp
    if op == "resize":
        rw, rh, sw, sh = args
        return (w + rw * sw) * (h + rh * sh) * bpp
    if op == "clip":
        return max(w - args[0] - args[1], 0) * max(h - args[2] - args[3], 0) * bpp
    if op == "reduce":
        rw, rh = args
        if rw <= 0 or rh <= 0:
            return 0
        return w * (h // rh) * bpp if rw == 1 else (w // rw) * (h // rh) * bpp
    return w * h * bpp


def gpu_case(ac, case, nframes=1, gap=0, src_gap=0):
    key, op, (w, h, bpp), args, kind, seed = case
    frames = np.stack([tcv_cases.image(kind, w, h, bpp, seed + 1000 * i) for i in range(nframes)])
    inplace = op in ("flip_v", "flip_h") and args[0]
    call_args = () if op in ("flip_v", "flip_h") else args
    ok, got = ac.plane_op_batch(OPS.get(op, op), frames, out_bytes(op, w, h, bpp, args), w, h, bpp, *call_args,
                                inplace=inplace, dst_gap=gap, src_gap=src_gap)
    return ok, got, frames


def test_case_list_matches_reference_digests_through_libacgpu(ac):
    """Every case the reference libtcvideo was recorded on (tests/golden/tcv_digests.json), through the device."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "tcv_digests.json")) as f:
        gold = json.load(f)["digests"]
    n = 0
    for case in tcv_cases.cases():
        want_ok, want = gold[case[0]]
        if case[1] == "resize" and case[3][0] and case[3][1]:
            continue
        ok, got, _ = gpu_case(ac, case)
        assert ok == want_ok, (case[0], ac.last_error())
        if ok:
            n_out = out_bytes(case[1], *case[2], case[3])
            assert hashlib.sha256(got[0, :n_out].tobytes()).hexdigest()[:16] == want, case[0]
            n += 1
    assert n > 350


def test_case_list_batched_with_gaps_against_the_checker(ac, tcv):
    """Three frames per launch with a 48-byte gap between destination planes and a 16- or 7-byte gap between source
    planes (the latter knocks the batch off the 16-byte-aligned paths): frames are independent, the gaps stay untouched."""
    for case in tcv_cases.cases():
        key, op, (w, h, bpp), args, kind, seed = case
        if op in ("deinterlace", "resize") or (op in ("flip_v", "flip_h") and args[0]):
            continue
        ok, got, frames = gpu_case(ac, case, nframes=3, gap=48, src_gap=(16 if (seed & 1) else 7))
        n_out = out_bytes(op, w, h, bpp, args)
        for i in range(3):
            c2 = (key, op, (w, h, bpp), args, kind, seed + 1000 * i)
            ok_r, want = tcv_cases.run_case(tcv, c2)
            assert ok == ok_r, key
            if ok:
                assert np.array_equal(got[i, :n_out], want[:n_out]), (key, i)
                assert (got[i, n_out:] == 0x55).all(), (key, "gap written")


@pytest.mark.parametrize("bpp", [1, 3])
def test_full_size_planes(ac, tcv, bpp):
    """1920x1080 planes (vector paths) and 1918x1079 (byte paths) for every operation."""
    for (w, h) in [(1920, 1080), (1918, 1079)]:
        src = ck.splitmix_bytes(w * h * bpp, 5)
        blocky = tcv_cases.blocky_image(w, h, bpp, 6)
        f = src[None, :]
        checks = [
            ("clip", (8, 24, 4, 12, 16), tcv.clip(src, w, h, bpp, 8, 24, 4, 12, black=16)),
            ("clip", (-16, -32, -2, -6, 128), tcv.clip(src, w, h, bpp, -16, -32, -2, -6, black=128)),
            ("clip", (3, -5, -7, 9, 0), tcv.clip(src, w, h, bpp, 3, -5, -7, 9, black=0)),
            ("reduce", (2, 2), tcv.reduce(src, w, h, bpp, 2, 2)),
            ("reduce", (1, 2), tcv.reduce(src, w, h, bpp, 1, 2)),
            ("reduce", (1, 1), tcv.reduce(src, w, h, bpp, 1, 1)),
            ("reduce", (4, 3), tcv.reduce(src, w, h, bpp, 4, 3)),
            ("reduce", (3, 2), tcv.reduce(src, w, h, bpp, 3, 2)),
            ("reduce", (3, 3), tcv.reduce(src, w, h, bpp, 3, 3)),
            ("reduce", (4, 4), tcv.reduce(src, w, h, bpp, 4, 4)),
            ("reduce", (5, 2), tcv.reduce(src, w, h, bpp, 5, 2)),
            ("reduce", (5, 5), tcv.reduce(src, w, h, bpp, 5, 5)),
            ("reduce", (6, 1), tcv.reduce(src, w, h, bpp, 6, 1)),
            ("reduce", (7, 3), tcv.reduce(src, w, h, bpp, 7, 3)),      # 1920 / 7 = 274: not on the 16-pixel grid -> gather kernel
            ("reduce", (8, 4), tcv.reduce(src, w, h, bpp, 8, 4)),
            ("reduce", (9, 2), tcv.reduce(src, w, h, bpp, 9, 2)),
            ("clip", (3, 5, 1, 1, 16), tcv.clip(src, w, h, bpp, 3, 5, 1, 1, black=16)),
            ("clip", (1, 0, 0, 0, 16), tcv.clip(src, w, h, bpp, 1, 0, 0, 0, black=16)),
            ("clip", (13, 2, 0, 3, 16), tcv.clip(src, w, h, bpp, 13, 2, 0, 3, black=16)),
            ("flip_v", (), tcv.flip_v(src, w, h, bpp)),
            ("flip_h", (), tcv.flip_h(src, w, h, bpp)),
            ("gamma_correct", (2.2,), tcv.gamma(src, w, h, bpp, 2.2)),
            ("gamma_correct", (0.45,), tcv.gamma(src, w, h, bpp, 0.45)),
        ]
        for op, args, (ok_r, want) in checks:
            ok, got = ac.plane_op_batch(op, f, want.size, w, h, bpp, *args)
            assert ok == ok_r == 1, (op, args, ac.last_error())
            assert np.array_equal(got[0], want), (op, args, w, h, bpp)
        for img in (src, blocky):
            ok_r, want = tcv.antialias(img, w, h, bpp, 0.333, 0.5)
            ok, got = ac.plane_op_batch("antialias", img[None, :], want.size, w, h, bpp, 0.333, 0.5)
            assert ok == ok_r == 1 and np.array_equal(got[0], want)
        # in-place flips (the reference allows src == dest: tcvideo.c:757-762, 806-814)
        for op in ("flip_v", "flip_h"):
            ok_r, want = getattr(tcv, op)(src, w, h, bpp, inplace=True)
            ok, got = ac.plane_op_batch(op, f, want.size, w, h, bpp, inplace=True)
            assert ok == 1 and np.array_equal(got[0], want), op


@pytest.mark.parametrize("bpp", [1, 3])
def test_antialias_batches_cross_frames_inside_one_persistent_launch(ac, tcv, bpp):
    """Five planes per launch (random, blocky, random, ...): the kernel's warps take 32-unit chunks of one frame in turn,
    so chunk ranges end inside frames and the last chunk of a frame is partial (1080p: 129600 units = 4050 chunks;
    720x576 at 4 pixels per unit is not a multiple of 32)."""
    for (w, h) in [(1920, 1080), (724, 578), (64, 3)]:
        imgs = [ck.splitmix_bytes(w * h * bpp, 11), tcv_cases.blocky_image(w, h, bpp, 12)]
        frames = np.stack([imgs[i % 2] for i in range(5)])
        want = [tcv.antialias(im, w, h, bpp, 0.333, 0.5)[1] for im in imgs]
        ok, got = ac.plane_op_batch("antialias", frames, w * h * bpp, w, h, bpp, 0.333, 0.5, dst_gap=64, src_gap=16)
        assert ok == 1, ac.last_error()
        for i in range(5):
            assert np.array_equal(got[i, : w * h * bpp], want[i % 2]), (w, h, bpp, i)
        assert (got[:, w * h * bpp:] == 0x55).all()


def test_in_place_rules(ac, tcv):
    """src == dest (libtcvideo/tcvideo.c:180 allows it): accepted where the reference's sequential in-place result is the
    out-of-place one -- and then equal to what the reference itself leaves in the buffer -- rejected elsewhere."""
    w, h = 192, 64
    for bpp in (1, 3):
        src = ck.splitmix_bytes(w * h * bpp, 21)
        f = src[None, :]

        def both(op, ref_op, out_bytes, *args, ref_args=None):
            ok, got = ac.plane_op_batch(op, f, out_bytes, w, h, bpp, *args, inplace=True)
            assert ok == 1, (op, args, ac.last_error())
            okr, want = ref_op(*(ref_args if ref_args is not None else args))
            assert okr == 1 and np.array_equal(got[0, :out_bytes], want[:out_bytes]), (op, args, bpp)

        ip = lambda name, *a: tcv._plane_op(name, src, 0, w, h, bpp, *a, inplace=True)
        both("clip", lambda *a: ip("clip", *a), (w - 24) * (h - 6) * bpp, 8, 16, 2, 4, 7)
        both("reduce", lambda *a: ip("reduce", *a), (w // 2) * (h // 2) * bpp, 2, 2)
        both("reduce", lambda *a: ip("reduce", *a), (w // 3) * (h // 4) * bpp, 3, 4)
        both("reduce", lambda *a: ip("reduce", *a), w * (h // 2) * bpp, 1, 2)
        both("deinterlace", lambda *a: ip("deinterlace", *a), w * h * bpp, 0)            # interpolate
        both("deinterlace", lambda *a: ip("deinterlace", *a), w * (h // 2) * bpp, 2)     # drop top field
        both("deinterlace", lambda *a: ip("deinterlace", *a), w * (h // 2) * bpp, 3)     # drop bottom field
        both("resize", lambda *a: ip("resize", *a), w * (h - 16) * bpp, 0, -2, 8, 8)
        both("resize", lambda *a: ip("resize", *a), (w - 24) * h * bpp, -3, 0, 8, 8)
        both("gamma_correct", lambda *a: ip("gamma", *a), w * h * bpp, 0.7)
        # rejected: the reference's own in-place result depends on bytes it has already overwritten
        for op, args in [("clip", (-8, 0, 0, 0, 0)), ("deinterlace", (1,)), ("resize", (0, 2, 8, 8)), ("resize", (3, 0, 8, 8)),
                         ("antialias", (0.3, 0.5))]:
            ok, _ = ac.plane_op_batch(op, f, w * h * bpp, w, h, bpp, *args, inplace=True)
            assert ok == 0 and ac.last_error(), (op, args)
    # partially overlapping device buffers are rejected outright
    buf = ac.malloc(4 * w * h)
    for fn, args in [(ac.lib.acgpu_flip_v_batch, ()), (ac.lib.acgpu_gamma_correct_batch, (1.5,)), (ac.lib.acgpu_reduce_batch, (2, 2))]:
        assert fn(buf.ptr, buf.ptr + 64, w, h, 1, *args, 0, 0, 1, None) == 0 and "overlap" in ac.last_error()
    buf.free()


def test_staged_calls_order_after_the_callers_stream(ac, tcv):
    """A batched call that hands a device-resident source to a HOST destination runs on the thread's private stream; it
    must wait for what the caller's own stream still has in flight (here: the conversion that produces the source)."""
    F = ck.F
    w, h, nf = 1920, 1080, 24
    chk = ck.best_checker()
    yuv = ck.random_frame(F.IMG_YUV420P, w, h, seed=31)
    _, rgb = chk.convert(yuv, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)
    _, want = tcv.flip_v(rgb, w, h, 3)
    sfb, dfb = yuv.size, rgb.size
    ds = ac.malloc(nf * sfb)
    for i in range(nf):
        ds.upload(yuv, offset=i * sfb)
    dd = ac.malloc(nf * dfb).fill(0)
    host = np.zeros(nf * dfb, np.uint8)
    stream = ac.lib.acgpu_stream_create()
    for _ in range(3):
        dd.fill(0)
        ac._ok(ac.imgconvert_batch(ds.ptr, F.IMG_YUV420P, sfb, dd.ptr, F.IMG_RGB24, dfb, w, h, nf, stream))     # asynchronous on `stream`
        ac._ok(ac.lib.acgpu_flip_v_batch(dd.ptr, host.ctypes.data, w, h, 3, dfb, dfb, nf, stream))                # device -> host, staged
        got = host.reshape(nf, dfb)
        for i in (0, nf // 2, nf - 1):
            assert np.array_equal(got[i], want), i
    ac.lib.acgpu_stream_destroy(stream)
    ds.free(); dd.free()


def test_flips_are_involutions_and_commute(ac):
    w, h, bpp = 1280, 720, 3
    f = ck.splitmix_bytes(w * h * bpp, 9)[None, :]
    _, v = ac.plane_op_batch("flip_v", f, f.size, w, h, bpp)
    _, vv = ac.plane_op_batch("flip_v", v, f.size, w, h, bpp)
    assert np.array_equal(vv, f)
    _, hh = ac.plane_op_batch("flip_h", f, f.size, w, h, bpp)
    _, hv = ac.plane_op_batch("flip_v", hh, f.size, w, h, bpp)
    _, vh = ac.plane_op_batch("flip_h", v, f.size, w, h, bpp)
    assert np.array_equal(hv, vh)
    img = f.reshape(h, w, bpp)
    assert np.array_equal(hv.reshape(h, w, bpp), img[::-1, ::-1])


def test_rejections_match_the_reference(ac):
    buf = ac.malloc(1 << 16)
    L = ac.lib
    assert L.acgpu_clip_batch(buf.ptr, buf.ptr, 64, 32, 2, 0, 0, 0, 0, 0, 0, 0, 1, None) == 0          # Bpp
    assert L.acgpu_clip_batch(buf.ptr, buf.ptr, 64, 32, 1, 32, 32, 0, 0, 0, 0, 0, 1, None) == 0        # nothing left
    assert L.acgpu_clip_batch(buf.ptr, buf.ptr, 64, 32, 1, 0, 0, 40, -8, 0, 0, 0, 1, None) == 0
    assert L.acgpu_reduce_batch(buf.ptr, buf.ptr, 64, 32, 1, 0, 1, 0, 0, 1, None) == 0
    assert L.acgpu_reduce_batch(None, buf.ptr, 64, 32, 1, 1, 1, 0, 0, 1, None) == 0
    assert L.acgpu_flip_v_batch(buf.ptr, buf.ptr, 0, 32, 1, 0, 0, 1, None) == 0
    assert L.acgpu_flip_h_batch(buf.ptr, buf.ptr, 64, -1, 3, 0, 0, 1, None) == 0
    assert L.acgpu_gamma_correct_batch(buf.ptr, buf.ptr, 64, 32, 1, 0.0, 0, 0, 1, None) == 0
    assert L.acgpu_gamma_correct_batch(buf.ptr, buf.ptr, 64, 32, 1, -2.0, 0, 0, 1, None) == 0
    assert L.acgpu_antialias_batch(buf.ptr, buf.ptr + 4096, 64, 32, 1, 1.5, 0.5, 0, 0, 1, None) == 0
    assert L.acgpu_antialias_batch(buf.ptr, buf.ptr + 4096, 64, 32, 1, 0.5, -0.5, 0, 0, 1, None) == 0
    assert L.acgpu_antialias_batch(buf.ptr, buf.ptr, 64, 32, 1, 0.5, 0.5, 0, 0, 1, None) == 0          # overlap
    assert b"overlap" in L.acgpu_last_error()
    buf.free()


def test_batches_longer_than_one_grid(ac, tcv):
    """grid.y carries the frame index, so batches above 65535 frames are cut into several launches."""
    w, h, bpp, nf = 16, 2, 1, 70000
    frames = ck.splitmix_bytes(nf * w * h * bpp, 77).reshape(nf, -1)
    _, got = ac.plane_op_batch("flip_v", frames, w * h * bpp, w, h, bpp)
    assert np.array_equal(got.reshape(nf, h, w), frames.reshape(nf, h, w)[:, ::-1])
    _, got = ac.plane_op_batch("gamma_correct", frames, w * h * bpp, w, h, bpp, 2.2)
    table = ck.Oracle().gamma_table(2.2)
    assert np.array_equal(got, table[frames])
    _, got = ac.plane_op_batch("clip", frames, (w - 4) * h * bpp, w, h, bpp, 2, 2, 0, 0, 0)
    assert np.array_equal(got.reshape(nf, h, w - 4), frames.reshape(nf, h, w)[:, :, 2:-2])
    for i in (0, 32767, 32768, 65535, 65536, nf - 1):
        assert np.array_equal(ac.plane_op_batch("flip_h", frames[i:i + 1], w * h, w, h, bpp)[1][0], tcv.flip_h(frames[i], w, h, bpp)[1])


def test_concurrent_threads_each_with_their_own_tables(ac):
    """transcode runs N frame threads (src/frame_threads.c:174-228); libacgpu keeps streams, staging and the table
    cache per caller thread.  Six threads mix every frame-granular operation on their own planes; each result is
    checked against that thread's own reference handle."""
    import threading

    errors = []

    def worker(tid):
        try:
            a = pkg.AcGpu()
            ref = ck.best_tcv_checker()
            rng = np.random.default_rng(700 + tid)
            w, h = 64 + 16 * tid, 24 + 2 * tid
            for it in range(40):
                bpp = 1 if (it + tid) % 2 else 3
                src = tcv_cases.blocky_image(w, h, bpp, 800 + tid * 100 + it)
                f = src[None, :]
                op = int(rng.integers(0, 7))
                if op == 0:
                    g = 0.4 + 0.05 * int(rng.integers(0, 40))
                    got, want = a.plane_op_batch("gamma_correct", f, src.size, w, h, bpp, g)[1][0], ref.gamma(src, w, h, bpp, g)[1]
                elif op == 1:
                    rw = -int(rng.integers(1, 4))
                    nw = w + rw * 8
                    got, want = a.plane_op_batch("resize", f, nw * h * bpp, w, h, bpp, rw, 0, 8, 2)[1][0], ref.resize(src, w, h, bpp, rw, 0, 8, 2)
                elif op == 2:
                    rh = int(rng.integers(1, 3))
                    nh = h + rh * 2
                    got, want = a.plane_op_batch("resize", f, w * nh * bpp, w, h, bpp, 0, rh, 8, 2)[1][0], ref.resize(src, w, h, bpp, 0, rh, 8, 2)
                elif op == 3:
                    args = (int(rng.integers(-3, 6)), int(rng.integers(-3, 6)), int(rng.integers(-2, 4)), int(rng.integers(-2, 4)))
                    ok_r, want = ref.clip(src, w, h, bpp, *args, black=16)
                    got = a.plane_op_batch("clip", f, want.size, w, h, bpp, *args, 16)[1][0]
                elif op == 4:
                    got, want = a.plane_op_batch("antialias", f, src.size, w, h, bpp, 0.333, 0.5)[1][0], ref.antialias(src, w, h, bpp, 0.333, 0.5)[1]
                elif op == 5:
                    mode = int(rng.integers(0, 4))
                    want = ref.deinterlace(src, w, h, bpp, mode)
                    got = a.plane_op_batch("deinterlace", f, want.size, w, h, bpp, mode)[1][0]
                else:
                    got, want = a.plane_op_batch("flip_h", f, src.size, w, h, bpp)[1][0], ref.flip_h(src, w, h, bpp)[1]
                if not np.array_equal(got[: want.size], want):
                    errors.append((tid, it, op))
                    return
        except Exception as e:      # noqa: BLE001 -- report through the main thread
            errors.append((tid, repr(e), a.last_error()))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


@pytest.mark.parametrize("kind", ["pageable", "pinned", "mixed"])
def test_host_frames_are_staged_once_per_call(ac, tcv, kind):
    """The frame-granular entry points also take HOST planes (what an unmodified libtcvideo caller holds): the batch is
    uploaded once, processed on the device, downloaded once, and the call returns with the result in place.  Pageable,
    page-locked and mixed (host source, device destination and the reverse) combinations; two planes per call with
    gaps between them that must stay untouched."""
    L = ac.lib
    nf = 2
    for case in tcv_cases.cases():
        key, op, (w, h, bpp), args, img_kind, seed = case
        if (w, h) not in ((64, 16), (33, 7), (64, 48), (32, 16)) or (op in ("flip_v", "flip_h") and args[0] and kind == "mixed"):
            continue
        inplace = op in ("flip_v", "flip_h") and args[0]
        call_args = () if op in ("flip_v", "flip_h") else args
        frames = [tcv_cases.image(img_kind, w, h, bpp, seed + 1000 * i) for i in range(nf)]
        refs = [tcv_cases.run_case(tcv, (key, op, (w, h, bpp), args, img_kind, seed + 1000 * i)) for i in range(nf)]
        ok_ref = refs[0][0]
        nout = out_bytes(op, w, h, bpp, args)
        sfb = w * h * bpp
        spitch, dpitch = sfb + 5, (sfb + 5 if inplace else nout + 9)
        hs = np.full(nf * spitch, 0x21, np.uint8)
        for i in range(nf):
            hs[i * spitch: i * spitch + sfb] = frames[i]
        hd = hs if inplace else np.full(max(nf * dpitch, 1), 0x55, np.uint8)
        keep = []
        if kind == "pinned":
            ps, pd = ac.pinned(hs.size), ac.pinned(hd.size)
            ps.array[:] = hs
            pd.array[:] = hd
            sptr, dptr = ps.ptr, (ps.ptr if inplace else pd.ptr)
            read = (lambda: ps.array.copy()) if inplace else (lambda: pd.array.copy())
            keep = [ps, pd]
        elif kind == "mixed" and (seed & 1):
            dsrc = ac.malloc(hs.size).upload(hs)                  # device source, pageable destination
            sptr, dptr, read = dsrc.ptr, hd.ctypes.data, (lambda: hd.copy())
            keep = [dsrc]
        elif kind == "mixed":
            ddst = ac.malloc(hd.size).upload(hd)                  # pageable source, device destination
            sptr, dptr, read = hs.ctypes.data, ddst.ptr, ddst.download
            keep = [ddst]
        else:
            sptr, dptr = hs.ctypes.data, (hs.ctypes.data if inplace else hd.ctypes.data)
            read = (lambda: hs.copy()) if inplace else (lambda: hd.copy())
        fn = getattr(L, f"acgpu_{OPS.get(op, op)}_batch")
        ok = fn(sptr, dptr, w, h, bpp, *call_args, spitch, dpitch, nf, None)
        assert ok == ok_ref, (key, kind, ac.last_error())
        if ok:
            got = read()
            ac.sync()
            for i in range(nf):
                assert np.array_equal(got[i * dpitch: i * dpitch + nout], refs[i][1][:nout]), (key, kind, i)
                gap = got[i * dpitch + nout: (i + 1) * dpitch]
                assert (gap == (0x21 if inplace else 0x55)).all(), (key, kind, "gap written")
        for k in keep:
            k.free()


@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160), (854, 480)])
def test_size_independent_properties_at_full_sizes(ac, size):
    """Relations that hold at any size, checked at BASELINE's frame sizes where the checker would take long:
    grow-then-clip is the identity and the grown border is the black value; reduce by (1,1) is a copy; reduce by
    (2,2) is the even pixels of the even rows; dropping the bottom field gives the even rows, which are also the even
    rows of the interpolating deinterlacer; vertical flip twice is the identity."""
    w, h = size
    for bpp in (1, 3):
        src = ck.splitmix_bytes(w * h * bpp, 17 + bpp)
        f = src[None, :]
        img = src.reshape(h, w, bpp)
        gw, gh = w + 24, h + 10
        _, grown = ac.plane_op_batch("clip", f, gw * gh * bpp, w, h, bpp, -8, -16, -4, -6, 99)
        g = grown[0].reshape(gh, gw, bpp)
        assert np.array_equal(g[4:4 + h, 8:8 + w], img)
        assert (g[:4] == 99).all() and (g[4 + h:] == 99).all() and (g[:, :8] == 99).all() and (g[:, 8 + w:] == 99).all()
        _, back = ac.plane_op_batch("clip", grown, w * h * bpp, gw, gh, bpp, 8, 16, 4, 6, 0)
        assert np.array_equal(back[0], src)
        _, same = ac.plane_op_batch("reduce", f, w * h * bpp, w, h, bpp, 1, 1)
        assert np.array_equal(same[0], src)
        _, half = ac.plane_op_batch("reduce", f, (w // 2) * (h // 2) * bpp, w, h, bpp, 2, 2)
        assert np.array_equal(half[0].reshape(h // 2, w // 2, bpp), img[0:h - h % 2:2, 0:w - w % 2:2])
        _, even = ac.plane_op_batch("deinterlace", f, w * (h // 2) * bpp, w, h, bpp, 3)
        assert np.array_equal(even[0].reshape(h // 2, w, bpp), img[0:2 * (h // 2):2])
        _, interp = ac.plane_op_batch("deinterlace", f, w * h * bpp, w, h, bpp, 0)
        assert np.array_equal(interp[0].reshape(h, w, bpp)[0::2], img[0::2])
        _, v = ac.plane_op_batch("flip_v", f, w * h * bpp, w, h, bpp)
        _, vv = ac.plane_op_batch("flip_v", v, w * h * bpp, w, h, bpp)
        assert np.array_equal(vv[0], src) and np.array_equal(v[0].reshape(h, w, bpp), img[::-1])
