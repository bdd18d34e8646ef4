"""-m gpu parity for the frame chains (include/acgpu.h acgpu_chain_*): several operations per PCIe round trip.

The reference applies its operations one call after the other on host frames (do_process_frame, src/video_trans.c:192-426;
tcv_convert pairs in the filter wrappers, filter/filter_ascii.c:367-373).  tests/chain_ref.py replays such a list with the
reference libraries themselves; the chain entry points must give the same bytes through all three doors: device-resident
batch, host-frame pipeline, multi-device host pipeline."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck
from chain_ref import (ANTIALIAS, CLIP, CONVERT, DECOLOR, DEINTERLACE, FLIP_H, FLIP_V, GAMMA, REDUCE, RESIZE, RGBSWAP, ref_chain)
from checkers import F

pkg = entry.load_package()
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def conv():
    return ck.best_checker()


@pytest.fixture(scope="module")
def tcv():
    return ck.best_tcv_checker()


def chain_output(ac, fmt, w, h, ops, nops):
    of, ow, oh = C.c_int(0), C.c_int(0), C.c_int(0)
    ok = ac.lib.acgpu_chain_output(fmt, w, h, ops, nops, C.byref(of), C.byref(ow), C.byref(oh))
    return ok, of.value, ow.value, oh.value


def run_device(ac, frames, fmt, w, h, stages, gap=0):
    ops = pkg.chain_ops(stages)
    ok, of, ow, oh = chain_output(ac, fmt, w, h, ops, len(stages))
    assert ok == 1, ac.last_error()
    nf, inb, outb = frames.shape[0], F.frame_bytes(fmt, w, h), F.frame_bytes(of, ow, oh)
    sp, dp = inb + gap, outb + gap
    hs = np.full((nf, sp), 0xEE, np.uint8)
    hs[:, :inb] = frames
    ds = ac.malloc(nf * sp).upload(hs.reshape(-1))
    dd = ac.malloc(nf * dp).fill(0x55)
    ac._ok(ac.lib.acgpu_chain_batch(ds.ptr, fmt, w, h, sp, dd.ptr, dp, ops, len(stages), nf, None))
    ac.sync()
    out = dd.download().reshape(nf, dp)
    src_after = ds.download().reshape(nf, sp)
    assert np.array_equal(src_after, hs), "the chain modified its source"
    assert (out[:, outb:] == 0x55).all(), "the chain wrote past its destination frames"
    ds.free(); dd.free()
    return out[:, :outb], (of, ow, oh)


def run_host(ac, frames, fmt, w, h, stages, multi=False, pinned=True, prefill=0x55):
    ops = pkg.chain_ops(stages)
    ok, of, ow, oh = chain_output(ac, fmt, w, h, ops, len(stages))
    assert ok == 1, ac.last_error()
    nf, inb, outb = frames.shape[0], F.frame_bytes(fmt, w, h), F.frame_bytes(of, ow, oh)
    if pinned:
        hs, hd = ac.pinned(nf * inb), ac.pinned(nf * outb + 64)
        hs.array[:] = frames.reshape(-1)
        hd.array[:] = prefill
        sptr, dptr, darr = hs.ptr, hd.ptr, hd.array
    else:
        s = np.ascontiguousarray(frames.reshape(-1))
        darr = np.full(nf * outb + 64, prefill, np.uint8)
        sptr, dptr = s.ctypes.data, darr.ctypes.data
    if multi:
        rc = ac.lib.acgpu_chain_frames_host_multi(sptr, fmt, w, h, dptr, ops, len(stages), nf, 0)
    else:
        rc = ac.lib.acgpu_chain_frames_host(sptr, fmt, w, h, dptr, ops, len(stages), nf)
    assert rc == 1, ac.last_error()
    out = np.array(darr[:nf * outb]).reshape(nf, outb)
    assert (np.asarray(darr[nf * outb:]) == prefill).all()
    if pinned:
        hs.free(); hd.free()
    return out, (of, ow, oh)


def expect(tcv, conv, frames, fmt, w, h, stages):
    outs, geo = [], None
    for f in frames:
        o, of, ow, oh = ref_chain(tcv, conv, f, fmt, w, h, stages)
        outs.append(o)
        geo = (of, ow, oh)
    return np.stack(outs), geo


def frames_of(fmt, w, h, nf, seed):
    return np.stack([ck.random_frame(fmt, w, h, seed=seed + i) for i in range(nf)])


# ---- BASELINE config 4: 3840x2160 YUV420P -> RGB24 -> YUV422P as ONE resident chain --------------------------------
def test_config4_uhd_round_trip_is_the_two_reference_conversions(ac, tcv, conv):
    w, h, nf = 3840, 2160, 3
    stages = [(CONVERT, F.IMG_RGB24), (CONVERT, F.IMG_YUV422P)]
    frames = frames_of(F.IMG_YUV420P, w, h, nf, 400)
    want, geo = expect(tcv, conv, frames, F.IMG_YUV420P, w, h, stages)
    assert geo == (F.IMG_YUV422P, w, h)
    got, g2 = run_device(ac, frames, F.IMG_YUV420P, w, h, stages)
    assert g2 == geo and np.array_equal(got, want)
    got, _ = run_host(ac, frames, F.IMG_YUV420P, w, h, stages)
    assert np.array_equal(got, want)
    got, _ = run_host(ac, frames, F.IMG_YUV420P, w, h, stages, multi=True)
    assert np.array_equal(got, want)


# ---- do_process_frame-shaped chains against the reference libtcvideo -----------------------------------------------
PROCESS_CHAINS = [
    # -j 16,8,16,8 -I 1 -B 4,6 -z -G 0.8 on a PAL YUV420P frame
    ("yuv420p_pal", F.IMG_YUV420P, 720, 576,
     [(CLIP, 8, 8, 16, 16), (DEINTERLACE, 1), (RESIZE, -6, -4), (FLIP_V,), (GAMMA, 0.8)]),
    # RGB24: linear blend, shrink rows, mirror, -k, -K, antialias
    ("rgb24_720p", F.IMG_RGB24, 1280, 720,
     [(DEINTERLACE, 5), (RESIZE, 0, -30), (FLIP_H,), (RGBSWAP,), (DECOLOR,), (ANTIALIAS, 0.333, 0.5)]),
    # YUV422P: drop field, enlarge both ways, grow with black borders, reduce, -k, -K
    ("yuv422p_small", F.IMG_YUV422P, 352, 288,
     [(DEINTERLACE, 4), (RESIZE, 4, 2), (CLIP, -16, -16, -8, -8), (REDUCE, 2, 2), (RGBSWAP,), (DECOLOR,)]),
    # the filter-wrapper shape: yuv -> rgb -> (rgb-only operation) -> yuv  (filter/filter_ascii.c:367-373)
    ("wrapper_1080p", F.IMG_YUV420P, 1920, 1080,
     [(CONVERT, F.IMG_RGB24), (DECOLOR,), (GAMMA, 1.6), (CONVERT, F.IMG_YUV420P)]),
    # in-place stages only, and a chain that opens with one (the source must survive)
    ("inplace_only", F.IMG_YUV420P, 640, 480, [(GAMMA, 2.2), (DECOLOR,), (RGBSWAP,)]),
    ("inplace_first", F.IMG_RGB24, 640, 480, [(RGBSWAP,), (FLIP_V,), (GAMMA, 0.45)]),
    # no-op stages vanish
    ("noops", F.IMG_YUV420P, 320, 240, [(DEINTERLACE, 2), (CONVERT, F.IMG_YUV420P), (RESIZE, 0, 0)]),
    ("empty", F.IMG_YUV422P, 320, 240, []),
    # a single 8-bit plane (set_vtd's default layout): config 3's row shapes as chains
    ("y8_rows", F.IMG_Y8, 1920, 1080, [(DEINTERLACE, 1), (RESIZE, 0, -45), (ANTIALIAS, 0.333, 0.5), (CLIP, 3, 5, 1, 2)]),
    # both-dimension resize in every buffer situation (first, middle, last stage)
    ("resize2_first", F.IMG_YUV420P, 640, 480, [(RESIZE, -8, 6), (FLIP_H,)]),
    ("resize2_mid", F.IMG_RGB24, 640, 480, [(FLIP_V,), (RESIZE, 10, -12), (FLIP_H,), (RESIZE, -3, 2)]),
    ("resize2_last", F.IMG_YUV422P, 640, 480, [(FLIP_V,), (FLIP_H,), (RESIZE, 2, 2)]),
    # long chain through several formats
    ("long", F.IMG_UYVY, 768, 576,
     [(CONVERT, F.IMG_YUV422P), (CLIP, 32, 32, 0, 0), (ANTIALIAS, 0.5, 0.25), (CONVERT, F.IMG_RGB24), (REDUCE, 2, 1),
      (CONVERT, F.IMG_YUV420P), (DEINTERLACE, 5), (CONVERT, F.IMG_YUY2)]),
]


@pytest.mark.parametrize("case", PROCESS_CHAINS, ids=[c[0] for c in PROCESS_CHAINS])
def test_process_frame_chains(ac, tcv, conv, case):
    _, fmt, w, h, stages = case
    nf = 3
    frames = frames_of(fmt, w, h, nf, 500)
    want, geo = expect(tcv, conv, frames, fmt, w, h, stages)
    got, g2 = run_device(ac, frames, fmt, w, h, stages, gap=256)
    assert g2 == geo
    assert np.array_equal(got, want), "device-resident chain"
    got, _ = run_host(ac, frames, fmt, w, h, stages)
    assert np.array_equal(got, want), "host-frame pipeline (pinned)"
    got, _ = run_host(ac, frames[:2], fmt, w, h, stages, pinned=False)
    assert np.array_equal(got, want[:2]), "host-frame pipeline (pageable)"
    got, _ = run_host(ac, frames, fmt, w, h, stages, multi=True)
    assert np.array_equal(got, want), "multi-device host pipeline"


def test_chain_many_frames_crosses_pipeline_slots_and_sub_batches(ac, tcv, conv):
    # 40 frames of 1080p: several chunks per pipeline slot on the host path; the device path is also run with a small
    # temporary so the batch is cut into several sub-batches (the budget is read once per process: tests/knob_worker.py
    # style subprocesses cover other values)
    w, h, nf = 1920, 1080, 40
    stages = [(CONVERT, F.IMG_RGB24), (FLIP_V,), (CONVERT, F.IMG_YUV422P)]
    uniq = frames_of(F.IMG_YUV420P, w, h, 4, 600)
    frames = np.stack([uniq[i % 4] for i in range(nf)])
    want4, _ = expect(tcv, conv, uniq, F.IMG_YUV420P, w, h, stages)
    got, _ = run_device(ac, frames, F.IMG_YUV420P, w, h, stages)
    for i in range(nf):
        assert np.array_equal(got[i], want4[i % 4]), i
    got, _ = run_host(ac, frames, F.IMG_YUV420P, w, h, stages)
    for i in range(nf):
        assert np.array_equal(got[i], want4[i % 4]), i


def test_chain_ending_in_a_32_bit_conversion_keeps_the_callers_alpha(ac, tcv, conv):
    # the C path never stores alpha (img_yuv_rgb.c:62-64): through host frames the caller's bytes must survive
    w, h, nf = 256, 64, 2
    stages = [(FLIP_V,), (CONVERT, F.IMG_BGRA32)]
    frames = frames_of(F.IMG_YUV420P, w, h, nf, 700)
    want, _ = expect(tcv, conv, frames, F.IMG_YUV420P, w, h, stages)     # checker prefill is 0x55 too
    got, _ = run_host(ac, frames, F.IMG_YUV420P, w, h, stages, prefill=0x55)
    assert np.array_equal(got, want)
    got, _ = run_device(ac, frames, F.IMG_YUV420P, w, h, stages)
    assert np.array_equal(got, want)


def test_chain_fuzz(ac, tcv, conv):
    rng = np.random.default_rng(20261018)
    planar = [F.IMG_YUV420P, F.IMG_YUV422P, F.IMG_RGB24]
    others = [F.IMG_YUY2, F.IMG_UYVY, F.IMG_YUV444P, F.IMG_BGR24, F.IMG_RGBA32, F.IMG_GRAY8, F.IMG_Y8, F.IMG_YUV411P]
    done = 0
    for it in range(60):
        fmt = int(rng.choice(planar))
        w, h = int(rng.integers(4, 40)) * 16, int(rng.integers(4, 30)) * 16
        stages, cf, cw, ch = [], fmt, w, h
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.integers(1, 12))
            if cf not in planar and k != CONVERT:
                k = CONVERT
            if k == CONVERT:
                df = int(rng.choice(planar + (others if rng.random() < 0.3 else [])))
                stages.append((CONVERT, df)); cf = df
            elif k == CLIP:
                l, r, t, b = (int(v) * 4 for v in rng.integers(-3, 6, 4))
                if cw - l - r < 32 or ch - t - b < 32:
                    continue
                stages.append((CLIP, l, r, t, b)); cw -= l + r; ch -= t + b
            elif k == DEINTERLACE:
                m = int(rng.choice([1, 2, 4, 5]))
                if m == 4 and (ch % 4 or ch < 16):
                    continue
                stages.append((DEINTERLACE, m)); ch = ch // 2 if m == 4 else ch
            elif k == RESIZE:
                if cw % 8 or ch % 8:
                    continue
                rw, rh = int(rng.integers(-2, 3)), int(rng.integers(-2, 3))
                if cw + rw * 8 < 32 or ch + rh * 8 < 32:
                    continue
                stages.append((RESIZE, rw, rh)); cw += rw * 8; ch += rh * 8
            elif k == REDUCE:
                rw, rh = int(rng.integers(1, 4)), int(rng.integers(1, 4))
                nw, nh = cw // rw, ch // rh
                if nw < 16 or nh < 16 or nw % 2 or nh % 2:
                    continue
                stages.append((REDUCE, rw, rh)); cw, ch = nw, nh
            elif k in (FLIP_V, FLIP_H, RGBSWAP, DECOLOR):
                stages.append((k,))
            elif k == GAMMA:
                stages.append((GAMMA, float(rng.choice([0.45, 0.8, 1.0, 2.2]))))
            else:
                stages.append((ANTIALIAS, float(rng.choice([0.2, 0.333, 0.9])), float(rng.choice([0.0, 0.5, 1.0]))))
        if cf not in planar and stages and stages[-1][0] != CONVERT:
            continue
        frames = frames_of(fmt, w, h, 2, 800 + it)
        want, geo = expect(tcv, conv, frames, fmt, w, h, stages)
        got, g2 = run_device(ac, frames, fmt, w, h, stages)
        assert g2 == geo, (it, stages)
        assert np.array_equal(got, want), (it, fmt, w, h, stages)
        got, _ = run_host(ac, frames, fmt, w, h, stages)
        assert np.array_equal(got, want), (it, fmt, w, h, stages, "host")
        done += 1
    assert done >= 40


# ---- frame lists: every frame in a buffer of its own (transcode's frame ring), in place, or inside a stream buffer -----
def ptr_list(addrs):
    return (C.c_void_p * len(addrs))(*addrs)


def run_list(ac, src_bufs, dst_bufs, fmt, w, h, stages, multi=False):
    ops = pkg.chain_ops(stages)
    sp, dp = ptr_list([b.ctypes.data for b in src_bufs]), ptr_list([b.ctypes.data for b in dst_bufs])
    if multi:
        return ac.lib.acgpu_chain_frame_list_host_multi(sp, fmt, w, h, dp, ops, len(stages), len(src_bufs), 0)
    return ac.lib.acgpu_chain_frame_list_host(sp, fmt, w, h, dp, ops, len(stages), len(src_bufs))


@pytest.mark.parametrize("multi", [False, True], ids=["one_device", "all_devices"])
def test_frame_list_scattered_buffers(ac, tcv, conv, multi):
    # 11 frames, each in its own (pageable) allocation with its own guard band, given in shuffled address order
    w, h, nf = 640, 480, 11
    stages = [(DEINTERLACE, 5), (CONVERT, F.IMG_RGB24), (FLIP_H,), (CONVERT, F.IMG_YUV422P)]
    frames = frames_of(F.IMG_YUV420P, w, h, nf, 900)
    want, geo = expect(tcv, conv, frames, F.IMG_YUV420P, w, h, stages)
    inb, outb = F.frame_bytes(F.IMG_YUV420P, w, h), F.frame_bytes(geo[0], geo[1], geo[2])
    order = np.random.default_rng(5).permutation(nf)
    src_bufs = [None] * nf
    dst_bufs = [None] * nf
    for i in order:          # allocation order != list order
        src_bufs[i] = np.concatenate([frames[i], np.full(32, 0xEE, np.uint8)])
        dst_bufs[i] = np.full(outb + 48, 0x55, np.uint8)
    assert run_list(ac, src_bufs, dst_bufs, F.IMG_YUV420P, w, h, stages, multi) == 1, ac.last_error()
    for i in range(nf):
        assert np.array_equal(dst_bufs[i][:outb], want[i]), i
        assert (dst_bufs[i][outb:] == 0x55).all(), "wrote past frame %d" % i
        assert np.array_equal(src_bufs[i][:inb], frames[i]) and (src_bufs[i][inb:] == 0xEE).all(), "source %d modified" % i


def test_frame_list_in_place_like_do_process_frame(ac, tcv, conv):
    # dest_frames[i] == src_frames[i]: the frame buffer is processed in place (video_trans.c works on video_buf itself);
    # the chain shrinks the frame, bytes behind the new frame keep the old picture
    w, h, nf = 720, 576, 5
    stages = [(CLIP, 8, 8, 16, 16), (DEINTERLACE, 1), (RESIZE, -6, -4), (GAMMA, 0.8)]
    frames = frames_of(F.IMG_YUV420P, w, h, nf, 910)
    want, geo = expect(tcv, conv, frames, F.IMG_YUV420P, w, h, stages)
    outb = F.frame_bytes(geo[0], geo[1], geo[2])
    bufs = [ac.pinned(frames.shape[1]) for _ in range(nf)]
    for i in range(nf):
        bufs[i].array[:] = frames[i]
    ops = pkg.chain_ops(stages)
    pl = ptr_list([b.ptr for b in bufs])
    assert ac.lib.acgpu_chain_frame_list_host(pl, F.IMG_YUV420P, w, h, pl, ops, len(stages), nf) == 1, ac.last_error()
    for i in range(nf):
        got = np.asarray(bufs[i].array)
        assert np.array_equal(got[:outb], want[i]), i
        assert np.array_equal(got[outb:], frames[i][outb:]), "bytes behind the new frame %d changed" % i
        bufs[i].free()


def test_frame_list_into_a_yuv4mpeg_stream_buffer(ac, tcv, conv):
    # encode_yuv4mpeg.c:256-289: tcv_convert to YUV420P, then "FRAME\n" + the planes go to the stream.  Here the frames of a
    # chain land behind their headers directly: dest_frames[i] = stream + header + i * (6 + frame bytes)
    w, h, nf = 352, 288, 9
    stages = [(CONVERT, F.IMG_YUV420P)]
    frames = frames_of(F.IMG_RGB24, w, h, nf, 920)
    want, geo = expect(tcv, conv, frames, F.IMG_RGB24, w, h, stages)
    outb = F.frame_bytes(F.IMG_YUV420P, w, h)
    head = b"YUV4MPEG2 W352 H288 F25:1 Ip A1:1 C420jpeg\n"
    stream = ac.pinned(len(head) + nf * (6 + outb))
    st = stream.array
    st[:] = 0x55
    st[:len(head)] = np.frombuffer(head, np.uint8)
    for i in range(nf):
        o = len(head) + i * (6 + outb)
        st[o:o + 6] = np.frombuffer(b"FRAME\n", np.uint8)
    src = ac.pinned(frames.size)
    src.array[:] = frames.reshape(-1)
    sp = ptr_list([src.ptr + i * frames.shape[1] for i in range(nf)])
    dp = ptr_list([stream.ptr + len(head) + i * (6 + outb) + 6 for i in range(nf)])
    ops = pkg.chain_ops(stages)
    assert ac.lib.acgpu_chain_frame_list_host(sp, F.IMG_RGB24, w, h, dp, ops, 1, nf) == 1, ac.last_error()
    expect_stream = bytearray(head)
    for i in range(nf):
        expect_stream += b"FRAME\n" + want[i].tobytes()
    assert bytes(np.asarray(st)) == bytes(expect_stream)
    src.free(); stream.free()


def test_frame_list_keeps_alpha_and_rejects_null_frames(ac, tcv, conv):
    w, h, nf = 128, 32, 3
    stages = [(CONVERT, F.IMG_ARGB32)]
    frames = frames_of(F.IMG_YUY2, w, h, nf, 930)
    want, _ = expect(tcv, conv, frames, F.IMG_YUY2, w, h, stages)       # the checker's destination prefill is 0x55
    src_bufs = [frames[i].copy() for i in range(nf)]
    dst_bufs = [np.full(w * h * 4, 0x55, np.uint8) for _ in range(nf)]
    assert run_list(ac, src_bufs, dst_bufs, F.IMG_YUY2, w, h, stages) == 1, ac.last_error()
    for i in range(nf):
        assert np.array_equal(dst_bufs[i], want[i]), i
    ops = pkg.chain_ops(stages)
    sp = ptr_list([src_bufs[0].ctypes.data, None, src_bufs[2].ctypes.data])
    dp = ptr_list([b.ctypes.data for b in dst_bufs])
    assert ac.lib.acgpu_chain_frame_list_host(sp, F.IMG_YUY2, w, h, dp, ops, 1, nf) == 0 and "frame 1" in ac.last_error()
    assert ac.lib.acgpu_chain_frame_list_host(None, F.IMG_YUY2, w, h, dp, ops, 1, nf) == 0
    assert ac.lib.acgpu_chain_frame_list_host(None, F.IMG_YUY2, w, h, None, ops, 1, 0) == 1     # nothing to do


def test_chain_rejections(ac):
    w, h = 64, 32
    bad = [
        (F.IMG_YUY2, [(FLIP_V,)]),                       # not a do_process_frame layout
        (F.IMG_YUV420P, [(CLIP, 1, 0, 0, 0)]),           # odd clip on subsampled chroma
        (F.IMG_YUV420P, [(DEINTERLACE, 3)]),             # needs tcv_zoom
        (F.IMG_YUV420P, [(DEINTERLACE, 9)]),
        (F.IMG_RGB24, [(CLIP, 40, 40, 0, 0)]),           # nothing left
        (F.IMG_RGB24, [(REDUCE, 0, 1)]),
        (F.IMG_RGB24, [(CONVERT, 0x7777)]),
        (F.IMG_RGB24, [(99,)]),
    ]
    for fmt, stages in bad:
        ops = pkg.chain_ops(stages)
        ok, *_ = chain_output(ac, fmt, w, h, ops, len(stages))
        assert ok == 0 and ac.last_error(), stages
    # overlapping device buffers
    buf = ac.malloc(4 * w * h * 3)
    ops = pkg.chain_ops([(FLIP_V,)])
    assert ac.lib.acgpu_chain_batch(buf.ptr, F.IMG_RGB24, w, h, 0, buf.ptr + 16, 0, ops, 1, 1, None) == 0
    assert ac.lib.acgpu_chain_batch(buf.ptr, F.IMG_RGB24, w, h, 0, buf.ptr, 0, ops, 1, 1, None) == 0
    buf.free()


@pytest.mark.parametrize("budget,fuse", [(1, 1), (4_000_000, 0), (9_000_000, 1), (1 << 31, 0)])
def test_chain_batch_walks_sub_batches_when_the_temporary_is_small(budget, fuse):
    """7 frames of 640x480 with room for 1, 2, 4 and all frames per sub-batch (the last one partial), with and without the
    fused form of conversion pairs ($ACGPU_CHAIN_FUSE)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    e = dict(os.environ)
    e["ACGPU_CHAIN_SCRATCH_BYTES"] = str(budget)
    e["ACGPU_CHAIN_FUSE"] = str(fuse)
    e["PYTHONPATH"] = os.pathsep.join([root, here, e.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, os.path.join(here, "chain_worker.py")], capture_output=True, text=True, env=e, cwd=root, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), (r.stdout[-400:], r.stderr[-400:])
