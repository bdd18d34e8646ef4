"""Reference side of the frame-chain parity tests (TEST INFRASTRUCTURE).

``ref_chain`` walks a stage list exactly as transcode's do_process_frame does (src/video_trans.c:192-426): the
PROCESS_FRAME stages call the libtcvideo function once per plane with the plane's divided size and arguments
(:37-46), the first-plane-only stages leave the other planes alone, -k / -K / conversions use the aclib calls the
reference uses.  Every pixel operation is executed by the CPU checkers of tests/checkers.py -- the unmodified reference
libraries when their build travelled with the repo (oracle/_ref), else the oracle restatement; nothing here computes
pixels itself.
"""
from __future__ import annotations

import numpy as np

import __graft_entry__ as entry
from checkers import F

pkg = entry.load_package()

# kinds (include/acgpu.h)
CONVERT, CLIP, DEINTERLACE, RESIZE, REDUCE, FLIP_V, FLIP_H, RGBSWAP, DECOLOR, GAMMA, ANTIALIAS = range(1, 12)


def plane_set(fmt):
    """(nplanes, Bpp, width_div, height_div, black) as set_vtd (video_trans.c:71-118)."""
    if fmt == F.IMG_YUV420P:
        return 3, 1, (1, 2, 2), (1, 2, 2), (0, 128, 128)
    if fmt == F.IMG_YUV422P:
        return 3, 1, (1, 2, 2), (1, 1, 1), (0, 128, 128)
    if fmt == F.IMG_RGB24:
        return 1, 3, (1,), (1,), (0,)
    if fmt in (F.IMG_Y8, F.IMG_GRAY8):       # set_vtd's defaults: one plane, one byte per pixel (video_trans.c:77-84)
        return 1, 1, (1,), (1,), (0,)
    raise ValueError("layout not handled by do_process_frame")


def split(frame, fmt, w, h):
    n, bpp, wd, hd, _ = plane_set(fmt)
    out, off = [], 0
    for i in range(n):
        sz = (w // wd[i]) * (h // hd[i]) * bpp
        out.append(np.array(frame[off:off + sz], dtype=np.uint8, copy=True))
        off += sz
    return out


def cdiv(a, b):
    """C integer division (truncates toward zero)."""
    q = abs(a) // b
    return q if a >= 0 else -q


def ref_chain(tcv, conv, frame, fmt, w, h, stages):
    """Returns (frame, fmt, w, h) after the stages; ``tcv`` / ``conv`` are the libtcvideo-shaped and the aclib-shaped
    checker."""
    frame = np.array(frame, dtype=np.uint8, copy=True)
    for st in stages:
        kind = st[0]
        if kind == CONVERT:
            dfmt = st[1]
            if dfmt == fmt:
                continue
            ok, out = conv.convert(frame, fmt, dfmt, w, h, prefill=0x55, pad=0)
            assert ok == 1
            frame, fmt = out, dfmt
            continue
        n, bpp, wd, hd, black = plane_set(fmt)
        planes = split(frame, fmt, w, h)
        dims = [(w // wd[i], h // hd[i]) for i in range(n)]
        if kind == CLIP:
            l, r, t, b = st[1:5]
            new = []
            for i in range(n):
                ok, d = tcv.clip(planes[i], dims[i][0], dims[i][1], bpp, cdiv(l, wd[i]), cdiv(r, wd[i]), cdiv(t, hd[i]), cdiv(b, hd[i]), black[i])
                assert ok == 1
                new.append(d)
            planes, w, h = new, w - l - r, h - t - b
        elif kind == DEINTERLACE:
            mode = st[1]
            if mode == 2:
                pass
            elif mode == 4:
                planes = [tcv.deinterlace(planes[i], dims[i][0], dims[i][1], bpp, 3) for i in range(n)]
                h //= 2
            else:
                planes[0] = tcv.deinterlace(planes[0], dims[0][0], dims[0][1], bpp, 0 if mode == 1 else 1)
        elif kind == RESIZE:
            rw, rh = st[1:3]
            if rh:
                planes = [tcv.resize(planes[i], dims[i][0], dims[i][1], bpp, 0, rh, 8 // wd[i], 8 // hd[i]) for i in range(n)]
                h += rh * 8
                dims = [(w // wd[i], h // hd[i]) for i in range(n)]
            if rw:
                planes = [tcv.resize(planes[i], dims[i][0], dims[i][1], bpp, rw, 0, 8 // wd[i], 8 // hd[i]) for i in range(n)]
                w += rw * 8
        elif kind == REDUCE:
            rw, rh = st[1:3]
            new = []
            for i in range(n):
                ok, d = tcv.reduce(planes[i], dims[i][0], dims[i][1], bpp, rw, rh)
                assert ok == 1
                new.append(d)
            planes, w, h = new, w // rw, h // rh
        elif kind in (FLIP_V, FLIP_H):
            fn = tcv.flip_v if kind == FLIP_V else tcv.flip_h
            new = []
            for i in range(n):
                ok, d = fn(planes[i], dims[i][0], dims[i][1], bpp)
                assert ok == 1
                new.append(d)
            planes = new
        elif kind == RGBSWAP:
            if fmt == F.IMG_RGB24:      # the byte loop of video_trans.c:352-358 is exactly RGB24 -> BGR24
                ok, d = conv.convert(planes[0], F.IMG_RGB24, F.IMG_BGR24, w, h, pad=0)
                assert ok == 1
                planes = [d]
            else:
                planes = [planes[0], planes[2], planes[1]]
        elif kind == DECOLOR:
            if fmt == F.IMG_RGB24:
                _, g = conv.convert(planes[0], F.IMG_RGB24, F.IMG_GRAY8, w, h, pad=0)
                _, d = conv.convert(g, F.IMG_GRAY8, F.IMG_RGB24, w, h, pad=0)
                planes = [d]
            else:
                planes = [planes[0], np.full_like(planes[1], 128), np.full_like(planes[2], 128)]
        elif kind == GAMMA:
            ok, d = tcv.gamma(planes[0], dims[0][0], dims[0][1], bpp, st[1])
            assert ok == 1
            planes[0] = d
        elif kind == ANTIALIAS:
            ok, d = tcv.antialias(planes[0], dims[0][0], dims[0][1], bpp, st[1], st[2])
            assert ok == 1
            planes[0] = d
        else:
            raise ValueError(kind)
        frame = np.concatenate(planes)
    return frame, fmt, w, h
