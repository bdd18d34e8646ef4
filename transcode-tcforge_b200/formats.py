"""Image-format metadata shared by the host binding, the tests and bench.py.

Pure bookkeeping (ids, plane sizes, size units) -- no pixel arithmetic lives here.
Numeric ids are the reference's ``ImageFormat`` enum (aclib/imgconvert.h:16-40); plane sizes follow
``UV_PLANE_SIZE`` / ``YUV_INIT_PLANES`` (aclib/imgconvert.h:54-65); the width/height units are the ones
the reference's own test uses to pick legal sizes (testsuite/test-imgconvert.c:38-61).
"""
from __future__ import annotations

IMG_YUV420P = 0x1001
IMG_YV12 = 0x1002
IMG_YUV411P = 0x1003
IMG_YUV422P = 0x1004
IMG_YUV444P = 0x1005
IMG_YUY2 = 0x1006
IMG_UYVY = 0x1007
IMG_YVYU = 0x1008
IMG_Y8 = 0x1009
IMG_RGB24 = 0x2001
IMG_BGR24 = 0x2002
IMG_RGBA32 = 0x2003
IMG_ABGR32 = 0x2004
IMG_ARGB32 = 0x2005
IMG_BGRA32 = 0x2006
IMG_GRAY8 = 0x2007

NAMES = {
    IMG_YUV420P: "yuv420p", IMG_YV12: "yv12", IMG_YUV411P: "yuv411p", IMG_YUV422P: "yuv422p",
    IMG_YUV444P: "yuv444p", IMG_YUY2: "yuy2", IMG_UYVY: "uyvy", IMG_YVYU: "yvyu", IMG_Y8: "y8",
    IMG_RGB24: "rgb24", IMG_BGR24: "bgr24", IMG_RGBA32: "rgba32", IMG_ABGR32: "abgr32",
    IMG_ARGB32: "argb32", IMG_BGRA32: "bgra32", IMG_GRAY8: "gray8",
}
BY_NAME = {v: k for k, v in NAMES.items()}

#: the 15 formats of the 225-pair table, in the reference test's order (YV12 is a pointer-swap alias)
FORMATS_15 = [IMG_YUV420P, IMG_YUV411P, IMG_YUV422P, IMG_YUV444P, IMG_YUY2, IMG_UYVY, IMG_YVYU, IMG_Y8,
              IMG_RGB24, IMG_BGR24, IMG_RGBA32, IMG_ABGR32, IMG_ARGB32, IMG_BGRA32, IMG_GRAY8]
FORMATS_16 = [IMG_YUV420P, IMG_YV12] + FORMATS_15[1:]

PLANAR = (IMG_YUV420P, IMG_YV12, IMG_YUV411P, IMG_YUV422P, IMG_YUV444P)

# (width unit, height unit): smallest block on which the format is meaningful
UNITS = {
    IMG_YUV420P: (2, 2), IMG_YV12: (2, 2), IMG_YUV411P: (4, 1), IMG_YUV422P: (2, 1), IMG_YUV444P: (1, 1),
    IMG_YUY2: (2, 1), IMG_UYVY: (2, 1), IMG_YVYU: (2, 1), IMG_Y8: (1, 1),
    IMG_RGB24: (1, 1), IMG_BGR24: (1, 1), IMG_RGBA32: (1, 1), IMG_ABGR32: (1, 1), IMG_ARGB32: (1, 1),
    IMG_BGRA32: (1, 1), IMG_GRAY8: (1, 1),
}


def uv_plane_size(fmt: int, w: int, h: int) -> int:
    if fmt in (IMG_YUV420P, IMG_YV12):
        return (w // 2) * (h // 2)
    if fmt == IMG_YUV411P:
        return (w // 4) * h
    if fmt == IMG_YUV422P:
        return (w // 2) * h
    if fmt == IMG_YUV444P:
        return w * h
    return 0


def plane_sizes(fmt: int, w: int, h: int) -> list[int]:
    """Bytes per plane; one entry for packed formats, three for planar YUV."""
    p = w * h
    if fmt in PLANAR:
        c = uv_plane_size(fmt, w, h)
        return [p, c, c]
    if fmt in (IMG_YUY2, IMG_UYVY, IMG_YVYU):
        return [2 * p]
    if fmt in (IMG_Y8, IMG_GRAY8):
        return [p]
    if fmt in (IMG_RGB24, IMG_BGR24):
        return [3 * p]
    return [4 * p]


def frame_bytes(fmt: int, w: int, h: int) -> int:
    return sum(plane_sizes(fmt, w, h))


def plane_offsets(fmt: int, w: int, h: int) -> list[int]:
    """Offsets of each plane inside one tightly packed frame buffer (YUV_INIT_PLANES order)."""
    offs, o = [], 0
    for s in plane_sizes(fmt, w, h):
        offs.append(o)
        o += s
    return offs


def size_unit(srcfmt: int, dstfmt: int) -> tuple[int, int]:
    """Coarser of the two formats' units -- the domain on which parity with the C path is defined."""
    a, b = UNITS[srcfmt], UNITS[dstfmt]
    return max(a[0], b[0]), max(a[1], b[1])


def algorithmic_bytes(srcfmt: int, dstfmt: int, w: int, h: int) -> int:
    """Bytes a single fused pass must read plus bytes it must write.

    yuv/Y8 -> 32-bit RGB leaves alpha untouched in the reference (img_yuv_rgb.c:62-64); libacgpu
    implements that as read-modify-write of the destination word, so those pairs count 4 B read + 4 B
    written per pixel on the RGB side.
    """
    rd = frame_bytes(srcfmt, w, h)
    wr = frame_bytes(dstfmt, w, h)
    if dstfmt in (IMG_RGBA32, IMG_ABGR32, IMG_ARGB32, IMG_BGRA32) and (srcfmt >> 12) == 1:
        rd += wr
    return rd + wr
