"""CPU checkers for the parity tests (TEST INFRASTRUCTURE).

Two checkers with one calling convention:
  * ``Oracle``  -- oracle/liboracle.so, this repo's C restatement of aclib's plain-C path;
  * ``RefLib``  -- oracle/_ref/libac_ref_{c,sse2}.so, the unmodified reference compiled from
                   /root/reference/aclib by oracle/Makefile (present only where it was built;
                   the .so travels to the GPU box, /root/reference does not).
and two with the libtcvideo calling convention (one plane, Bpp 1 or 3):
  * ``Oracle``  again (oracle_deinterlace / resize / clip / reduce / flip / gamma / antialias);
  * ``TcvRef``  -- oracle/_ref/libtcv_ref.so, the unmodified reference libtcvideo (tcvideo.c + zoom.c) over the
                   plain-C aclib, with oracle/tcv_ref_stubs.c standing in for libtc's allocator and logger.
Nothing under transcode-tcforge_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as _entry  # noqa: E402

F = _entry.load_package().F     # format ids / plane sizes only; no pixel code comes from the product

_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray, off: int = 0):
    return C.cast(a.ctypes.data + off, _u8p)


def _planes(buf: np.ndarray, fmt: int, w: int, h: int):
    arr = (_u8p * 3)()
    offs = F.plane_offsets(fmt, w, h)
    for i in range(3):
        arr[i] = _ptr(buf, offs[i] if i < len(offs) else 0)
    return arr


def build_oracle() -> None:
    """Compile oracle/liboracle.so + cpubench (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


class _Base:
    name = "?"

    def convert(self, src: np.ndarray, srcfmt: int, dstfmt: int, w: int, h: int,
                prefill: int | np.ndarray = 0x55, pad: int = 64):
        """Run one conversion on tightly packed frames.

        ``src`` holds one frame (YUV_INIT_PLANES layout).  Returns ``(ok, dest)`` where dest has
        ``frame_bytes(dstfmt) + pad`` bytes pre-filled with ``prefill`` so untouched bytes and overruns
        are visible.  The source is copied first: the reference rewrites UYVY/YVYU sources in place
        (aclib/img_yuv_mixed.c:24,30-32).
        """
        s = np.array(src, dtype=np.uint8, copy=True)
        n = F.frame_bytes(dstfmt, w, h)
        if isinstance(prefill, np.ndarray):
            d = np.array(prefill[: n + pad], dtype=np.uint8, copy=True)
            assert d.size == n + pad
        else:
            d = np.full(n + pad, prefill, dtype=np.uint8)
        ok = self._convert(_planes(s, srcfmt, w, h), srcfmt, _planes(d, dstfmt, w, h), dstfmt, w, h)
        return int(ok), d

    def average(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        d = np.zeros_like(a)
        self._average(_ptr(a), _ptr(b), _ptr(d), a.size)
        return d

    def rescale(self, a: np.ndarray, b: np.ndarray, w1: int, w2: int) -> np.ndarray:
        d = np.zeros_like(a)
        self._rescale(_ptr(a), _ptr(b), _ptr(d), a.size, w1, w2)
        return d


def _bind(lib, prefix):
    conv = getattr(lib, prefix + "imgconvert")
    conv.restype = C.c_int
    conv.argtypes = [C.POINTER(_u8p), C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int]
    avg = getattr(lib, prefix + "average")
    avg.restype = None
    avg.argtypes = [_u8p, _u8p, _u8p, C.c_int]
    rs = getattr(lib, prefix + "rescale")
    rs.restype = None
    rs.argtypes = [_u8p, _u8p, _u8p, C.c_int, C.c_uint32, C.c_uint32]
    return conv, avg, rs


class _TcvOps:
    """libtcvideo-shaped operations on one tightly packed plane (Bpp 1 or 3).  Subclasses provide ``_tcv(name)`` ->
    callable(src_ptr, dest_ptr, *args) -> int.  Deinterlace modes use libacgpu's numbering: 0 interpolate,
    1 linear blend, 2 drop top field, 3 drop bottom field."""

    def _plane_op(self, name, src, out_bytes, *args, prefill=0x55, inplace=False):
        s = np.array(src, dtype=np.uint8, copy=True)
        d = s if inplace else np.full(max(out_bytes, 1), prefill, dtype=np.uint8)
        ok = self._tcv(name)(_ptr(s), _ptr(d), *args)
        return int(ok), d[:out_bytes] if not inplace else d

    def deinterlace(self, src: np.ndarray, w: int, h: int, bpp: int, mode: int) -> np.ndarray:
        rows = h // 2 if mode >= 2 else h
        ok, d = self._plane_op("deinterlace", src, w * rows * bpp, w, h, bpp, mode)
        assert ok == 1
        return d

    def resize(self, src: np.ndarray, w: int, h: int, bpp: int, rw: int, rh: int, sw: int, sh: int):
        nw, nh = w + rw * sw, h + rh * sh
        ok, d = self._plane_op("resize", src, nw * nh * bpp, w, h, bpp, rw, rh, sw, sh)
        assert ok == 1
        return d

    def clip(self, src, w, h, bpp, left, right, top, bottom, black=0, prefill=0x55):
        nw, nh = w - left - right, h - top - bottom
        return self._plane_op("clip", src, max(nw, 0) * max(nh, 0) * bpp, w, h, bpp, left, right, top, bottom, black,
                              prefill=prefill)

    def reduce(self, src, w, h, bpp, rw, rh, prefill=0x55):
        if rw <= 0 or rh <= 0:
            n = w * h * bpp
        elif rw == 1 and rh == 1:
            n = w * h * bpp
        elif rw == 1:
            n = w * (h // rh) * bpp
        else:
            n = (w // rw) * (h // rh) * bpp
        return self._plane_op("reduce", src, n, w, h, bpp, rw, rh, prefill=prefill)

    def flip_v(self, src, w, h, bpp, inplace=False):
        return self._plane_op("flip_v", src, w * h * bpp, w, h, bpp, inplace=inplace)

    def flip_h(self, src, w, h, bpp, inplace=False):
        return self._plane_op("flip_h", src, w * h * bpp, w, h, bpp, inplace=inplace)

    def gamma(self, src, w, h, bpp, gamma):
        return self._plane_op("gamma", src, w * h * bpp, w, h, bpp, float(gamma))

    def antialias(self, src, w, h, bpp, weight, bias):
        return self._plane_op("antialias", src, w * h * bpp, w, h, bpp, float(weight), float(bias))


_TCV_SIGS = {   # argument types after (src, dest)
    "deinterlace": [C.c_int] * 4,
    "resize": [C.c_int] * 7,
    "clip": [C.c_int] * 7 + [C.c_uint8],
    "reduce": [C.c_int] * 5,
    "flip_v": [C.c_int] * 3,
    "flip_h": [C.c_int] * 3,
    "gamma": [C.c_int] * 3 + [C.c_double],
    "antialias": [C.c_int] * 3 + [C.c_double, C.c_double],
}


class Oracle(_Base, _TcvOps):
    name = "oracle"
    _NAMES = {"gamma": "oracle_gamma_correct"}

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self._convert, self._average, self._rescale = _bind(self.lib, "oracle_")
        self.lib.oracle_resize_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        for op, sig in _TCV_SIGS.items():
            fn = getattr(self.lib, self._NAMES.get(op, "oracle_" + op))
            fn.restype = C.c_int
            fn.argtypes = [_u8p, _u8p] + sig
        self.lib.oracle_gamma_table.argtypes = [C.c_double, _u8p]
        self.lib.oracle_aa_tables.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_uint32)]

    def _tcv(self, op):
        return getattr(self.lib, self._NAMES.get(op, "oracle_" + op))

    def resize_table(self, oldsize: int, newsize: int):
        n = newsize // 8
        s = (C.c_int32 * max(n, 1))()
        w1 = (C.c_uint32 * max(n, 1))()
        w2 = (C.c_uint32 * max(n, 1))()
        self.lib.oracle_resize_table(oldsize, newsize, s, w1, w2)
        return (np.array(s[:n], dtype=np.int32), np.array(w1[:n], dtype=np.uint32),
                np.array(w2[:n], dtype=np.uint32))

    def gamma_table(self, gamma: float) -> np.ndarray:
        t = np.zeros(256, np.uint8)
        self.lib.oracle_gamma_table(float(gamma), _ptr(t))
        return t

    def aa_tables(self, weight: float, bias: float) -> np.ndarray:
        t = np.zeros(1024, np.uint32)
        self.lib.oracle_aa_tables(float(weight), float(bias), t.ctypes.data_as(C.POINTER(C.c_uint32)))
        return t


class TcvRef(_TcvOps):
    """The unmodified reference libtcvideo (oracle/_ref/libtcv_ref.so); aclib inside it runs its plain-C path."""
    name = "ref_tcv"
    _NAMES = {"gamma": "tcv_gamma_correct"}
    _MODES = {0: 2, 1: 3, 2: 0, 3: 1}    # libacgpu numbering -> TCVDeinterlaceMode (libtcvideo/tcvideo.h:29-34)

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libtcv_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ac_init.restype = C.c_int
        self.lib.ac_init.argtypes = [C.c_int]
        if self.lib.ac_init(0) != 1:
            raise RuntimeError("reference ac_init failed")
        self.lib.tcv_init.restype = C.c_void_p
        self.handle = C.c_void_p(self.lib.tcv_init())
        assert self.handle
        for op, sig in _TCV_SIGS.items():
            fn = getattr(self.lib, self._NAMES.get(op, "tcv_" + op))
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, _u8p, _u8p] + sig

    def _tcv(self, op):
        fn = getattr(self.lib, self._NAMES.get(op, "tcv_" + op))
        if op == "deinterlace":
            return lambda s, d, w, h, bpp, mode: fn(self.handle, s, d, w, h, bpp, self._MODES[mode])
        return lambda s, d, *a: fn(self.handle, s, d, *a)


def have_tcv_ref() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libtcv_ref.so"))


def best_tcv_checker():
    """The real libtcvideo when its build travelled with the repo, else the restatement."""
    return TcvRef() if have_tcv_ref() else Oracle()


class RefLib(_Base):
    """The unmodified reference; ``variant`` is "c" (AC_NONE) or "sse2" (AC_ALL)."""

    def __init__(self, variant: str = "c"):
        self.name = "ref_" + variant
        path = os.path.join(ORACLE_DIR, "_ref", f"libac_ref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ac_init.restype = C.c_int
        self.lib.ac_init.argtypes = [C.c_int]
        if self.lib.ac_init(0 if variant == "c" else -1) != 1:
            raise RuntimeError("reference ac_init failed")
        self._convert, self._average, self._rescale = _bind(self.lib, "ac_")


def have_ref(variant: str = "c") -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", f"libac_ref_{variant}.so"))


def best_checker():
    """The real reference when its build travelled with the repo, else the restatement."""
    return RefLib("c") if have_ref("c") else Oracle()


# ---------------------------------------------------------------------------------------------
# Deterministic inputs

def splitmix_bytes(n: int, seed: int = 0) -> np.ndarray:
    """n uniform bytes from a counter-based splitmix64 stream: byte i = low byte of mix(seed + (i+1)*phi).
    Self-contained so the committed golden digests do not depend on numpy's RNG versioning."""
    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(seed & 0xFFFFFFFFFFFFFFFF))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(0xFF)).astype(np.uint8)


def random_frame(fmt: int, w: int, h: int, seed: int = 0) -> np.ndarray:
    """Uniform random bytes (includes out-of-range YUV so every clamp path is hit)."""
    return splitmix_bytes(F.frame_bytes(fmt, w, h), seed * 1000003 + fmt * 31 + w * 7 + h)


def colour_bars_yuv420p(w: int, h: int) -> np.ndarray:
    """The reference end-to-end test's colour-bar frame (testsuite/newtest.pl:1462-1502), restated."""
    bar = (w // 64) * 16
    white = w - 3 * bar
    h1 = (h // 64) * 16
    right = [81] * bar + [145] * bar + [41] * bar
    rows = [np.array([g] * white + right, dtype=np.uint8) for g in (16, 89, 162, 235)]
    ys = [rows[0]] * h1 + [rows[1]] * h1 + [rows[2]] * h1 + [rows[3]] * (h - 3 * h1)
    urow = np.array([128] * ((w - 3 * bar) // 2) + [91] * (bar // 2) + [54] * (bar // 2) + [240] * (bar // 2), dtype=np.uint8)
    vrow = np.array([128] * ((w - 3 * bar) // 2) + [239] * (bar // 2) + [35] * (bar // 2) + [111] * (bar // 2), dtype=np.uint8)
    return np.concatenate(ys + [urow] * (h // 2) + [vrow] * (h // 2))


def colour_bars_rgb24(w: int, h: int) -> np.ndarray:
    """RGB twin of the colour bars (testsuite/newtest.pl:1504-1537); test_raw_raw_csp (:544-566)
    requires the two to convert into each other exactly."""
    bar = (w // 64) * 16
    white = w - 3 * bar
    h1 = (h // 64) * 16
    colour = [253, 0, 1] * bar + [2, 255, 1] * bar + [2, 0, 255] * bar
    rows = [np.array([g] * (white * 3) + colour, dtype=np.uint8) for g in (0, 85, 170, 255)]
    return np.concatenate([rows[0]] * h1 + [rows[1]] * h1 + [rows[2]] * h1 + [rows[3]] * (h - 3 * h1))


def glibc_random_bytes(n: int, seed: int = 0) -> np.ndarray:
    """glibc random() TYPE_3 low bytes == the reference test's generator
    (testsuite/test-imgconvert.c:361-363: srandom(0); buf[i] = random())."""
    libc = C.CDLL(None)
    libc.srandom(C.c_uint(seed))
    libc.random.restype = C.c_long
    out = np.empty(n, dtype=np.uint8)
    for i in range(n):
        out[i] = libc.random() & 0xFF
    return out
