"""CPU checkers for the parity tests (TEST INFRASTRUCTURE).

Two checkers with one calling convention:
  * ``Oracle``  -- oracle/liboracle.so, this repo's C restatement of aclib's plain-C path;
  * ``RefLib``  -- oracle/_ref/libac_ref_{c,sse2}.so, the unmodified reference compiled from
                   /root/reference/aclib by oracle/Makefile (present only where it was built;
                   the .so travels to the GPU box, /root/reference does not).
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


class Oracle(_Base):
    name = "oracle"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self._convert, self._average, self._rescale = _bind(self.lib, "oracle_")
        self.lib.oracle_resize_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        self.lib.oracle_deinterlace.restype = C.c_int
        self.lib.oracle_deinterlace.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.oracle_resize.restype = C.c_int
        self.lib.oracle_resize.argtypes = [_u8p, _u8p] + [C.c_int] * 7

    def resize_table(self, oldsize: int, newsize: int):
        n = newsize // 8
        s = (C.c_int32 * max(n, 1))()
        w1 = (C.c_uint32 * max(n, 1))()
        w2 = (C.c_uint32 * max(n, 1))()
        self.lib.oracle_resize_table(oldsize, newsize, s, w1, w2)
        return (np.array(s[:n], dtype=np.int32), np.array(w1[:n], dtype=np.uint32),
                np.array(w2[:n], dtype=np.uint32))

    def deinterlace(self, src: np.ndarray, w: int, h: int, bpp: int, mode: int) -> np.ndarray:
        s = np.array(src, dtype=np.uint8, copy=True)
        d = np.full(w * h * bpp, 0x55, dtype=np.uint8)
        assert self.lib.oracle_deinterlace(_ptr(s), _ptr(d), w, h, bpp, mode) == 1
        return d

    def resize(self, src: np.ndarray, w: int, h: int, bpp: int, rw: int, rh: int, sw: int, sh: int):
        nw, nh = w + rw * sw, h + rh * sh
        s = np.array(src, dtype=np.uint8, copy=True)
        d = np.full(nw * nh * bpp, 0x55, dtype=np.uint8)
        assert self.lib.oracle_resize(_ptr(s), _ptr(d), w, h, bpp, rw, rh, sw, sh) == 1
        return d


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
