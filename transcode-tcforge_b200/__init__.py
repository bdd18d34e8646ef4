"""ctypes binding of libacgpu -- the host-side mirror of aclib's C interface for Python callers.

The product is ``libacgpu.so`` (C ABI, include/*.h).  This module only loads it and exposes the same
entry points under the same names (``ac_init``, ``ac_imgconvert``, ``ac_average``, ``ac_rescale``, ...)
plus thin helpers for device buffers, so tests and bench.py read like the reference's own C tests
(testsuite/test-imgconvert.c, testsuite/test-average.c).  There is no Python or CPU implementation of
any pixel operation here: if the shared library is missing or no B200 is present, loading / ac_init
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import formats as F

_HERE = os.path.dirname(os.path.abspath(__file__))

LIB_PATH = os.environ.get("ACGPU_LIB") or os.path.join(_HERE, "libacgpu.so")   # ACGPU_LIB: A/B builds while profiling
TCV_LIB_PATH = os.path.join(_HERE, "libtcvgpu.so")                              # libtcvideo's interface over libacgpu

AC_NONE = 0
AC_ALL = -1
AC_SSE2 = 0x0100
AC_CUDA = 0x8000

ROW_RESCALE, ROW_AVERAGE, ROW_COPY, ROW_AVERAGE3 = 0, 1, 2, 3
DEINT_INTERPOLATE, DEINT_LINEAR_BLEND = 0, 1

_u8p = C.POINTER(C.c_uint8)
_planes_t = _u8p * 3


# include/acgpu.h row-operation codes and deinterlace modes
ACGPU_ROW_RESCALE, ACGPU_ROW_AVERAGE, ACGPU_ROW_COPY, ACGPU_ROW_AVERAGE3 = 0, 1, 2, 3
ACGPU_DEINT_INTERPOLATE, ACGPU_DEINT_LINEAR_BLEND, ACGPU_DEINT_DROP_FIELD_TOP, ACGPU_DEINT_DROP_FIELD_BOTTOM = 0, 1, 2, 3


class RowOp(C.Structure):
    _fields_ = [("src1_off", C.c_int64), ("src2_off", C.c_int64), ("src3_off", C.c_int64),
                ("dest_off", C.c_int64), ("weight1", C.c_uint32), ("weight2", C.c_uint32),
                ("op", C.c_uint32), ("reserved", C.c_uint32)]


CHAIN_CONVERT, CHAIN_CLIP, CHAIN_DEINTERLACE, CHAIN_RESIZE, CHAIN_REDUCE, CHAIN_FLIP_V, CHAIN_FLIP_H, CHAIN_RGBSWAP, \
    CHAIN_DECOLOR, CHAIN_GAMMA, CHAIN_ANTIALIAS = range(1, 12)


class ChainOp(C.Structure):
    """include/acgpu.h acgpu_chain_op."""
    _fields_ = [("kind", C.c_int32), ("p", C.c_int32 * 5), ("d", C.c_double * 2)]


def chain_ops(stages):
    """[(kind, ints..., floats...)] -> C array of acgpu_chain_op; ints fill p[], floats fill d[]."""
    arr = (ChainOp * max(len(stages), 1))()
    for i, st in enumerate(stages):
        arr[i].kind = st[0]
        ints = [v for v in st[1:] if isinstance(v, (int, np.integer))]
        flts = [v for v in st[1:] if isinstance(v, float)]
        for j, v in enumerate(ints):
            arr[i].p[j] = int(v)
        for j, v in enumerate(flts):
            arr[i].d[j] = float(v)
    return arr


class AcGpuError(RuntimeError):
    pass


#: every symbol include/*.h declares; tests/test_abi.py checks the .so exports exactly these
ABI_SYMBOLS = [
    "ac_init", "ac_cpuinfo", "ac_endian", "ac_flagstotext", "ac_parseflags", "ac_memcpy", "ac_average",
    "ac_rescale", "ac_imgconvert_init", "ac_imgconvert",
    "acgpu_version", "acgpu_last_error", "acgpu_device_count", "acgpu_set_device", "acgpu_get_device",
    "acgpu_device_sm_count", "acgpu_last_kernel_tier", "acgpu_launch_count", "acgpu_force_tier",
    "acgpu_malloc", "acgpu_free", "acgpu_host_alloc", "acgpu_host_free", "acgpu_memcpy_h2d",
    "acgpu_memcpy_d2h", "acgpu_memcpy_d2d", "acgpu_memset", "acgpu_stream_create", "acgpu_stream_destroy",
    "acgpu_stream_sync", "acgpu_event_create", "acgpu_event_destroy", "acgpu_event_record",
    "acgpu_event_sync", "acgpu_event_elapsed_ms", "acgpu_imgconvert_batch", "acgpu_imgconvert_frames_host", "acgpu_imgconvert_frames_host_multi",
    "acgpu_rowops_run", "acgpu_average", "acgpu_rescale", "acgpu_deinterlace_batch", "acgpu_resize_batch",
    "acgpu_convert_batch", "acgpu_decolor_rgb24_batch",
    "acgpu_clip_batch", "acgpu_reduce_batch", "acgpu_flip_v_batch", "acgpu_flip_h_batch",
    "acgpu_gamma_correct_batch", "acgpu_antialias_batch",
    "acgpu_chain_output", "acgpu_chain_batch", "acgpu_chain_frames_host", "acgpu_chain_frames_host_multi", "acgpu_shutdown",
    "acgpu_chain_frame_list_host", "acgpu_chain_frame_list_host_multi",
    "acgpu_bufalloc", "acgpu_buffree", "acgpu_host_register", "acgpu_host_unregister", "acgpu_pointer_kind",
]


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise AcGpuError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(libacgpu has no Python/CPU fallback)")
    lib = C.CDLL(path)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
    sig = {
        "ac_init": (i32, [i32]), "ac_cpuinfo": (i32, []), "ac_endian": (i32, []),
        "ac_flagstotext": (C.c_char_p, [i32]), "ac_parseflags": (i32, [C.c_char_p, C.POINTER(i32)]),
        "ac_memcpy": (vp, [vp, vp, sz]), "ac_average": (None, [vp, vp, vp, i32]),
        "ac_rescale": (None, [vp, vp, vp, i32, u32, u32]), "ac_imgconvert_init": (i32, [i32]),
        "ac_imgconvert": (i32, [C.POINTER(vp), i32, C.POINTER(vp), i32, i32, i32]),
        "acgpu_version": (C.c_char_p, []), "acgpu_last_error": (C.c_char_p, []),
        "acgpu_device_count": (i32, []), "acgpu_set_device": (i32, [i32]), "acgpu_get_device": (i32, []),
        "acgpu_device_sm_count": (i32, []), "acgpu_last_kernel_tier": (i32, []),
        "acgpu_launch_count": (C.c_uint64, [i32]), "acgpu_force_tier": (None, [i32]),
        "acgpu_malloc": (vp, [sz]), "acgpu_free": (None, [vp]), "acgpu_host_alloc": (vp, [sz]),
        "acgpu_host_free": (None, [vp]), "acgpu_memcpy_h2d": (i32, [vp, vp, sz, vp]),
        "acgpu_memcpy_d2h": (i32, [vp, vp, sz, vp]), "acgpu_memcpy_d2d": (i32, [vp, vp, sz, vp]),
        "acgpu_memset": (i32, [vp, i32, sz, vp]), "acgpu_stream_create": (vp, []),
        "acgpu_stream_destroy": (None, [vp]), "acgpu_stream_sync": (i32, [vp]),
        "acgpu_event_create": (vp, []), "acgpu_event_destroy": (None, [vp]),
        "acgpu_event_record": (i32, [vp, vp]), "acgpu_event_sync": (i32, [vp]),
        "acgpu_event_elapsed_ms": (C.c_float, [vp, vp]),
        "acgpu_imgconvert_batch": (i32, [C.POINTER(vp), i32, sz, C.POINTER(vp), i32, sz, i32, i32, i32, vp]),
        "acgpu_imgconvert_frames_host": (i32, [vp, i32, vp, i32, i32, i32, i32]),
        "acgpu_imgconvert_frames_host_multi": (i32, [vp, i32, vp, i32, i32, i32, i32, i32]),
        "acgpu_rowops_run": (i32, [vp, sz, vp, sz, C.POINTER(RowOp), i32, i32, i32, vp]),
        "acgpu_average": (i32, [vp, vp, vp, sz, vp]), "acgpu_rescale": (i32, [vp, vp, vp, sz, u32, u32, vp]),
        "acgpu_deinterlace_batch": (i32, [vp, vp, i32, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_resize_batch": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_convert_batch": (i32, [vp, vp, i32, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_decolor_rgb24_batch": (i32, [vp, i32, i32, sz, i32, vp]),
        "acgpu_clip_batch": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, C.c_uint8, sz, sz, i32, vp]),
        "acgpu_reduce_batch": (i32, [vp, vp, i32, i32, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_flip_v_batch": (i32, [vp, vp, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_flip_h_batch": (i32, [vp, vp, i32, i32, i32, sz, sz, i32, vp]),
        "acgpu_gamma_correct_batch": (i32, [vp, vp, i32, i32, i32, C.c_double, sz, sz, i32, vp]),
        "acgpu_antialias_batch": (i32, [vp, vp, i32, i32, i32, C.c_double, C.c_double, sz, sz, i32, vp]),
        "acgpu_chain_output": (i32, [i32, i32, i32, C.POINTER(ChainOp), i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "acgpu_chain_batch": (i32, [vp, i32, i32, i32, sz, vp, sz, C.POINTER(ChainOp), i32, i32, vp]),
        "acgpu_chain_frames_host": (i32, [vp, i32, i32, i32, vp, C.POINTER(ChainOp), i32, i32]),
        "acgpu_chain_frames_host_multi": (i32, [vp, i32, i32, i32, vp, C.POINTER(ChainOp), i32, i32, i32]),
        "acgpu_chain_frame_list_host": (i32, [C.POINTER(vp), i32, i32, i32, C.POINTER(vp), C.POINTER(ChainOp), i32, i32]),
        "acgpu_chain_frame_list_host_multi": (i32, [C.POINTER(vp), i32, i32, i32, C.POINTER(vp), C.POINTER(ChainOp), i32, i32, i32]),
        "acgpu_shutdown": (None, []),
        "acgpu_bufalloc": (vp, [sz]), "acgpu_buffree": (None, [vp]), "acgpu_host_register": (i32, [vp, sz]),
        "acgpu_host_unregister": (i32, [vp]), "acgpu_pointer_kind": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


#: every symbol include/tcvideo.h declares (libtcvideo/tcvideo.h:54-98)
TCV_ABI_SYMBOLS = ["tcv_init", "tcv_free", "tcv_clip", "tcv_deinterlace", "tcv_resize", "tcv_zoom", "tcv_reduce", "tcv_flip_v",
                   "tcv_flip_h", "tcv_gamma_correct", "tcv_antialias", "tcv_convert", "tcv_zoom_filter_to_string",
                   "tcv_zoom_filter_from_string"]


def load_tcv_library(path: str = TCV_LIB_PATH) -> C.CDLL:
    """libtcvgpu.so with the reference's prototypes (handle first, then src, dest, width, height, Bpp, ...)."""
    if not os.path.exists(path):
        raise AcGpuError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    vp, i32, u8, dbl = C.c_void_p, C.c_int, C.c_uint8, C.c_double
    plane = [vp, vp, vp, i32, i32, i32]
    sig = {
        "tcv_init": (vp, []), "tcv_free": (None, [vp]),
        "tcv_clip": (i32, plane + [i32, i32, i32, i32, u8]), "tcv_deinterlace": (i32, plane + [i32]),
        "tcv_resize": (i32, plane + [i32, i32, i32, i32]), "tcv_zoom": (i32, plane + [i32, i32, i32]),
        "tcv_reduce": (i32, plane + [i32, i32]), "tcv_flip_v": (i32, plane), "tcv_flip_h": (i32, plane),
        "tcv_gamma_correct": (i32, plane + [dbl]), "tcv_antialias": (i32, plane + [dbl, dbl]),
        "tcv_convert": (i32, [vp, vp, vp, i32, i32, i32, i32]),
        "tcv_zoom_filter_to_string": (C.c_char_p, [i32]), "tcv_zoom_filter_from_string": (i32, [C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def _vp3(ptrs):
    arr = (C.c_void_p * 3)()
    for i in range(3):
        arr[i] = ptrs[i] if i < len(ptrs) and ptrs[i] else None
    return arr


class DeviceBuffer:
    """A device allocation owned by Python (acgpu_malloc / acgpu_free)."""

    def __init__(self, lib: "AcGpu", nbytes: int):
        self._lib = lib
        self.nbytes = int(nbytes)
        self.ptr = lib.lib.acgpu_malloc(max(self.nbytes, 1))
        if not self.ptr:
            raise AcGpuError(lib.last_error())

    def upload(self, a: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        assert offset + a.size <= self.nbytes
        self._lib._ok(self._lib.lib.acgpu_memcpy_h2d(self.ptr + offset, a.ctypes.data, a.size, None))
        self._lib.sync()
        return self

    def download(self, nbytes: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, dtype=np.uint8)
        self._lib._ok(self._lib.lib.acgpu_memcpy_d2h(out.ctypes.data, self.ptr + offset, n, None))
        self._lib.sync()
        return out

    def fill(self, value: int):
        self._lib._ok(self._lib.lib.acgpu_memset(self.ptr, value, self.nbytes, None))
        self._lib.sync()
        return self

    def free(self):
        if self.ptr:
            self._lib.lib.acgpu_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory (acgpu_host_alloc) viewed as a numpy uint8 array."""

    def __init__(self, lib: "AcGpu", nbytes: int):
        self._lib = lib
        self.nbytes = int(nbytes)
        self.ptr = lib.lib.acgpu_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise AcGpuError(lib.last_error())
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.lib.acgpu_host_free(self.ptr)
            self.ptr = None


class AcGpu:
    """One loaded libacgpu.  Method names and argument order are aclib's."""

    def __init__(self, path: str = LIB_PATH):
        self.lib = load_library(path)

    # ---- aclib core -----------------------------------------------------------------------------
    def ac_init(self, accel: int = AC_ALL) -> int:
        return self.lib.ac_init(accel)

    def ac_cpuinfo(self) -> int:
        return self.lib.ac_cpuinfo()

    def ac_flagstotext(self, accel: int) -> str:
        return self.lib.ac_flagstotext(accel).decode()

    def ac_parseflags(self, text: str):
        v = C.c_int(0)
        ok = self.lib.ac_parseflags(text.encode(), C.byref(v))
        return ok, v.value

    def last_error(self) -> str:
        return self.lib.acgpu_last_error().decode()

    def _ok(self, rc: int):
        if rc != 1:
            raise AcGpuError(self.last_error())

    def sync(self, stream=None):
        self._ok(self.lib.acgpu_stream_sync(stream))

    # ---- legacy, host-pointer API on numpy buffers ------------------------------------------------
    def ac_imgconvert(self, src: np.ndarray, srcfmt: int, dest: np.ndarray, destfmt: int, w: int, h: int) -> int:
        """``src`` / ``dest`` are tightly packed frames (YUV_INIT_PLANES layout); dest is modified in place."""
        so, do = F.plane_offsets(srcfmt, w, h), F.plane_offsets(destfmt, w, h)
        sp = _vp3([src.ctypes.data + o for o in so])
        dp = _vp3([dest.ctypes.data + o for o in do])
        return self.lib.ac_imgconvert(sp, srcfmt, dp, destfmt, w, h)

    def convert(self, src: np.ndarray, srcfmt: int, destfmt: int, w: int, h: int, prefill=0x55, pad: int = 64):
        """Convenience form: returns (ok, dest) with dest pre-filled and followed by a guard band."""
        s = np.array(src, dtype=np.uint8, copy=True)
        n = F.frame_bytes(destfmt, w, h)
        if isinstance(prefill, np.ndarray):
            d = np.array(prefill[: n + pad], dtype=np.uint8, copy=True)
        else:
            d = np.full(n + pad, prefill, dtype=np.uint8)
        ok = self.ac_imgconvert(s, srcfmt, d, destfmt, w, h)
        return ok, d

    def ac_average(self, a: np.ndarray, b: np.ndarray, dest: np.ndarray | None = None) -> np.ndarray:
        d = np.zeros_like(a) if dest is None else dest
        self.lib.ac_average(a.ctypes.data, b.ctypes.data, d.ctypes.data, a.size)
        return d

    def ac_rescale(self, a: np.ndarray, b: np.ndarray, w1: int, w2: int, dest: np.ndarray | None = None) -> np.ndarray:
        d = np.zeros_like(a) if dest is None else dest
        self.lib.ac_rescale(a.ctypes.data, b.ctypes.data, d.ctypes.data, a.size, w1, w2)
        return d

    # ---- device-resident API --------------------------------------------------------------------------
    def malloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self, nbytes)

    def imgconvert_batch(self, dsrc: int, srcfmt: int, src_pitch: int, ddst: int, destfmt: int, dst_pitch: int,
                         w: int, h: int, nframes: int, stream=None) -> int:
        """Device pointers to frame 0 of tightly packed frames; planes derived like YUV_INIT_PLANES."""
        so, do = F.plane_offsets(srcfmt, w, h), F.plane_offsets(destfmt, w, h)
        return self.lib.acgpu_imgconvert_batch(_vp3([dsrc + o for o in so]), srcfmt, src_pitch,
                                               _vp3([ddst + o for o in do]), destfmt, dst_pitch,
                                               w, h, nframes, stream)

    def plane_op_batch(self, op: str, frames: np.ndarray, out_bytes: int, w: int, h: int, bpp: int, *args,
                       prefill: int = 0x55, inplace: bool = False, dst_gap: int = 0, src_gap: int = 0):
        """Upload [nframes, w*h*bpp] planes, run ``acgpu_<op>_batch`` (clip / reduce / flip_v / flip_h / gamma_correct /
        antialias / deinterlace / resize) once over the batch, download [nframes, out_bytes + dst_gap].
        Returns ``(ok, planes)``; ``inplace`` passes the source buffer as destination."""
        nf, sfb = frames.shape[0], w * h * bpp
        fn = getattr(self.lib, f"acgpu_{op}_batch")
        sp = sfb + src_gap
        hs = np.full((nf, sp), 0xEE, dtype=np.uint8)
        hs[:, :sfb] = np.ascontiguousarray(frames, dtype=np.uint8).reshape(nf, sfb)
        ds = self.malloc(nf * sp).upload(hs.reshape(-1))
        if inplace:
            dd, dp = ds, sp
        else:
            dp = out_bytes + dst_gap
            dd = self.malloc(max(nf * dp, 1)).fill(prefill)
        ok = fn(ds.ptr, dd.ptr, w, h, bpp, *args, sp, dp, nf, None)
        self.sync()
        out = dd.download(nf * dp).reshape(nf, dp) if ok else None
        ds.free()
        if not inplace:
            dd.free()
        return int(ok), out

    def convert_batch(self, frames: np.ndarray, srcfmt: int, destfmt: int, w: int, h: int, prefill: int = 0x55,
                      src_pitch: int | None = None, dst_pitch: int | None = None) -> np.ndarray:
        """Upload [nframes, frame_bytes] -> convert on device -> download [nframes, dst_pitch]."""
        nf = frames.shape[0]
        sfb, dfb = F.frame_bytes(srcfmt, w, h), F.frame_bytes(destfmt, w, h)
        sp = sfb if src_pitch is None else src_pitch
        dp = dfb if dst_pitch is None else dst_pitch
        hs = np.zeros((nf, sp), dtype=np.uint8)
        hs[:, :sfb] = frames.reshape(nf, sfb)
        ds, dd = self.malloc(nf * sp), self.malloc(nf * dp)
        ds.upload(hs.reshape(-1))
        dd.fill(prefill)
        self._ok(self.imgconvert_batch(ds.ptr, srcfmt, sp, dd.ptr, destfmt, dp, w, h, nf))
        self.sync()
        out = dd.download().reshape(nf, dp)
        ds.free()
        dd.free()
        return out
