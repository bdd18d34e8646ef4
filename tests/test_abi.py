"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/*.h
declares, keeps aclib's flag vocabulary (aclib/accore.c:76-167) and fails loudly without a device."""
import os
import re
import subprocess

import pytest

import __graft_entry__ as entry

pkg = entry.load_package()
ROOT = entry.ROOT


@pytest.fixture(scope="module")
def ac():
    if not os.path.exists(pkg.LIB_PATH):
        entry.build()
    return pkg.AcGpu()


def declared_symbols():
    names = set()
    for hdr in ("ac.h", "imgconvert.h", "acgpu.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(ac_[a-z_0-9]+|acgpu_[a-z_0-9]+)\s*\(", text))
    return {n for n in names if not n.endswith("_t")}


def test_library_exports_every_declared_symbol(ac):
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    decl = declared_symbols()
    assert decl == set(pkg.ABI_SYMBOLS), decl ^ set(pkg.ABI_SYMBOLS)
    assert decl <= exported, decl - exported
    # nothing but the C ABI leaks out
    assert all(s.startswith(("ac_", "acgpu_")) for s in exported), exported


def test_reference_symbols_are_all_present(ac):
    """The ten functions of aclib/ac.h:59-91 + aclib/imgconvert.h:81-90."""
    for name in ["ac_init", "ac_cpuinfo", "ac_endian", "ac_flagstotext", "ac_parseflags", "ac_memcpy",
                 "ac_average", "ac_rescale", "ac_imgconvert_init", "ac_imgconvert"]:
        assert hasattr(ac.lib, name)


def test_flag_vocabulary(ac):
    assert ac.ac_flagstotext(0) == "none"
    assert ac.ac_flagstotext(pkg.AC_CUDA) == "cuda"
    assert ac.ac_flagstotext(0x8000 | 0x0100 | 0x0080 | 0x0008 | 0x0002) == "cuda sse2 sse mmx asm"
    assert ac.ac_parseflags("cuda") == (1, 0x8000)
    assert ac.ac_parseflags("CUDA,sse2,mmx") == (1, 0x8108)
    assert ac.ac_parseflags("C") == (1, 0)
    assert ac.ac_parseflags("asm") == (1, 0x0002)
    assert ac.ac_parseflags("sse6")[0] == 0
    assert ac.ac_parseflags("")[0] == 0
    for text in ["sse5", "sse4a", "sse42", "sse41", "ssse3", "sse3", "sse2", "sse", "3dnowext", "3dnow", "mmxext", "mmx", "cmove"]:
        ok, v = ac.ac_parseflags(text)
        assert ok == 1 and ac.ac_flagstotext(v) == text


def test_endianness(ac):
    import sys
    assert ac.lib.ac_endian() == (1 if sys.byteorder == "little" else 2)


def test_host_memcpy_is_memmove(ac):
    """ac.h:80-82: ascending copy, so overlapping dest<src works (decode_lavc.c:296-305 relies on it)."""
    import numpy as np
    a = np.arange(64, dtype=np.uint8)
    ac.lib.ac_memcpy(a.ctypes.data, a.ctypes.data + 1, 63)
    assert list(a[:63]) == list(range(1, 64))


def test_fails_loudly_without_a_device(ac):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert ac.ac_cpuinfo() == 0
    assert ac.ac_init(pkg.AC_ALL) == 0
    assert "no CPU fallback" in ac.last_error()
    import numpy as np
    src = np.zeros(64 * 16 * 3 // 2, dtype=np.uint8)
    ok, _ = ac.convert(src, pkg.F.IMG_YUV420P, pkg.F.IMG_RGB24, 64, 16)
    assert ok == 0       # never silently computed on the CPU


def test_frame_entry_points_reject_what_libtcvideo_rejects(ac):
    """Host logic, no device needed: the parameter checks of tcv_clip / tcv_reduce / tcv_flip_* / tcv_gamma_correct /
    tcv_antialias / tcv_deinterlace / tcv_resize (libtcvideo/tcvideo.c:192-202, 291-311, 436-457, 690-699, 848-851,
    899-903) run before anything touches the GPU, return 0 and say why."""
    L = ac.lib
    p = 0x1000          # never dereferenced: every call below is rejected by its argument checks

    def why():
        return L.acgpu_last_error().decode()

    assert L.acgpu_clip_batch(p, p, 64, 32, 2, 0, 0, 0, 0, 0, 0, 0, 1, None) == 0 and "invalid frame parameters" in why()
    assert L.acgpu_clip_batch(p, p, 64, 32, 1, 40, 24, 0, 0, 0, 0, 0, 1, None) == 0 and "clipping parameters" in why()
    assert L.acgpu_clip_batch(p, p, 64, 32, 3, 0, 0, 33, -1, 0, 0, 0, 1, None) == 0 and "clipping parameters" in why()
    assert L.acgpu_clip_batch(None, p, 64, 32, 1, 0, 0, 0, 0, 0, 0, 0, 1, None) == 0 and "invalid frame parameters" in why()
    assert L.acgpu_reduce_batch(p, p, 64, 32, 1, 0, 1, 0, 0, 1, None) == 0 and "reduction parameters" in why()
    assert L.acgpu_reduce_batch(p, p, 64, 32, 1, 2, -3, 0, 0, 1, None) == 0 and "reduction parameters" in why()
    assert L.acgpu_flip_v_batch(p, p, 0, 32, 1, 0, 0, 1, None) == 0 and "invalid frame parameters" in why()
    assert L.acgpu_flip_h_batch(p, None, 64, 32, 3, 0, 0, 1, None) == 0 and "invalid frame parameters" in why()
    assert L.acgpu_gamma_correct_batch(p, p, 64, 32, 1, 0.0, 0, 0, 1, None) == 0 and "invalid gamma" in why()
    assert L.acgpu_gamma_correct_batch(p, p, 64, 32, 1, float("nan"), 0, 0, 1, None) == 0 and "invalid gamma" in why()
    assert L.acgpu_antialias_batch(p, p + 4096, 64, 32, 1, 1.01, 0.5, 0, 0, 1, None) == 0 and "antialiasing parameters" in why()
    assert L.acgpu_antialias_batch(p, p + 4096, 64, 32, 1, 0.5, -0.01, 0, 0, 1, None) == 0 and "antialiasing parameters" in why()
    assert L.acgpu_antialias_batch(p, p, 64, 32, 1, 0.5, 0.5, 0, 0, 1, None) == 0 and "overlap" in why()
    assert L.acgpu_deinterlace_batch(p, p, 64, 32, 1, 7, 0, 0, 1, None) == 0 and "invalid mode" in why()
    assert L.acgpu_deinterlace_batch(p, p, 64, 32, 4, 0, 0, 0, 1, None) == 0 and "invalid frame parameters" in why()
    assert L.acgpu_resize_batch(p, p, 64, 32, 1, 0, -1, 3, 8, 0, 0, 1, None) == 0 and "scale" in why()
    assert L.acgpu_resize_batch(p, p, 60, 32, 1, 0, -1, 8, 8, 0, 0, 1, None) == 0 and "divide" in why()
    assert L.acgpu_resize_batch(p, p, 64, 32, 1, -8, 0, 8, 8, 0, 0, 1, None) == 0 and "not positive" in why()


def test_product_never_touches_the_oracle():
    """The product path must not import, link or load anything under oracle/."""
    for dirpath, _, files in os.walk(entry.PKG_DIR):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile", ".map")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("liboracle", "ac_oracle", "oracle_", "oracle/", "libac_ref", "checkers"):
                    assert needle not in text, f"{f} references {needle}"
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libac_ref" not in out
    # the profiling aids under tools/ drive the product only; scripts that need the reference live under tests/
    tools = os.path.join(ROOT, "tools")
    for f in os.listdir(tools):
        if f.endswith((".py", ".sh", ".c", ".cu")):
            text = open(os.path.join(tools, f), errors="ignore").read()
            for needle in ("liboracle", "ac_oracle", "oracle_", "oracle/", "libac_ref", "libtcv_ref", "import checkers"):
                assert needle not in text, f"tools/{f} references {needle}"


def test_chain_planning_needs_no_device(ac):
    """acgpu_chain_output walks a stage list on the host: geometry as do_process_frame computes it
    (src/video_trans.c:213-345), rejections with a reason."""
    import ctypes as C
    F = pkg.F

    def plan(fmt, w, h, stages):
        ops = pkg.chain_ops(stages)
        of, ow, oh = C.c_int(0), C.c_int(0), C.c_int(0)
        ok = ac.lib.acgpu_chain_output(fmt, w, h, ops, len(stages), C.byref(of), C.byref(ow), C.byref(oh))
        return ok, of.value, ow.value, oh.value

    assert plan(F.IMG_YUV420P, 3840, 2160, [(pkg.CHAIN_CONVERT, F.IMG_RGB24), (pkg.CHAIN_CONVERT, F.IMG_YUV422P)]) == (1, F.IMG_YUV422P, 3840, 2160)
    assert plan(F.IMG_YUV420P, 720, 576, [(pkg.CHAIN_CLIP, 8, 8, 16, 16), (pkg.CHAIN_DEINTERLACE, 1), (pkg.CHAIN_RESIZE, -6, -4),
                                          (pkg.CHAIN_FLIP_V,), (pkg.CHAIN_GAMMA, 0.8)]) == (1, F.IMG_YUV420P, 656, 512)
    assert plan(F.IMG_YUV422P, 352, 288, [(pkg.CHAIN_DEINTERLACE, 4), (pkg.CHAIN_RESIZE, 4, 2), (pkg.CHAIN_CLIP, -16, -16, -8, -8),
                                          (pkg.CHAIN_REDUCE, 2, 2)]) == (1, F.IMG_YUV422P, 208, 88)
    assert plan(F.IMG_RGB24, 64, 32, []) == (1, F.IMG_RGB24, 64, 32)
    for fmt, stages, why in [
        (F.IMG_YUY2, [(pkg.CHAIN_FLIP_V,)], "YUV420P, YUV422P, RGB24"),
        (F.IMG_Y8, [(pkg.CHAIN_DECOLOR,)], "colour planes"),
        (F.IMG_YUV420P, [(pkg.CHAIN_CLIP, 1, 0, 0, 0)], "multiples"),
        (F.IMG_YUV420P, [(pkg.CHAIN_DEINTERLACE, 3)], "tcv_zoom"),
        (F.IMG_RGB24, [(pkg.CHAIN_CLIP, 40, 40, 0, 0)], "no frame"),
        (F.IMG_RGB24, [(pkg.CHAIN_REDUCE, 0, 1)], "reduce"),
        (F.IMG_RGB24, [(pkg.CHAIN_CONVERT, 0x7777)], "unknown format"),
        (F.IMG_RGB24, [(99,)], "unknown stage"),
    ]:
        ok, *_ = plan(fmt, 64, 32, stages)
        assert ok == 0 and why in ac.last_error(), (stages, ac.last_error())
