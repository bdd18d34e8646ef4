"""Frame-buffer plumbing: transcode's frame buffers on libacgpu's page-locked allocator.

tests/c/frame_caller.c is libtc/tcframes.c in miniature with ONE change, `#define tc_bufalloc acgpu_bufalloc`
(libtcutil/memutils.c:89-123 is the reference allocator).  CPU: it compiles and links against include/ + libacgpu +
libtcvgpu, and the allocator keeps tc_bufalloc's contract without a device.  GPU: the frame's buffers are page-locked,
the legacy per-frame calls on them give the checker's bytes, and a pre-existing buffer can be registered once."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry
from test_drop_in_c import fnv, lcg_bytes

ROOT = entry.ROOT
pkg = entry.load_package()
SRC = os.path.join(ROOT, "tests", "c", "frame_caller.c")
EXE = os.path.join(ROOT, "tests", "c", "frame_caller")


def build():
    if not os.path.exists(pkg.TCV_LIB_PATH):
        entry.build()
    subprocess.run(["gcc", "-std=gnu99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", EXE, SRC,
                    "-L", entry.PKG_DIR, "-ltcvgpu", "-lacgpu", "-Wl,-rpath," + entry.PKG_DIR, "-Wl,--no-undefined"], check=True)


def test_frame_caller_compiles_and_links():
    build()
    out = subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True, check=True).stdout
    assert {"acgpu_bufalloc", "acgpu_buffree", "acgpu_host_register", "ac_imgconvert", "tcv_flip_v", "tcv_convert",
            "acgpu_chain_frame_list_host"} <= \
        {l.split()[-1] for l in out.splitlines()}


def test_bufalloc_keeps_the_reference_contract_without_a_device():
    """tc_bufalloc promises a page-aligned writable buffer that tc_buffree releases (memutils.c:89-123)."""
    lib = pkg.load_library()
    for size in (1, 4096, 720 * 576 * 3 + 128):
        p = lib.acgpu_bufalloc(size)
        assert p and p % 4096 == 0
        C.memset(p, 0x5A, size)
        assert (np.ctypeslib.as_array((C.c_uint8 * size).from_address(p)) == 0x5A).all()
        lib.acgpu_buffree(p)
    lib.acgpu_buffree(None)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(352, 288), (720, 576)])
def test_frame_buffers_are_page_locked_and_calls_match_the_checker(size):
    import checkers as ck
    F = ck.F
    build()
    w, h = size
    r = subprocess.run([EXE, str(w), str(h)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split(None, 1) for l in r.stdout.strip().splitlines())
    assert got["page_aligned"] == "1 1"
    assert got["kind"] == "1 1"                      # page-locked: the legacy calls DMA straight from / to the frame
    assert (got["plain_kind_before"], got["plain_kind_after"], got["plain_kind_end"]) == ("0", "1", "0")
    chk, tcv = ck.best_checker(), ck.best_tcv_checker()
    yuv = lcg_bytes(F.frame_bytes(F.IMG_YUV420P, w, h), 5)
    _, rgb = chk.convert(yuv, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)
    assert got["yuv420p_rgb24"] == fnv(rgb)
    _, flipped = tcv.flip_v(rgb, w, h, 3)
    assert got["flip_v"] == fnv(flipped)
    _, y422 = chk.convert(flipped, F.IMG_RGB24, F.IMG_YUV422P, w, h, pad=0)
    assert got["rgb24_yuv422p"] == fnv(y422)
    # by now buffer 0 holds the flipped RGB frame: its first w*h*3/2 bytes are what the last conversion reads as YUV420P
    _, bgr = chk.convert(flipped[: F.frame_bytes(F.IMG_YUV420P, w, h)], F.IMG_YUV420P, F.IMG_BGR24, w, h, pad=0)
    assert got["yuv420p_bgr24"] == fnv(bgr)
    # the ring segment: four frames in buffers of their own, one acgpu_chain_frame_list_host call, in place (-I 5 -G 0.8)
    from chain_ref import DEINTERLACE, GAMMA, ref_chain
    for i in range(4):
        frame = lcg_bytes(F.frame_bytes(F.IMG_YUV420P, w, h), 10 + i)
        want, of, ow, oh = ref_chain(tcv, chk, frame, F.IMG_YUV420P, w, h, [(DEINTERLACE, 5), (GAMMA, 0.8)])
        assert (of, ow, oh) == (F.IMG_YUV420P, w, h) and got["ring_%d" % i] == fnv(want), i
