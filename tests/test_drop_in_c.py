"""A C program written purely against aclib's public interface compiles against include/*.h and links against
libacgpu.so unchanged (CPU test), and produces the checker's bytes when run on a B200 (-m gpu test)."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = entry.ROOT
SRC = os.path.join(ROOT, "tests", "c", "drop_in_caller.c")
EXE = os.path.join(ROOT, "tests", "c", "drop_in_caller")
pkg = entry.load_package()


def build_caller():
    if not os.path.exists(pkg.LIB_PATH):
        entry.build()
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", EXE, SRC,
                    "-L", entry.PKG_DIR, "-lacgpu", "-Wl,-rpath," + entry.PKG_DIR, "-Wl,--no-undefined"], check=True)


def test_c_caller_compiles_and_links_against_libacgpu():
    build_caller()
    out = subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True, check=True).stdout
    used = {l.split()[-1] for l in out.splitlines() if " ac_" in l}
    assert {"ac_init", "ac_cpuinfo", "ac_flagstotext", "ac_imgconvert", "ac_average", "ac_rescale", "ac_memcpy"} <= used


def fnv(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def lcg_bytes(n: int, seed: int) -> np.ndarray:
    out = np.empty(n, np.uint8)
    z = seed
    for i in range(n):
        z = (z * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = z >> 56
    return out


@pytest.mark.gpu
def test_c_caller_output_matches_checker():
    import checkers as ck
    F = ck.F
    build_caller()
    w, h = 64, 32
    r = subprocess.run([EXE, str(w), str(h)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split(None, 1) for l in r.stdout.strip().splitlines())
    assert "cuda" in got["accel"]
    chk = ck.best_checker()
    yuv = lcg_bytes(F.frame_bytes(F.IMG_YUV420P, w, h), 1)
    _, rgb = chk.convert(yuv, F.IMG_YUV420P, F.IMG_RGB24, w, h, pad=0)
    assert got["yuv420p_rgb24"] == fnv(rgb)
    _, y422 = chk.convert(rgb, F.IMG_RGB24, F.IMG_YUV422P, w, h, pad=0)
    assert got["rgb24_yuv422p"] == fnv(y422)
    _, bgr = chk.convert(yuv, F.IMG_YV12, F.IMG_BGR24, w, h, pad=0)
    assert got["yv12_bgr24"] == fnv(bgr)
    n = w * 3
    assert got["average"] == fnv(chk.average(bgr[:n].copy(), bgr[2 * n:3 * n].copy()))
    assert got["rescale"] == fnv(chk.rescale(bgr[:n].copy(), bgr[n:2 * n].copy(), 49152, 16384))
    assert got["rescale_copy"] == fnv(bgr[:n])
    assert got["memcpy"] == fnv(bgr[1:n + 1])
