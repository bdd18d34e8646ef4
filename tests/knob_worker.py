"""Child process of tests/test_gpu_knobs.py: libacgpu reads its profiling knobs ($ACGPU_*) once per process, so each
setting is checked in a fresh interpreter.  Runs a fixed set of 4:2:0 conversions against the checker; prints OK."""
import sys

import numpy as np

import __graft_entry__ as entry
import checkers as ck

pkg = entry.load_package()
F = pkg.F


def main():
    tier = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ac = pkg.AcGpu()
    assert ac.ac_init(pkg.AC_ALL) == 1, ac.last_error()
    chk = ck.best_checker()
    ac.lib.acgpu_force_tier(tier)
    pairs = [(F.IMG_YUV420P, F.IMG_RGB24), (F.IMG_YUV420P, F.IMG_BGR24)]
    if tier == 0:
        pairs += [(F.IMG_YUV420P, F.IMG_ARGB32), (F.IMG_RGB24, F.IMG_YUV420P), (F.IMG_BGRA32, F.IMG_YV12),
                  (F.IMG_YUV420P, F.IMG_YUV422P), (F.IMG_YUV444P, F.IMG_YUV420P), (F.IMG_UYVY, F.IMG_YUV420P),
                  (F.IMG_YUV420P, F.IMG_YVYU), (F.IMG_YUV411P, F.IMG_YUV420P), (F.IMG_YUV420P, F.IMG_YUV420P),
                  # the row form of the tensor-map staged loads ($ACGPU_TMA_AUTO bits 1 and 2)
                  (F.IMG_YUV422P, F.IMG_RGB24), (F.IMG_YUV444P, F.IMG_BGR24), (F.IMG_YUV411P, F.IMG_RGB24),
                  (F.IMG_YUY2, F.IMG_RGB24), (F.IMG_UYVY, F.IMG_BGR24), (F.IMG_YVYU, F.IMG_BGR24)]
    n = 0
    for (w, h, nf) in [(1920, 16, 2), (720, 36, 3), (1280, 6, 1), (64, 4, 2), (4128, 4, 1), (4160, 6, 2)]:
        for sf, df in pairs:
            frames = np.stack([ck.random_frame(sf, w, h, seed=70 + i) for i in range(nf)])
            got = ac.convert_batch(frames, sf, df, w, h, prefill=0x33)
            for i in range(nf):
                want = chk.convert(frames[i], sf, df, w, h, prefill=0x33, pad=0)[1]
                if not np.array_equal(got[i], want):
                    print(f"MISMATCH {F.NAMES[sf]}->{F.NAMES[df]} {w}x{h} frame {i} tier {ac.lib.acgpu_last_kernel_tier()}")
                    return 1
            n += 1
    print("OK", n)
    return 0


if __name__ == "__main__":
    sys.exit(main())
