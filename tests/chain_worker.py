"""Child process of tests/test_gpu_chain.py: $ACGPU_CHAIN_SCRATCH_BYTES is read once per process, so the sub-batched walk of
acgpu_chain_batch (a temporary smaller than the batch) is checked in a fresh interpreter.  Prints OK <sub-batches>."""
import os
import sys

import numpy as np

import __graft_entry__ as entry
import checkers as ck
import test_gpu_chain as t
from chain_ref import CONVERT, FLIP_H, FLIP_V, RESIZE
from checkers import F

pkg = entry.load_package()


def main():
    ac = pkg.AcGpu()
    assert ac.ac_init(pkg.AC_ALL) == 1, ac.last_error()
    tcv, conv = ck.best_tcv_checker(), ck.best_checker()
    w, h, nf = 640, 480, 7
    budget = int(os.environ["ACGPU_CHAIN_SCRATCH_BYTES"])
    for fmt, stages in [(F.IMG_YUV420P, [(CONVERT, F.IMG_RGB24), (FLIP_V,), (CONVERT, F.IMG_YUV422P)]),
                        (F.IMG_RGB24, [(RESIZE, -4, 3), (FLIP_V,)]),
                        # two conversions through an RGB frame nobody looks at: fused into one pass unless $ACGPU_CHAIN_FUSE=0
                        (F.IMG_YUV420P, [(CONVERT, F.IMG_RGB24), (CONVERT, F.IMG_YUV422P)]),
                        (F.IMG_YUV420P, [(CONVERT, F.IMG_BGRA32), (CONVERT, F.IMG_YUV444P)]),
                        (F.IMG_YUV420P, [(FLIP_V,), (CONVERT, F.IMG_BGR24), (CONVERT, F.IMG_YUV420P), (FLIP_H,)]),
                        (F.IMG_YUV420P, [(CONVERT, F.IMG_ARGB32), (CONVERT, F.IMG_YUY2)])]:
        frames = t.frames_of(fmt, w, h, nf, 900)
        want, _ = t.expect(tcv, conv, frames, fmt, w, h, stages)
        got, _ = t.run_device(ac, frames, fmt, w, h, stages, gap=512)
        if not np.array_equal(got, want):
            print("MISMATCH", fmt, stages)
            return 1
    print("OK", budget)
    return 0


if __name__ == "__main__":
    sys.exit(main())
