"""Longer run of the seeded differential fuzz tests (tests/test_gpu_fuzz.py) with seeds beyond the ones pytest uses:
    python tests/soak_fuzz.py [chunks]        (needs a GPU; every chunk = 60 batched conversions + 70 plane operations
                                               + 60 legacy host calls, all compared with the checker)"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import checkers as ck  # noqa: E402
import test_gpu_fuzz as t  # noqa: E402


def main():
    chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ac = t.pkg.AcGpu()
    assert ac.ac_init(t.pkg.AC_ALL) == 1, ac.last_error()
    chk = ck.best_checker()
    t0 = time.time()
    for chunk in range(100, 100 + chunks):
        t.test_fuzz_batched_conversions(ac, chk, chunk)
        t.test_fuzz_plane_operations(ac, chunk)
        t.test_fuzz_legacy_host_calls(ac, chk, chunk)
    print("soak ok: %d chunks x (60 conversions + 70 plane operations + 60 legacy calls) in %.0f s" % (chunks, time.time() - t0))


if __name__ == "__main__":
    main()
