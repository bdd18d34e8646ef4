"""Generates tests/golden/imgconvert_digests.json from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref/libac_ref_c.so, i.e. /root/reference):
    python tests/golden/make_golden.py
For each of the 256 reachable (src,dst) pairs it converts a seeded splitmix64 frame at two sizes with the
reference's plain-C path (ac_init(AC_NONE)) into a 0x55-prefilled, 64-byte-padded buffer and records the
first 16 hex digits of the SHA-256 of that buffer.  tests/test_oracle.py replays the same inputs through
oracle/ac_oracle.c; the -m gpu tests replay them through libacgpu.
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import checkers as ck  # noqa: E402
from checkers import F  # noqa: E402

SEED = 7
SIZES = [(64, 16), (176, 144)]


def main():
    ref = ck.RefLib("c")
    out = {}
    for sf in F.FORMATS_16:
        for df in F.FORMATS_16:
            for (w, h) in SIZES:
                src = ck.random_frame(sf, w, h, seed=SEED)
                ok, got = ref.convert(src, sf, df, w, h, prefill=0x55)
                assert ok == 1
                out[f"{F.NAMES[sf]}:{F.NAMES[df]}:{w}x{h}"] = hashlib.sha256(got.tobytes()).hexdigest()[:16]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "imgconvert_digests.json")
    with open(path, "w") as f:
        json.dump({"generator": "oracle/_ref/libac_ref_c.so (reference aclib, ac_init(AC_NONE))",
                   "seed": SEED, "prefill": 0x55, "pad": 64, "digests": out}, f, indent=0, sort_keys=True)
    print("wrote", path, len(out), "digests")


if __name__ == "__main__":
    main()
