"""Generates tests/golden/tcv_digests.json from the UNMODIFIED reference libtcvideo.

Run in the build container (needs oracle/_ref/libtcv_ref.so, i.e. /root/reference):
    python tests/golden/make_golden_tcv.py
Every case of tests/tcv_cases.py is run through tcv_deinterlace / tcv_resize / tcv_clip / tcv_reduce / tcv_flip_v /
tcv_flip_h / tcv_gamma_correct / tcv_antialias (libtcvideo/tcvideo.c) with aclib on its plain-C path; the file records
the call's return value and the first 16 hex digits of the SHA-256 of the output plane (0x55-prefilled).
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import checkers as ck  # noqa: E402
import tcv_cases  # noqa: E402


def main():
    ref = ck.TcvRef()
    out = {}
    for case in tcv_cases.cases():
        ok, d = tcv_cases.run_case(ref, case)
        out[case[0]] = [int(ok), hashlib.sha256(d.tobytes()).hexdigest()[:16] if ok else ""]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tcv_digests.json")
    with open(path, "w") as f:
        json.dump({"generator": "oracle/_ref/libtcv_ref.so (reference libtcvideo over plain-C aclib)", "prefill": 0x55,
                   "digests": out}, f, indent=0, sort_keys=True)
    print("wrote", path, len(out), "digests")


if __name__ == "__main__":
    main()
