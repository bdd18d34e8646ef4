#!/usr/bin/env python
"""Regenerates SURVEY.md Appendix B: where aclib's SSE2 path differs from its own plain-C path (CPU only).

    python tests/golden/simd_divergence.py > profiles/r1_reference_simd_vs_c.md
(TEST INFRASTRUCTURE: it loads the reference builds under oracle/_ref, like the other scripts in this directory.)
Both libraries are the unmodified reference (oracle/_ref).  768x512 uniform-random bytes, dest pre-filled 0x55.
libacgpu's parity target is the C path; this table documents what a user switching from --accel sse2 will see.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import checkers as ck  # noqa: E402

F = ck.F
c, s = ck.RefLib("c"), ck.RefLib("sse2")
w, h = 768, 512
rows = []
for sf in F.FORMATS_15:
    for df in F.FORMATS_15:
        src = ck.random_frame(sf, w, h, seed=3)
        _, a = c.convert(src, sf, df, w, h, prefill=0x55, pad=0)
        _, b = s.convert(src, sf, df, w, h, prefill=0x55, pad=0)
        d = np.abs(a.astype(np.int16) - b.astype(np.int16))
        n = int((d != 0).sum())
        if n:
            rows.append((F.NAMES[sf], F.NAMES[df], n, a.size, int(d.max())))
print("# aclib SSE2 path vs aclib C path (both unmodified reference builds), 768x512 random bytes, dest pre-filled 0x55\n")
print(f"{225 - len(rows)} of 225 pairs are bit-identical; the {len(rows)} below are not (libacgpu follows the C path on all of them).\n")
print("| src | dst | differing bytes | of | max abs diff |\n|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} |")
a8 = np.repeat(np.arange(256, dtype=np.uint8), 256)
b8 = np.tile(np.arange(256, dtype=np.uint8), 256)
print("\n`ac_average`: SSE2 == C on all 65 536 byte pairs:", bool(np.array_equal(c.average(a8, b8), s.average(a8, b8))))
tot = diff = 0
for w1 in range(0, 65537, 7):
    x, y = c.rescale(a8, b8, w1, 65536 - w1), s.rescale(a8, b8, w1, 65536 - w1)
    tot += x.size
    diff += int((x != y).sum())
print(f"\n`ac_rescale`, weights w1 + w2 = 65536, every 7th w1, all byte pairs: {diff} of {tot} bytes differ (max 1).")
