"""Seeded case list for the libtcvideo-shaped plane operations (TEST INFRASTRUCTURE).

One list, three users: tests/golden/make_golden_tcv.py records the REFERENCE's output digest for every case,
tests/test_oracle.py replays them through oracle/ac_oracle.c, and the -m gpu tests replay them through libacgpu.
A case is ``(key, op, (w, h, bpp), args, image_kind)``; ``run_case(checker, case)`` returns ``(ok, dest)``.
"""
from __future__ import annotations

import numpy as np

import checkers as ck

SIZES = [(64, 16), (48, 18), (33, 7), (176, 144), (16, 2), (40, 9)]


def blocky_image(w: int, h: int, bpp: int, seed: int) -> np.ndarray:
    """Flat blocks with a little noise: neighbouring pixels are often 'the same colour' (difference < 25), which is
    what tcv_antialias keys on (libtcvideo/tcvideo.c:917-949); random bytes almost never trigger it."""
    base = (np.add.outer(np.arange(h) // 3 * 40, np.arange(w) // 5 * 30) % 256).astype(np.int32)
    img = np.repeat(base.reshape(-1), bpp)
    if bpp == 3:
        img = img + np.tile(np.array([0, 9, 17]), w * h)
    noise = ck.splitmix_bytes(img.size, seed) % 12
    return ((img + noise) % 256).astype(np.uint8)


def image(kind: str, w: int, h: int, bpp: int, seed: int) -> np.ndarray:
    return blocky_image(w, h, bpp, seed) if kind == "blocky" else ck.splitmix_bytes(w * h * bpp, seed)


def cases():
    out = []
    seed = 100
    for bpp in (1, 3):
        for (w, h) in SIZES:
            dims = (w, h, bpp)
            tag = f"{w}x{h}x{bpp}"
            for mode in range(4):
                if mode == 1 and h < 2:
                    continue
                out.append((f"deinterlace:{tag}:{mode}", "deinterlace", dims, (mode,), "random"))
            for a in [(0, 0, 0, 0), (2, 3, 1, 2), (-2, -3, -1, -4), (5, -2, 0, 3), (w + 2, -6, 1, 1), (-4, w + 1, h + 1, -5),
                      (w, 0, 0, 0), (0, 0, h, 0)]:
                out.append((f"clip:{tag}:{a}", "clip", dims, a + (16 + bpp,), "random"))
            for a in [(1, 1), (1, 2), (2, 1), (2, 2), (3, 2), (4, 4), (0, 1), (1, -1), (w + 1, 1)]:
                out.append((f"reduce:{tag}:{a}", "reduce", dims, a, "random"))
            for inplace in (False, True):
                out.append((f"flip_v:{tag}:{int(inplace)}", "flip_v", dims, (inplace,), "random"))
                out.append((f"flip_h:{tag}:{int(inplace)}", "flip_h", dims, (inplace,), "random"))
            for g in (0.45, 1.0, 2.2, 0.0, -1.0):
                out.append((f"gamma:{tag}:{g}", "gamma", dims, (g,), "random"))
            for a in [(0.333, 0.5), (0.0, 0.0), (1.0, 1.0), (0.7, 0.1), (1.5, 0.5), (0.5, -0.1)]:
                out.append((f"antialias:{tag}:{a}:random", "antialias", dims, a, "random"))
                out.append((f"antialias:{tag}:{a}:blocky", "antialias", dims, a, "blocky"))
    for (w, h, rw, rh, sw, sh) in [(64, 48, -2, 0, 8, 8), (64, 48, 0, -1, 8, 8), (64, 48, 3, 0, 8, 8), (64, 48, 0, 2, 4, 8),
                                   (32, 16, 1, 0, 2, 1), (176, 144, 0, -6, 8, 8), (176, 144, -5, 0, 8, 8), (64, 48, 0, 5, 8, 1)]:
        for bpp in (1, 3):
            out.append((f"resize:{w}x{h}x{bpp}:{(rw, rh, sw, sh)}", "resize", (w, h, bpp), (rw, rh, sw, sh), "random"))
    return [(k, op, d, a, img, seed + i) for i, (k, op, d, a, img) in enumerate(out)]


def run_case(chk, case):
    key, op, (w, h, bpp), args, kind, seed = case
    src = image(kind, w, h, bpp, seed)
    if op == "deinterlace":
        return 1, chk.deinterlace(src, w, h, bpp, *args)
    if op == "resize":
        return 1, chk.resize(src, w, h, bpp, *args)
    if op == "clip":
        return chk.clip(src, w, h, bpp, *args[:4], black=args[4])
    if op == "reduce":
        return chk.reduce(src, w, h, bpp, *args)
    if op in ("flip_v", "flip_h"):
        return getattr(chk, op)(src, w, h, bpp, inplace=args[0])
    if op == "gamma":
        return chk.gamma(src, w, h, bpp, *args)
    if op == "antialias":
        return chk.antialias(src, w, h, bpp, *args)
    raise ValueError(op)
