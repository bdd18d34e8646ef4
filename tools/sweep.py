#!/usr/bin/env python
"""Device-resident throughput sweep over format pairs (profiling aid, not the contract bench).

    python tools/sweep.py [--size 1920x1080] [--pairs cfg2|all|a:b,c:d] [--out gpurun_out/sweep.md]
Each pair: one acgpu_imgconvert_batch launch per step over a batch sized to ~1.5 GB of traffic, CUDA-event timed.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
F = pkg.F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--pairs", default="cfg2")
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tier", type=int, default=0)
    ap.add_argument("--content", default="random", choices=["random", "smooth"],
                    help="random: uniform bytes (worst case for the chroma-table bank conflicts); smooth: a soft gradient "
                         "with +-2 noise, closer to a natural picture (neighbouring lanes hit the same table words)")
    args = ap.parse_args()
    w, h = map(int, args.size.split("x"))
    if args.pairs == "cfg2":
        S = [F.IMG_YUV420P, F.IMG_YUV422P, F.IMG_YUV444P, F.IMG_YUY2, F.IMG_UYVY, F.IMG_Y8]
        D = [F.IMG_RGB24, F.IMG_BGR24, F.IMG_RGBA32]
        pairs = [(s, d) for s in S for d in D] + [(d, s) for s in S for d in D]
    elif args.pairs == "all":
        pairs = [(s, d) for s in F.FORMATS_15 for d in F.FORMATS_15]
    else:
        pairs = [tuple(F.BY_NAME[x] for x in p.split(":")) for p in args.pairs.split(",")]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ac = pkg.AcGpu()
    assert ac.ac_init(pkg.AC_CUDA) == 1, ac.last_error()
    ac.lib.acgpu_force_tier(args.tier)
    lib = ac.lib
    stream = lib.acgpu_stream_create()
    e0, e1 = lib.acgpu_event_create(), lib.acgpu_event_create()
    rows = []
    import numpy as np
    rng = np.random.default_rng(1)
    for sf, df in pairs:
        sfb, dfb = F.frame_bytes(sf, w, h), F.frame_bytes(df, w, h)
        ab = F.algorithmic_bytes(sf, df, w, h)
        batch = max(8, int(1.5e9 // ab))
        src, dst = ac.malloc(batch * sfb), ac.malloc(batch * dfb)
        if args.content == "random":
            host = rng.integers(0, 256, size=sfb, dtype=np.uint8)
        else:
            i = np.arange(sfb, dtype=np.int64)
            host = ((i // 61) % 200 + 28 + rng.integers(-2, 3, size=sfb)).astype(np.uint8)
        for i in range(batch):
            lib.acgpu_memcpy_h2d(src.ptr + i * sfb, host.ctypes.data, sfb, None)
        ac.sync()
        for _ in range(3):
            ac._ok(ac.imgconvert_batch(src.ptr, sf, sfb, dst.ptr, df, dfb, w, h, batch, stream))
        lib.acgpu_event_record(e0, stream)
        for _ in range(args.steps):
            ac._ok(ac.imgconvert_batch(src.ptr, sf, sfb, dst.ptr, df, dfb, w, h, batch, stream))
        lib.acgpu_event_record(e1, stream)
        ac.sync(stream)
        ms = lib.acgpu_event_elapsed_ms(e0, e1) / args.steps
        gbs = batch * ab / ms / 1e6
        rows.append((F.NAMES[sf], F.NAMES[df], batch * 1000.0 / ms, gbs, gbs / peak, lib.acgpu_last_kernel_tier()))
        print(f"{F.NAMES[sf]:8s} -> {F.NAMES[df]:8s} {batch*1000.0/ms:12.0f} frames/s {gbs:8.1f} GB/s {gbs/peak:6.3f} tier {rows[-1][5]}", flush=True)
        src.free(); dst.free()
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"| src | dst | frames/s @ {w}x{h} | algorithmic GB/s | frac of measured HBM peak ({peak:.0f} GB/s) | tier |\n|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r[0]} | {r[1]} | {r[2]:.0f} | {r[3]:.1f} | {r[4]:.3f} | {r[5]} |\n")


if __name__ == "__main__":
    main()
