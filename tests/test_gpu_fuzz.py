"""-m gpu: seeded differential fuzzing of acgpu_imgconvert_batch against the checker.

Random format pairs, sizes on the unit grid (small, odd-ish, ragged 4:2:0 widths, vector-friendly), 1-3 frames per
launch, planes placed at random offsets inside each frame slab (aligned for about half of the cases, so both the
vectorised and the generic tier are exercised) with canary-filled gaps between planes and frames.  Every destination
byte is compared -- the converted planes with the checker's output, everything else with the canary."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as entry
import checkers as ck

pkg = entry.load_package()
F = pkg.F
pytestmark = pytest.mark.gpu

CANARY = 0xC3


@pytest.fixture(scope="module")
def ac():
    a = pkg.AcGpu()
    assert a.ac_init(pkg.AC_ALL) == 1, a.last_error()
    return a


@pytest.fixture(scope="module")
def chk():
    return ck.best_checker()


def layout(rng, sizes, aligned):
    """Plane offsets inside a frame slab and the frame pitch."""
    offs, o = [], 0
    for s in sizes:
        o += int(rng.integers(0, 4)) * 16 if aligned else int(rng.integers(0, 40))
        offs.append(o)
        o += s
    pitch = o + (int(rng.integers(0, 4)) * 16 if aligned else int(rng.integers(0, 40)))
    if aligned:
        pitch = (pitch + 15) // 16 * 16
    return offs, max(pitch, 1)


def pick_size(rng, sf, df, aligned):
    uw, uh = F.size_unit(sf, df)
    kind = int(rng.integers(0, 4))
    if kind == 0:                                   # small
        w, h = int(rng.integers(1, 24)) * uw, int(rng.integers(1, 12)) * uh
    elif kind == 1:                                 # wide and flat (long rows, several warps per row)
        w, h = int(rng.integers(30, 520)) * uw, int(rng.integers(1, 4)) * uh
    elif kind == 2:                                 # sixteen-pixel grid
        w, h = int(rng.integers(1, 40)) * 16, int(rng.integers(1, 10)) * 2 * uh
    else:                                           # ragged: even width, width*height a multiple of 16
        w = int(rng.integers(9, 120)) * 2 * max(uw // 2, 1)
        h = int(rng.integers(1, 6)) * 8
    w -= w % uw
    h -= h % uh
    return max(w, uw), max(h, uh)


@pytest.mark.parametrize("chunk", range(8))
def test_fuzz_batched_conversions(ac, chk, chunk):
    rng = np.random.default_rng(9000 + chunk)
    fmts = F.FORMATS_15
    tiers = {1: 0, 2: 0}
    for case in range(60):
        sf, df = fmts[int(rng.integers(0, len(fmts)))], fmts[int(rng.integers(0, len(fmts)))]
        aligned = bool(rng.integers(0, 2))
        w, h = pick_size(rng, sf, df, aligned)
        nf = int(rng.integers(1, 4))
        ssz, dsz = F.plane_sizes(sf, w, h), F.plane_sizes(df, w, h)
        soff, spitch = layout(rng, ssz, aligned)
        doff, dpitch = layout(rng, dsz, aligned)
        frames = [ck.random_frame(sf, w, h, seed=int(rng.integers(0, 1 << 30))) for _ in range(nf)]
        hs = np.full(nf * spitch + 64, 0x11, np.uint8)
        for f in range(nf):
            o = 0
            for p, s in enumerate(ssz):
                hs[f * spitch + soff[p]: f * spitch + soff[p] + s] = frames[f][o:o + s]
                o += s
        dsrc = ac.malloc(hs.size).upload(hs)
        ddst = ac.malloc(nf * dpitch + 64).fill(CANARY)
        sp = (C.c_void_p * 3)(*[dsrc.ptr + soff[p] if p < len(ssz) else None for p in range(3)])
        dp = (C.c_void_p * 3)(*[ddst.ptr + doff[p] if p < len(dsz) else None for p in range(3)])
        ok = ac.lib.acgpu_imgconvert_batch(sp, sf, spitch, dp, df, dpitch, w, h, nf, None)
        ac.sync()
        what = f"chunk {chunk} case {case}: {F.NAMES[sf]}->{F.NAMES[df]} {w}x{h} nf={nf} aligned={aligned} tier={ac.lib.acgpu_last_kernel_tier()}"
        assert ok == 1, (what, ac.last_error())
        tiers[min(ac.lib.acgpu_last_kernel_tier(), 2)] += 1      # 3 = the tensor-map staged form of the vectorised tier
        got = ddst.download()
        want = np.full_like(got, CANARY)
        for f in range(nf):
            _, d = chk.convert(frames[f], sf, df, w, h, prefill=CANARY, pad=0)
            o = 0
            for p, s in enumerate(dsz):
                want[f * dpitch + doff[p]: f * dpitch + doff[p] + s] = d[o:o + s]
                o += s
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            raise AssertionError(f"{what}: {bad.size} bytes differ, first at {bad[0]} (got {got[bad[0]]}, want {want[bad[0]]})")
        assert np.array_equal(dsrc.download(), hs), what + ": source modified"
        dsrc.free()
        ddst.free()
    assert tiers[1] >= 3 and tiers[2] >= 3, tiers        # the mix must keep exercising both tiers


# ---- the frame-granular libtcvideo operations ----------------------------------------------------------------------
import tcv_cases  # noqa: E402


def _tcv_call(ac, rng, op, w, h, bpp):
    """Returns (lib function name, argument tuple after (src, dest, w, h, bpp), checker thunk, output bytes)."""
    if op == "clip":
        a = tuple(int(x) for x in (rng.integers(-9, w // 2 + 2), rng.integers(-9, w // 2 + 2), rng.integers(-5, h // 2 + 2), rng.integers(-5, h // 2 + 2)))
        black = int(rng.integers(0, 256))
        nw, nh = w - a[0] - a[1], h - a[2] - a[3]
        return "clip", a + (black,), (lambda t, s: t.clip(s, w, h, bpp, *a, black=black, prefill=CANARY)), max(nw, 0) * max(nh, 0) * bpp
    if op == "reduce":
        rw, rh = int(rng.integers(1, 9)), int(rng.integers(1, 5))      # 3 .. 6 have a kernel of their own on the 16-pixel grid
        n = w * h * bpp if (rw == 1 and rh == 1) else w * (h // rh) * bpp if rw == 1 else (w // rw) * (h // rh) * bpp
        return "reduce", (rw, rh), (lambda t, s: t.reduce(s, w, h, bpp, rw, rh, prefill=CANARY)), n
    if op == "flip_v":
        return "flip_v", (), (lambda t, s: t.flip_v(s, w, h, bpp)), w * h * bpp
    if op == "flip_h":
        return "flip_h", (), (lambda t, s: t.flip_h(s, w, h, bpp)), w * h * bpp
    if op == "gamma":
        g = 0.2 + 0.1 * int(rng.integers(0, 30))
        return "gamma_correct", (g,), (lambda t, s: t.gamma(s, w, h, bpp, g)), w * h * bpp
    if op == "deinterlace":
        mode = int(rng.integers(0, 4))
        if mode == 1 and h < 2:
            mode = 0
        return "deinterlace", (mode,), (lambda t, s: (1, t.deinterlace(s, w, h, bpp, mode))), w * (h // 2 if mode >= 2 else h) * bpp
    if op == "resize":
        horizontal = bool(rng.integers(0, 2))
        scales = [sc for sc in (1, 2, 4, 8) if (w if horizontal else h) % sc == 0]
        sc = scales[int(rng.integers(0, len(scales)))]
        size = w if horizontal else h
        lo = -(size // sc) + 1
        r = int(rng.integers(max(lo, -12), 13)) or 1
        a = (r, 0, sc, 1) if horizontal else (0, r, 1, sc)
        nw, nh = w + a[0] * a[2], h + a[1] * a[3]
        return "resize", a, (lambda t, s: (1, t.resize(s, w, h, bpp, *a))), nw * nh * bpp
    wt, bs = 0.05 * int(rng.integers(0, 21)), 0.05 * int(rng.integers(0, 21))
    return "antialias", (wt, bs), (lambda t, s: t.antialias(s, w, h, bpp, wt, bs)), w * h * bpp


@pytest.mark.parametrize("chunk", range(4))
def test_fuzz_plane_operations(ac, chunk):
    tcv = ck.best_tcv_checker()
    rng = np.random.default_rng(7700 + chunk)
    ops = ["clip", "reduce", "flip_v", "flip_h", "gamma", "antialias", "deinterlace", "resize"]
    for case in range(70):
        op = ops[int(rng.integers(0, len(ops)))]
        bpp = 1 if rng.integers(0, 2) else 3
        aligned = bool(rng.integers(0, 2))
        w = int(rng.integers(1, 20)) * 16 if aligned else int(rng.integers(1, 300))
        h = int(rng.integers(1, 40))
        nf = int(rng.integers(1, 4))
        name, args, ref, nout = _tcv_call(ac, rng, op, w, h, bpp)
        sfb = w * h * bpp
        s_off, d_off = (0, 0) if aligned else (int(rng.integers(0, 16)), int(rng.integers(0, 16)))
        spitch = sfb + (int(rng.integers(0, 3)) * 16 if aligned else int(rng.integers(0, 30)))
        dpitch = nout + (int(rng.integers(0, 3)) * 16 if aligned else int(rng.integers(0, 30)))
        if aligned:
            spitch, dpitch = (spitch + 15) // 16 * 16, (dpitch + 15) // 16 * 16
        frames = [tcv_cases.image("blocky" if op == "antialias" and rng.integers(0, 2) else "random", w, h, bpp, int(rng.integers(0, 1 << 30)))
                  for _ in range(nf)]
        hs = np.full(s_off + nf * spitch + 64, 0x22, np.uint8)
        for f in range(nf):
            hs[s_off + f * spitch: s_off + f * spitch + sfb] = frames[f]
        dsrc = ac.malloc(hs.size).upload(hs)
        ddst = ac.malloc(d_off + nf * dpitch + 64).fill(CANARY)
        fn = getattr(ac.lib, f"acgpu_{name}_batch")
        ok = fn(dsrc.ptr + s_off, ddst.ptr + d_off, w, h, bpp, *args, spitch, dpitch, nf, None)
        ac.sync()
        what = f"chunk {chunk} case {case}: {name}{args} {w}x{h}x{bpp} nf={nf} aligned={aligned}"
        got = ddst.download()
        want = np.full_like(got, CANARY)
        ok_ref = 1
        for f in range(nf):
            ok_ref, d = ref(tcv, frames[f])
            if ok_ref:
                want[d_off + f * dpitch: d_off + f * dpitch + nout] = d[:nout]
        assert ok == ok_ref, (what, ac.last_error())
        if ok and not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            raise AssertionError(f"{what}: {bad.size} bytes differ, first at {bad[0]} (got {got[bad[0]]}, want {want[bad[0]]})")
        assert np.array_equal(dsrc.download(), hs), what + ": source modified"
        dsrc.free()
        ddst.free()


@pytest.mark.parametrize("chunk", range(3))
def test_fuzz_legacy_host_calls(ac, chk, chunk):
    """The unmodified ac_imgconvert(src planes, fmt, dest planes, fmt, w, h) call on pageable host memory: random pairs
    and sizes, planes at arbitrary (unaligned) host addresses, canaries around every destination plane."""
    rng = np.random.default_rng(4400 + chunk)
    fmts = F.FORMATS_15
    for case in range(60):
        sf, df = fmts[int(rng.integers(0, len(fmts)))], fmts[int(rng.integers(0, len(fmts)))]
        w, h = pick_size(rng, sf, df, bool(rng.integers(0, 2)))
        ssz, dsz = F.plane_sizes(sf, w, h), F.plane_sizes(df, w, h)
        soff, stot = layout(rng, ssz, False)
        doff, dtot = layout(rng, dsz, False)
        frame = ck.random_frame(sf, w, h, seed=int(rng.integers(0, 1 << 30)))
        hs = np.full(stot + 64, 0x11, np.uint8)
        o = 0
        for p, s in enumerate(ssz):
            hs[soff[p]: soff[p] + s] = frame[o:o + s]
            o += s
        keep = hs.copy()
        hd = np.full(dtot + 64, CANARY, np.uint8)
        sp = (C.c_void_p * 3)(*[hs.ctypes.data + soff[p] if p < len(ssz) else None for p in range(3)])
        dp = (C.c_void_p * 3)(*[hd.ctypes.data + doff[p] if p < len(dsz) else None for p in range(3)])
        ok = ac.lib.ac_imgconvert(sp, sf, dp, df, w, h)
        what = f"chunk {chunk} case {case}: {F.NAMES[sf]}->{F.NAMES[df]} {w}x{h}"
        assert ok == 1, (what, ac.last_error())
        _, d = chk.convert(frame, sf, df, w, h, prefill=CANARY, pad=0)
        want = np.full_like(hd, CANARY)
        o = 0
        for p, s in enumerate(dsz):
            want[doff[p]: doff[p] + s] = d[o:o + s]
            o += s
        assert np.array_equal(hd, want), what
        assert np.array_equal(hs, keep), what + ": source modified"
