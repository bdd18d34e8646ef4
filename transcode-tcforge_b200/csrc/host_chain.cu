// host_chain.cu -- resident multi-operation chains (include/acgpu.h, "frame chains").
//
// transcode applies several operations to every frame: do_process_frame (src/video_trans.c:192-426) runs clip ->
// deinterlace -> resize -> clip -> reduce -> flip -> mirror -> -k -> -K -> gamma -> antialias, each through libtcvideo and
// so through ac_imgconvert / ac_rescale / ac_average, and the filter wrappers round-trip yuv -> rgb -> filter -> yuv
// (filter/filter_ascii.c:367-373).  On the CPU every stage is one more pass over host memory.  Here a frame crosses PCIe
// once in each direction and all stages run on the device copy: one upload, N kernels, one download.
//
// A chain acts on whole frames in transcode's frame layouts (YUV_INIT_PLANES, tightly packed): the stages other than
// CONVERT accept the three layouts do_process_frame knows -- YUV420P, YUV422P (three planes, Bpp 1, chroma planes
// divided as video_trans.c:85-103) and RGB24 (one plane, Bpp 3) -- and treat the planes exactly as the reference's
// PROCESS_FRAME macro and its per-stage special cases do.  Host side only; every stage is one of the existing launches.
#include "host_ctx.h"

#include <math.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace acgpu {
namespace {

struct Geo {               // a frame layout between two stages
    int fmt, w, h;
};

struct PlaneSet {          // how do_process_frame sees a layout (set_vtd, video_trans.c:71-118)
    int n, Bpp;
    int wd[3], hd[3];
    uint8_t black[3];
};

bool plane_set(int fmt, PlaneSet *ps)
{
    PlaneSet s{};
    s.n = 1; s.Bpp = 1;
    for (int i = 0; i < 3; i++) { s.wd[i] = s.hd[i] = 1; s.black[i] = 0; }
    if (fmt == IMG_YUV420P || fmt == IMG_YUV422P) {
        s.n = 3;
        s.wd[1] = s.wd[2] = 2;
        s.hd[1] = s.hd[2] = fmt == IMG_YUV420P ? 2 : 1;
        s.black[1] = s.black[2] = 128;
    } else if (fmt == IMG_RGB24) {
        s.Bpp = 3;
    } else if (fmt != IMG_Y8 && fmt != IMG_GRAY8) {     // one 8-bit plane: set_vtd's defaults (video_trans.c:77-84)
        return false;
    }
    *ps = s;
    return true;
}

size_t geo_bytes(const Geo &g) { return frame_bytes(g.fmt == IMG_YV12 ? IMG_YUV420P : g.fmt, g.w, g.h); }

// byte offset of plane i inside a tightly packed frame of the layout (video_trans.c:104-117)
size_t plane_off(const PlaneSet &ps, const Geo &g, int i)
{
    size_t off = 0;
    for (int k = 0; k < i; k++) off += (size_t)(g.w / ps.wd[k]) * (size_t)(g.h / ps.hd[k]) * ps.Bpp;
    return off;
}

const char *kind_name(int k)
{
    static const char *const n[] = {"?", "convert", "clip", "deinterlace", "resize", "reduce", "flip_v", "flip_h", "rgbswap",
                                    "decolor", "gamma", "antialias"};
    return k >= 1 && k <= ACGPU_CHAIN_ANTIALIAS ? n[k] : "?";
}

// Geometry after one stage; false (error set) when the stage cannot apply to the layout.
bool stage_output(const acgpu_chain_op &op, const Geo &in, Geo *out)
{
    *out = in;
    if (op.kind == ACGPU_CHAIN_CONVERT) {
        const int df = op.p[0] == IMG_YV12 ? IMG_YUV420P : op.p[0];
        if (describe(df).kind == K_NONE) { set_error("chain: convert to unknown format 0x%x", op.p[0]); return false; }
        out->fmt = op.p[0];
        return true;
    }
    PlaneSet ps;
    if (!plane_set(in.fmt, &ps)) {
        set_error("chain: %s needs a YUV420P, YUV422P, RGB24 or single-plane (Y8 / GRAY8) frame (src/video_trans.c:77-103), got 0x%x",
                  kind_name(op.kind), in.fmt);
        return false;
    }
    if (ps.n == 1 && ps.Bpp == 1 && (op.kind == ACGPU_CHAIN_RGBSWAP || op.kind == ACGPU_CHAIN_DECOLOR)) {
        set_error("chain: %s needs colour planes", kind_name(op.kind));     // video_trans.c:359-365, 379-385 index planes 1 and 2
        return false;
    }
    const int wd = ps.wd[ps.n - 1], hd = ps.hd[ps.n - 1];
    switch (op.kind) {
    case ACGPU_CHAIN_CLIP: {
        // video_trans.c:213-223: each plane is clipped by the frame's amounts divided by its divisors; the plane slots of the
        // result are laid out for the new frame size, so the amounts must divide or planes would spill into each other
        for (int k = 0; k < 4; k++)
            if (op.p[k] % (k < 2 ? wd : hd)) { set_error("chain: clip amounts must be multiples of the chroma subsampling"); return false; }
        out->w = in.w - op.p[0] - op.p[1];
        out->h = in.h - op.p[2] - op.p[3];
        if (out->w <= 0 || out->h <= 0) { set_error("chain: clip leaves no frame"); return false; }
        return true;
    }
    case ACGPU_CHAIN_DEINTERLACE:
        // transcode's -I modes (video_trans.c:227-277): 1 interpolate, 5 linear blend (first plane only), 4 drop field
        // (every plane, half height); 2 is left to the encoder; 3 needs tcv_zoom, which is outside this library
        if (op.p[0] == 1 || op.p[0] == 5 || op.p[0] == 2) return true;
        if (op.p[0] == 4) {
            if ((in.h / 2) % hd) { set_error("chain: drop-field needs height/2 divisible by the chroma subsampling"); return false; }
            out->h = in.h / 2;
            return true;
        }
        set_error("chain: deinterlace mode %d not available (3 needs tcv_zoom)", op.p[0]);
        return false;
    case ACGPU_CHAIN_RESIZE:
        // video_trans.c:281-297: units of 8 pixels, rows first
        out->w = in.w + op.p[0] * 8;
        out->h = in.h + op.p[1] * 8;
        if (out->w <= 0 || out->h <= 0) { set_error("chain: resize leaves no frame"); return false; }
        return true;
    case ACGPU_CHAIN_REDUCE:
        if (op.p[0] <= 0 || op.p[1] <= 0) { set_error("chain: invalid reduce factors"); return false; }
        out->w = in.w / op.p[0];
        out->h = in.h / op.p[1];
        if (out->w <= 0 || out->h <= 0 || out->w % wd || out->h % hd) { set_error("chain: reduce leaves no usable frame"); return false; }
        return true;
    case ACGPU_CHAIN_FLIP_V: case ACGPU_CHAIN_FLIP_H: case ACGPU_CHAIN_RGBSWAP: case ACGPU_CHAIN_DECOLOR:
    case ACGPU_CHAIN_GAMMA: case ACGPU_CHAIN_ANTIALIAS:
        return true;
    default:
        set_error("chain: unknown stage kind %d", op.kind);
        return false;
    }
}

// Stages that leave the frame as it is: -I 2 (left to the encoder, video_trans.c:278), a conversion to the same format
// (tcvideo.c:1037-1042 with src == dest), a resize by nothing.
bool stage_noop(const acgpu_chain_op &op, const Geo &in)
{
    switch (op.kind) {
    case ACGPU_CHAIN_DEINTERLACE: return op.p[0] == 2;
    case ACGPU_CHAIN_CONVERT: return op.p[0] == in.fmt;
    case ACGPU_CHAIN_RESIZE: return op.p[0] == 0 && op.p[1] == 0;
    default: return false;
    }
}

// Does the stage rewrite its frame where it lies (true) or produce a new frame (false)?
bool stage_in_place(const acgpu_chain_op &op)
{
    return op.kind == ACGPU_CHAIN_RGBSWAP || op.kind == ACGPU_CHAIN_DECOLOR || op.kind == ACGPU_CHAIN_GAMMA;
}

struct Plan {
    std::vector<Geo> geo;          // geo[k] = layout entering stage k; geo[nops] = result
    size_t max_bytes = 0;          // largest frame anywhere along the chain
    int    n_out = 0;              // stages that produce a new frame
    int    n_live = 0;             // stages that do anything
    bool   two_pass = false;       // a resize of both dimensions (it needs a buffer between its passes)
};

bool make_plan(int fmt, int w, int h, const acgpu_chain_op *ops, int nops, Plan *pl)
{
    if (describe(fmt == IMG_YV12 ? IMG_YUV420P : fmt).kind == K_NONE || w <= 0 || h <= 0) { set_error("chain: invalid source frame"); return false; }
    if (nops < 0 || (nops && !ops)) { set_error("chain: null stage list"); return false; }
    pl->geo.assign(1, Geo{fmt, w, h});
    pl->max_bytes = frame_bytes(fmt == IMG_YV12 ? IMG_YUV420P : fmt, w, h);
    for (int k = 0; k < nops; k++) {
        Geo o;
        if (!stage_output(ops[k], pl->geo.back(), &o)) return false;
        if (!stage_noop(ops[k], pl->geo.back())) {
            pl->n_live++;
            if (!stage_in_place(ops[k])) pl->n_out++;
            if (ops[k].kind == ACGPU_CHAIN_RESIZE && ops[k].p[0] && ops[k].p[1]) {
                pl->two_pass = true;
                const Geo mid{o.fmt, pl->geo.back().w, o.h};      // after the row pass, before the column pass
                if (geo_bytes(mid) > pl->max_bytes) pl->max_bytes = geo_bytes(mid);
            }
        }
        if (geo_bytes(o) > pl->max_bytes) pl->max_bytes = geo_bytes(o);
        pl->geo.push_back(o);
    }
    return true;
}

struct Buf { uint8_t *p; size_t pitch; };

struct DeviceOnly {      // pointers handed to the stage launchers below are device memory: skip the per-call classification
    DeviceOnly() { tls.device_only++; }
    ~DeviceOnly() { tls.device_only--; }
};

// One out-of-place stage, in -> out, on `nf` frames.
bool run_out_stage(DevCtx *c, acgpu_stream_t as, const acgpu_chain_op &op, const Geo &gi, const Geo &go, Buf in, Buf out, int nf)
{
    cudaStream_t st = pick_stream(c, as);
    if (op.kind == ACGPU_CHAIN_CONVERT) {
        const int sf = gi.fmt == IMG_YV12 ? IMG_YUV420P : gi.fmt, df = go.fmt == IMG_YV12 ? IMG_YUV420P : go.fmt;
        uint8_t *sp[3], *dp[3];
        sp[0] = in.p;  sp[1] = in.p + (size_t)gi.w * gi.h;  sp[2] = sp[1] + chroma_plane_bytes(sf, gi.w, gi.h);
        dp[0] = out.p; dp[1] = out.p + (size_t)go.w * go.h; dp[2] = dp[1] + chroma_plane_bytes(df, go.w, go.h);
        return acgpu_imgconvert_batch(sp, (ImageFormat)gi.fmt, in.pitch, dp, (ImageFormat)go.fmt, out.pitch, gi.w, gi.h, nf, as) == 1;
    }
    PlaneSet ps;
    if (!plane_set(gi.fmt, &ps)) return false;
    auto copy_plane = [&](int i) {
        const size_t n = (size_t)(gi.w / ps.wd[i]) * (size_t)(gi.h / ps.hd[i]) * ps.Bpp;
        return check(cudaMemcpy2DAsync(out.p + plane_off(ps, go, i), out.pitch, in.p + plane_off(ps, gi, i), in.pitch, n, (size_t)nf,
                                       cudaMemcpyDeviceToDevice, st), "chain plane copy");
    };
    bool ok = true;
    for (int i = 0; i < ps.n && ok; i++) {
        const uint8_t *s = in.p + plane_off(ps, gi, i);
        uint8_t *d = out.p + plane_off(ps, go, i);
        const int pw = gi.w / ps.wd[i], ph = gi.h / ps.hd[i];
        switch (op.kind) {
        case ACGPU_CHAIN_CLIP:
            ok = acgpu_clip_batch(s, d, pw, ph, ps.Bpp, op.p[0] / ps.wd[i], op.p[1] / ps.wd[i], op.p[2] / ps.hd[i], op.p[3] / ps.hd[i],
                                  ps.black[i], in.pitch, out.pitch, nf, as) == 1;
            break;
        case ACGPU_CHAIN_DEINTERLACE:
            if (op.p[0] == 4)        // PROCESS_FRAME(tcv_deinterlace, DROP_FIELD_BOTTOM), video_trans.c:255-259
                ok = acgpu_deinterlace_batch(s, d, pw, ph, ps.Bpp, ACGPU_DEINT_DROP_FIELD_BOTTOM, in.pitch, out.pitch, nf, as) == 1;
            else if (i == 0)         // first plane only, the others are copied (:238-249, :267-277)
                ok = acgpu_deinterlace_batch(s, d, pw, ph, ps.Bpp, op.p[0] == 1 ? ACGPU_DEINT_INTERPOLATE : ACGPU_DEINT_LINEAR_BLEND,
                                             in.pitch, out.pitch, nf, as) == 1;
            else
                ok = copy_plane(i);
            break;
        case ACGPU_CHAIN_REDUCE:
            ok = acgpu_reduce_batch(s, d, pw, ph, ps.Bpp, op.p[0], op.p[1], in.pitch, out.pitch, nf, as) == 1;
            break;
        case ACGPU_CHAIN_FLIP_V:
            ok = acgpu_flip_v_batch(s, d, pw, ph, ps.Bpp, in.pitch, out.pitch, nf, as) == 1;
            break;
        case ACGPU_CHAIN_FLIP_H:
            ok = acgpu_flip_h_batch(s, d, pw, ph, ps.Bpp, in.pitch, out.pitch, nf, as) == 1;
            break;
        case ACGPU_CHAIN_ANTIALIAS:      // only the first plane is smoothed (video_trans.c:409-421)
            ok = i == 0 ? acgpu_antialias_batch(s, d, pw, ph, ps.Bpp, op.d[0], op.d[1], in.pitch, out.pitch, nf, as) == 1 : copy_plane(i);
            break;
        default:
            set_error("chain: stage %s has no out-of-place form", kind_name(op.kind));
            ok = false;
        }
    }
    return ok;
}

// -B / -X (video_trans.c:281-297): rows first, then columns, each over every plane; `mid` holds the frame between the two
// passes (unused when only one dimension changes).
bool run_resize(acgpu_stream_t as, const acgpu_chain_op &op, const Geo &gi, const Geo &go, Buf in, Buf mid, Buf out, int nf)
{
    PlaneSet ps;
    if (!plane_set(gi.fmt, &ps)) return false;
    Geo gm = gi;
    gm.h = go.h;
    const bool two = op.p[0] && op.p[1];
    Buf cur = in;
    Geo gc = gi;
    if (op.p[1]) {
        Buf to = two ? mid : out;
        for (int i = 0; i < ps.n; i++)
            if (acgpu_resize_batch(cur.p + plane_off(ps, gc, i), to.p + plane_off(ps, gm, i), gc.w / ps.wd[i], gc.h / ps.hd[i], ps.Bpp,
                                   0, op.p[1], 8 / ps.wd[i], 8 / ps.hd[i], cur.pitch, to.pitch, nf, as) != 1) return false;
        cur = to;
        gc = gm;
    }
    if (op.p[0])
        for (int i = 0; i < ps.n; i++)
            if (acgpu_resize_batch(cur.p + plane_off(ps, gc, i), out.p + plane_off(ps, go, i), gc.w / ps.wd[i], gc.h / ps.hd[i], ps.Bpp,
                                   op.p[0], 0, 8 / ps.wd[i], 8 / ps.hd[i], cur.pitch, out.pitch, nf, as) != 1) return false;
    return true;
}

// One in-place stage on `nf` frames at `b`; `tmp` is a scratch frame buffer of at least the same size.
bool run_in_place_stage(DevCtx *c, acgpu_stream_t as, const acgpu_chain_op &op, const Geo &g, Buf b, Buf tmp, int nf)
{
    cudaStream_t st = pick_stream(c, as);
    PlaneSet ps;
    if (!plane_set(g.fmt, &ps)) return false;
    const size_t uv = ps.n == 3 ? (size_t)(g.w / ps.wd[1]) * (size_t)(g.h / ps.hd[1]) : 0;
    uint8_t *u = b.p + plane_off(ps, g, 1), *v = b.p + plane_off(ps, g, ps.n - 1);
    switch (op.kind) {
    case ACGPU_CHAIN_RGBSWAP:
        if (g.fmt == IMG_RGB24) {       // video_trans.c:352-358: R <-> B of every pixel = RGB24 -> BGR24 on itself
            uint8_t *pl[3] = {b.p, nullptr, nullptr};
            return acgpu_imgconvert_batch(pl, IMG_RGB24, b.pitch, pl, IMG_BGR24, b.pitch, g.w, g.h, nf, as) == 1;
        }
        // :359-365 U and V planes change places
        return check(cudaMemcpy2DAsync(tmp.p, tmp.pitch, u, b.pitch, uv, (size_t)nf, cudaMemcpyDeviceToDevice, st), "chain uv swap")
            && check(cudaMemcpy2DAsync(u, b.pitch, v, b.pitch, uv, (size_t)nf, cudaMemcpyDeviceToDevice, st), "chain uv swap")
            && check(cudaMemcpy2DAsync(v, b.pitch, tmp.p, tmp.pitch, uv, (size_t)nf, cudaMemcpyDeviceToDevice, st), "chain uv swap");
    case ACGPU_CHAIN_DECOLOR:
        if (g.fmt == IMG_RGB24) return acgpu_decolor_rgb24_batch(b.p, g.w, g.h, b.pitch, nf, as) == 1;     // :371-378
        return check(cudaMemset2DAsync(u, b.pitch, 128, 2 * uv, (size_t)nf, st), "chain decolor");           // :379-385 (U and V are adjacent)
    case ACGPU_CHAIN_GAMMA:             // :391-396 first plane only
        return acgpu_gamma_correct_batch(b.p, b.p, g.w, g.h, ps.Bpp, op.d[0], b.pitch, b.pitch, nf, as) == 1;
    default:
        set_error("chain: stage %s has no in-place form", kind_name(op.kind));
        return false;
    }
}

// Runs the whole chain on `nf` device-resident frames.  `in` is read (and, when `in_mutable`, may be rewritten); s0 / s1 are
// two scratch frame buffers (pitch >= plan.max_bytes); when `out.p` is set the result is delivered there, otherwise it is
// left wherever the last stage put it.  Returns the location of the result in *res.
bool run_chain(DevCtx *c, acgpu_stream_t as, const Plan &pl, const acgpu_chain_op *ops, int nops, Buf in, bool in_mutable,
               Buf s0, Buf s1, Buf out, int nf, Buf *res)
{
    DeviceOnly guard;
    cudaStream_t st = pick_stream(c, as);
    Buf cur = in;
    bool cur_mutable = in_mutable;
    int outs_left = pl.n_out;
    auto next_scratch = [&]() { return cur.p == s0.p ? s1 : s0; };       // the scratch buffer the current frame is not in
    static const bool fuse = [] { const char *e = getenv("ACGPU_CHAIN_FUSE"); return !e || atoi(e) != 0; }();
    auto copy_to = [&](Buf to, const Geo &g) {
        return check(cudaMemcpy2DAsync(to.p, to.pitch, cur.p, cur.pitch, geo_bytes(g), (size_t)nf, cudaMemcpyDeviceToDevice, st), "chain copy");
    };
    for (int k = 0; k < nops; k++) {
        const Geo &gi = pl.geo[(size_t)k], &go = pl.geo[(size_t)k + 1];
        if (stage_noop(ops[k], gi)) continue;
        if (stage_in_place(ops[k])) {
            if (!cur_mutable) {          // the caller's source is read-only: continue on a copy (straight in `out` if nothing moves later)
                Buf to = (outs_left == 0 && out.p) ? out : next_scratch();
                if (!copy_to(to, gi)) return false;
                cur = to;
                cur_mutable = true;
            }
            Buf tmp = cur.p == s0.p ? s1 : s0;
            if (!run_in_place_stage(c, as, ops[k], gi, cur, tmp, nf)) return false;
            continue;
        }
        // Two conversions through an RGB frame nobody looks at (YUV420P -> RGB -> planar YUV: BASELINE config 4, the filter
        // wrappers' round trip with nothing in between) are one fused pass when the geometry allows it.
        if (fuse && k + 1 < nops && ops[k].kind == ACGPU_CHAIN_CONVERT && ops[k + 1].kind == ACGPU_CHAIN_CONVERT && gi.fmt == IMG_YUV420P
            && describe(ops[k].p[0]).kind == K_RGB && !stage_noop(ops[k + 1], go)) {
            const Geo &g2 = pl.geo[(size_t)k + 2];
            Buf to2 = (outs_left - 2 == 0 && out.p) ? out : next_scratch();
            ConvertArgs a{};
            a.srcfmt = IMG_YUV420P; a.dstfmt = g2.fmt; a.w = gi.w; a.h = gi.h; a.nframes = nf; a.stream = st;
            a.src.p[0] = cur.p; a.src.p[1] = cur.p + (size_t)gi.w * gi.h; a.src.p[2] = a.src.p[1] + chroma_plane_bytes(IMG_YUV420P, gi.w, gi.h);
            a.src.pitch = cur.pitch;
            a.dst.p[0] = to2.p; a.dst.p[1] = to2.p + (size_t)g2.w * g2.h; a.dst.p[2] = a.dst.p[1] + chroma_plane_bytes(g2.fmt, g2.w, g2.h);
            a.dst.pitch = to2.pitch;
            tls.err[0] = 0;
            // grid.y carries the frame index: longer sub-batches are cut into launches of 32768 frames, like every batch call
            bool fused = true;
            for (int f0 = 0; f0 < nf && fused; f0 += 32768) {
                ConvertArgs b = a;
                b.nframes = nf - f0 < 32768 ? nf - f0 : 32768;
                for (int pl_ = 0; pl_ < 3; pl_++) { b.src.p[pl_] += (size_t)f0 * a.src.pitch; b.dst.p[pl_] += (size_t)f0 * a.dst.pitch; }
                fused = convert_fused_yuv420_rgb_yuv(b);
                if (!fused && f0 > 0) return false;       // cannot happen (same geometry as the first part); never mix the two forms
            }
            if (fused) {
                tls.last_tier = 2;
                outs_left -= 2;
                cur = to2;
                cur_mutable = true;
                k++;                      // both stages done
                continue;
            }
            if (tls.err[0]) return false;     // a launch failed (outside the fused domain it returns false silently)
        }
        outs_left--;
        const bool to_out = outs_left == 0 && out.p;
        Buf to = to_out ? out : next_scratch();
        if (ops[k].kind == ACGPU_CHAIN_RESIZE) {
            // the frame between the two passes goes to a scratch buffer that is neither the input nor the result; when
            // both scratch buffers are taken (input in one, result due in the other) the passes go input -> other -> input's
            // buffer, which is free again once the first pass has read it
            Buf mid = to, res = to;
            if (to_out) {
                mid = cur.p == s0.p ? s1 : s0;
            } else {
                Buf other = to.p == s0.p ? s1 : s0;
                if (other.p != cur.p) mid = other;
                else if (ops[k].p[0] && ops[k].p[1]) res = cur;      // cur is a scratch buffer here: ours to overwrite
            }
            if (!run_resize(as, ops[k], gi, go, cur, mid, res, nf)) return false;
            to = res;
        } else if (!run_out_stage(c, as, ops[k], gi, go, cur, to, nf)) {
            return false;
        }
        cur = to;
        cur_mutable = true;
    }
    if (out.p && cur.p != out.p) {       // no stage moved the frame (empty chain, or in-place stages on a mutable source)
        if (!copy_to(out, pl.geo.back())) return false;
        cur = out;
    }
    *res = cur;
    return true;
}

// ---- per-device host threads for the *_multi calls ---------------------------------------------------------------------
struct DeviceWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool pending = false, quit = false;

    DeviceWorker()
    {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return pending || quit; });
                if (quit) return;
                lk.unlock();
                job();
                lk.lock();
                pending = false;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> f)
    {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f);
        pending = true;
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return !pending; });
    }
    ~DeviceWorker()
    {
        { std::lock_guard<std::mutex> lk(m); quit = true; cv.notify_all(); }
        if (th.joinable()) th.join();
    }
};

std::mutex g_multi_mutex;                                     // one multi-device call at a time
// Raw pointers, never deleted by a static destructor: joining a worker at process exit would run its thread-local CUDA
// cleanup after the statically linked runtime may be gone.  acgpu_shutdown() tears them down while CUDA is alive.
DeviceWorker *g_workers[kMaxDev];

}  // namespace

// Frames are independent: a run of host frames is shared out among the devices, each served by its own
// long-lived host thread (created on first use, parked on a condition variable: streams, pipeline buffers and device
// contexts persist from call to call).  No exchange between devices, no collective.  The devices of one box do not all get
// the same share of the host's DMA bandwidth (profiles/r2_host_dma_probe_8gpu.md: four GPUs at half the rate of the other
// four when all eight copy), so the run is not cut into equal blocks: every device thread takes the next grain of frames
// from a shared counter until none are left.  job(device, first_frame, end_frame) runs on the device's thread with that
// device selected.
int run_on_devices(const char *who, int ndevices, int nframes, const std::function<bool(int, int, int)> &job)
{
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) { cudaGetLastError(); set_error("%s: no CUDA device", who); return 0; }
    if (ndevices <= 0) ndevices = visible;
    if (ndevices > visible || ndevices > kMaxDev) { set_error("%s: %d devices requested, %d visible", who, ndevices, visible); return 0; }
    if (nframes <= 0) return 1;
    if (ndevices > nframes) ndevices = nframes;
    // a grain is one pipelined call: long enough to amortise the pipeline's fill and drain, short enough to balance
    int grain = nframes / (ndevices * 6);
    if (grain < 16) grain = 16;
    if (grain > (nframes + ndevices - 1) / ndevices) grain = (nframes + ndevices - 1) / ndevices;
    if (ndevices == 1) grain = nframes;        // nothing to balance: one pipelined call
    std::lock_guard<std::mutex> call(g_multi_mutex);
    std::vector<int> ok((size_t)ndevices, 0);
    std::vector<std::string> why((size_t)ndevices);
    std::atomic<int> next{0};
    for (int d = 0; d < ndevices; d++) {
        if (!g_workers[d]) g_workers[d] = new DeviceWorker();
        g_workers[d]->submit([=, &ok, &why, &job, &next] {
            bool good = acgpu_set_device(d) == 1;
            for (int f0; good && (f0 = next.fetch_add(grain)) < nframes;) good = job(d, f0, f0 + grain < nframes ? f0 + grain : nframes);
            ok[(size_t)d] = good;
            if (!good) why[(size_t)d] = tls.err;
        });
    }
    for (int d = 0; d < ndevices; d++) g_workers[d]->wait();
    for (int d = 0; d < ndevices; d++)
        if (!ok[(size_t)d]) { set_error("device %d: %s", d, why[(size_t)d].c_str()); return 0; }
    return 1;
}

void stop_device_workers()
{
    std::lock_guard<std::mutex> call(g_multi_mutex);
    for (int d = 0; d < kMaxDev; d++) {
        delete g_workers[d];
        g_workers[d] = nullptr;
    }
}

}  // namespace acgpu

using namespace acgpu;

extern "C" {

int acgpu_chain_output(ImageFormat fmt, int width, int height, const acgpu_chain_op *ops, int nops,
                       ImageFormat *out_fmt, int *out_width, int *out_height)
{
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    if (out_fmt) *out_fmt = (ImageFormat)pl.geo.back().fmt;
    if (out_width) *out_width = pl.geo.back().w;
    if (out_height) *out_height = pl.geo.back().h;
    return 1;
}

int acgpu_chain_batch(const uint8_t *src, ImageFormat fmt, int width, int height, size_t spitch, uint8_t *dest, size_t dpitch,
                      const acgpu_chain_op *ops, int nops, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!src || !dest) { set_error("acgpu_chain_batch: null frame pointer"); return 0; }
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    if (nframes <= 0) return 1;
    const size_t inb = geo_bytes(pl.geo.front()), outb = geo_bytes(pl.geo.back());
    if (spitch == 0) spitch = inb;
    if (dpitch == 0) dpitch = outb;
    if (nframes > 1 && (spitch < inb || dpitch < outb)) { set_error("acgpu_chain_batch: frame pitch smaller than a frame"); return 0; }
    {
        const size_t sspan = spitch * (size_t)(nframes - 1) + inb, dspan = dpitch * (size_t)(nframes - 1) + outb;
        if (src < dest + dspan && dest < src + sspan) { set_error("acgpu_chain_batch: src and dest overlap"); return 0; }
    }
    cudaStream_t st = pick_stream(c, stream);
    // Intermediate frames live in the thread's temporary, two buffers of `sub` frames each; the batch is walked in
    // sub-batches of that many frames.  Bigger is better: sub-batches small enough to keep the intermediates in the 126 MB
    // L2 were measured and lost (UHD 420P -> RGB24 -> 422P: 48.9 k frames/s with one frame per launch, 64.9 k with two,
    // 82.9 k with 32 or more -- launches of one or two frames leave most of the machine in the tail of each kernel, which
    // costs more than the HBM round trip of the intermediate saves).  $ACGPU_CHAIN_SCRATCH_BYTES bounds the temporary
    // (default 2 GiB); a chain of one moving stage has no intermediate and runs as one launch over the whole batch.
    const size_t tp = align_up(pl.max_bytes + 256, 256);
    const bool needs_scratch = !(pl.n_live == 0 || (pl.n_live == 1 && pl.n_out == 1 && !pl.two_pass));
    int sub = nframes;
    if (needs_scratch) {
        static const size_t budget = [] { const char *e = getenv("ACGPU_CHAIN_SCRATCH_BYTES"); return e ? (size_t)atoll(e) : (size_t)2 << 30; }();
        sub = (int)(budget / (2 * tp));
        if (sub < 1) sub = 1;
        if (sub > nframes) sub = nframes;
    }
    Buf s0{nullptr, tp}, s1{nullptr, tp};
    if (needs_scratch) {
        if (!ensure_arena(c, 2 * tp * (size_t)sub) || !arena_acquire(c, st)) return 0;
        s0.p = c->arena;
        s1.p = c->arena + tp * (size_t)sub;
    }
    for (int f0 = 0; f0 < nframes; f0 += sub) {
        const int nf = nframes - f0 < sub ? nframes - f0 : sub;
        Buf res;
        if (!run_chain(c, stream, pl, ops, nops, Buf{const_cast<uint8_t *>(src) + (size_t)f0 * spitch, spitch}, false, s0, s1,
                       Buf{dest + (size_t)f0 * dpitch, dpitch}, nf, &res))
            return 0;
    }
    return needs_scratch ? (arena_release(c, st) ? 1 : 0) : 1;
}

// The pipeline behind the host-frame chain calls.  Frame i of the run lies at src_at(i) / dst_at(i) in host memory (best:
// acgpu_host_alloc / acgpu_bufalloc memory).  A few pipeline slots, each with its own stream: upload chunk k+1 and download
// chunk k-1 while chunk k is processed.  `contiguous`: the frames are one tightly packed run, so a chunk travels as one
// strided copy; otherwise every frame is a copy of its own (transcode's frame ring: one tc_bufalloc'd buffer per
// vframe_list_t, tccore/frame.h:215-253).
extern "C++" {
template <class SrcAt, class DstAt>
static int chain_pipeline(const char *who, DevCtx *c, const Plan &pl, const acgpu_chain_op *ops, int nops, int nframes,
                          bool contiguous, SrcAt src_at, DstAt dst_at)
{
    const size_t inb = geo_bytes(pl.geo.front()), outb = geo_bytes(pl.geo.back());
    const size_t ip = align_up(inb, 256), tp = align_up(pl.max_bytes + 256, 256);
    // The last conversion may leave destination bytes alone (alpha of YUV -> 32-bit RGB): the caller's destination frames
    // are then uploaded first, as the one-conversion call does, into a result region of their own.
    bool preload = false;
    for (int k = nops - 1; k >= 0; k--) {
        if (stage_noop(ops[k], pl.geo[(size_t)k])) continue;
        if (ops[k].kind == ACGPU_CHAIN_CONVERT) {
            const int sfmt = pl.geo[(size_t)k].fmt, dfmt = ops[k].p[0];
            const FmtDesc sd = describe(sfmt == IMG_YV12 ? IMG_YUV420P : sfmt), dd = describe(dfmt == IMG_YV12 ? IMG_YUV420P : dfmt);
            preload = dd.kind == K_RGB && dd.bpp == 4 && (sd.kind == K_PLANAR || sd.kind == K_PACKED || sd.kind == K_Y8);
        }
        break;
    }
    const size_t big = ip > tp ? ip : tp;
    size_t per = pipe_chunk_bytes() / big;
    if (per < 1) per = 1;
    if (per > (size_t)nframes) per = (size_t)nframes;
    const size_t slot_bytes = per * (ip + (preload ? 3 : 2) * tp);
    for (int s = 0; s < pipe_slots(); s++) {
        if (!c->pipe_stream[s] && !check(cudaStreamCreateWithFlags(&c->pipe_stream[s], cudaStreamNonBlocking), "pipe stream")) return 0;
        if (c->pipe_cap[s] < slot_bytes) {
            if (c->pipe_buf[s]) { cudaStreamSynchronize(c->pipe_stream[s]); cudaFree(c->pipe_buf[s]); c->pipe_buf[s] = nullptr; c->pipe_cap[s] = 0; }
            if (!check(cudaMalloc(&c->pipe_buf[s], slot_bytes), "cudaMalloc(pipeline slot)")) return 0;
            c->pipe_cap[s] = slot_bytes;
        }
    }
    // host run <-> device rows of `dev_pitch` bytes, frames f0 .. f0+n-1
    auto copy_frames = [&](uint8_t *dev, size_t dev_pitch, int f0, int n, size_t bytes, bool to_device, bool from_dst, cudaStream_t st) {
        if (contiguous) {
            uint8_t *host = from_dst ? dst_at(f0) : const_cast<uint8_t *>(src_at(f0));
            return to_device ? check(cudaMemcpy2DAsync(dev, dev_pitch, host, bytes, bytes, (size_t)n, cudaMemcpyHostToDevice, st), "H2D frames")
                             : check(cudaMemcpy2DAsync(host, bytes, dev, dev_pitch, bytes, (size_t)n, cudaMemcpyDeviceToHost, st), "D2H frames");
        }
        for (int i = 0; i < n; i++) {
            uint8_t *host = from_dst ? dst_at(f0 + i) : const_cast<uint8_t *>(src_at(f0 + i));
            const bool ok = to_device ? check(cudaMemcpyAsync(dev + (size_t)i * dev_pitch, host, bytes, cudaMemcpyHostToDevice, st), "H2D frame")
                                      : check(cudaMemcpyAsync(host, dev + (size_t)i * dev_pitch, bytes, cudaMemcpyDeviceToHost, st), "D2H frame");
            if (!ok) return false;
        }
        return true;
    };
    int chunk = 0;
    for (int f0 = 0; f0 < nframes; f0 += (int)per, chunk++) {
        const int s = chunk % pipe_slots();
        const int n = nframes - f0 < (int)per ? nframes - f0 : (int)per;
        cudaStream_t st = c->pipe_stream[s];
        acgpu_stream_t as = reinterpret_cast<acgpu_stream_t>(st);
        Buf in{c->pipe_buf[s], ip}, s0{in.p + per * ip, tp}, s1{s0.p + per * tp, tp}, out{nullptr, 0};
        if (!copy_frames(in.p, ip, f0, n, inb, true, false, st)) return 0;
        if (preload) {
            out = Buf{s1.p + per * tp, tp};
            if (!copy_frames(out.p, tp, f0, n, outb, true, true, st)) return 0;
        }
        Buf res;
        if (!run_chain(c, as, pl, ops, nops, in, true, s0, s1, out, n, &res)) return 0;
        if (!copy_frames(res.p, res.pitch, f0, n, outb, false, true, st)) return 0;
    }
    for (int s = 0; s < pipe_slots(); s++)
        if (!check(cudaStreamSynchronize(c->pipe_stream[s]), who)) return 0;
    return 1;
}
}  // extern "C++"

// Host frames: tightly packed runs in `src_frames` / `dest_frames`.
int acgpu_chain_frames_host(const uint8_t *src_frames, ImageFormat fmt, int width, int height, uint8_t *dest_frames,
                            const acgpu_chain_op *ops, int nops, int nframes)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!src_frames || !dest_frames) { set_error("acgpu_chain_frames_host: null frame pointer"); return 0; }
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    if (nframes <= 0) return 1;
    const size_t inb = geo_bytes(pl.geo.front()), outb = geo_bytes(pl.geo.back());
    return chain_pipeline("acgpu_chain_frames_host", c, pl, ops, nops, nframes, true,
                          [=](int i) { return src_frames + (size_t)i * inb; }, [=](int i) { return dest_frames + (size_t)i * outb; });
}

// Host frames that each have a buffer of their own: src_frames[i] / dest_frames[i] point at frame i.
int acgpu_chain_frame_list_host(const uint8_t *const *src_frames, ImageFormat fmt, int width, int height,
                                uint8_t *const *dest_frames, const acgpu_chain_op *ops, int nops, int nframes)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    if (nframes <= 0) return 1;
    if (!src_frames || !dest_frames) { set_error("acgpu_chain_frame_list_host: null frame list"); return 0; }
    for (int i = 0; i < nframes; i++)
        if (!src_frames[i] || !dest_frames[i]) { set_error("acgpu_chain_frame_list_host: frame %d is a null pointer", i); return 0; }
    return chain_pipeline("acgpu_chain_frame_list_host", c, pl, ops, nops, nframes, false,
                          [=](int i) { return src_frames[i]; }, [=](int i) { return dest_frames[i]; });
}

int acgpu_chain_frames_host_multi(const uint8_t *src_frames, ImageFormat fmt, int width, int height, uint8_t *dest_frames,
                                  const acgpu_chain_op *ops, int nops, int nframes, int ndevices)
{
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    const size_t inb = geo_bytes(pl.geo.front()), outb = geo_bytes(pl.geo.back());
    return run_on_devices("acgpu_chain_frames_host_multi", ndevices, nframes, [=](int, int f0, int f1) {
        return acgpu_chain_frames_host(src_frames + (size_t)f0 * inb, fmt, width, height, dest_frames + (size_t)f0 * outb, ops, nops, f1 - f0) == 1;
    });
}

int acgpu_chain_frame_list_host_multi(const uint8_t *const *src_frames, ImageFormat fmt, int width, int height,
                                      uint8_t *const *dest_frames, const acgpu_chain_op *ops, int nops, int nframes, int ndevices)
{
    Plan pl;
    if (!make_plan(fmt, width, height, ops, nops, &pl)) return 0;
    if (nframes > 0 && (!src_frames || !dest_frames)) { set_error("acgpu_chain_frame_list_host_multi: null frame list"); return 0; }
    return run_on_devices("acgpu_chain_frame_list_host_multi", ndevices, nframes, [=](int, int f0, int f1) {
        return acgpu_chain_frame_list_host(src_frames + f0, fmt, width, height, dest_frames + f0, ops, nops, f1 - f0) == 1;
    });
}

}  // extern "C"
