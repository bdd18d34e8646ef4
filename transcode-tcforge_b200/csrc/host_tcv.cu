// host_tcv.cu -- the frame-granular libtcvideo entry points of libacgpu's C ABI (include/acgpu.h): tcv_deinterlace,
// tcv_resize, tcv_convert, the -K grayscale chain, tcv_clip, tcv_reduce, tcv_flip_v/h, tcv_gamma_correct, tcv_antialias
// (libtcvideo/tcvideo.c, src/video_trans.c).  Host side only: argument checks with the reference's rules and return
// values, the small tables the reference builds on the host (same double arithmetic), and one launch per batch.
#include "host_ctx.h"

#include <math.h>

#include <algorithm>
#include <vector>

namespace acgpu {
namespace {

// libtcvideo/tcvideo.c:1138-1165 -- sin^2-weighted two-tap table, newsize/8 entries.
void build_resize_table(int oldsize, int newsize, std::vector<int32_t> &src, std::vector<uint32_t> &w1,
                        std::vector<uint32_t> &w2)
{
    const int n = newsize / 8;
    src.resize(n); w1.resize(n); w2.resize(n);
    const double ratio = (double)oldsize / (double)newsize;
    for (int i = 0; i < n; i++) {
        const double pos = (double)i * (double)oldsize / (double)newsize;
        const int s = (int)pos;
        src[i] = s;
        if (pos + ratio < s + 1) {
            w1[i] = 65536; w2[i] = 0;
        } else {
            const double t = ((s + 1) - pos) / ratio * M_PI / 2;
            w1[i] = (uint32_t)(sin(t) * sin(t) * 65536 + 0.5);
            w2[i] = 65536 - w1[i];
        }
    }
}

// Host frames for the frame-granular entry points.  These functions are written for device-resident planes; when the
// caller's source and/or destination is host memory (unmodified libtcvideo callers hold host frame buffers) the planes go
// through a per-thread device staging buffer: one upload of the batch, the operation on the device copies, one download, and the
// call returns after the result has landed -- the same contract as the legacy per-frame ac_imgconvert path, but one
// round trip per FRAME instead of one per ROW (tcv_deinterlace / tcv_resize call ac_average / ac_rescale per row).
struct HostStage {
    DevCtx *c = nullptr;
    bool staged = false, ok = true, dst_host = false, lone = true;
    int dst_kind = 0;
    StagedCall *busy = nullptr;
    const uint8_t *dsrc = nullptr;
    uint8_t *ddst = nullptr, *dest = nullptr;
    size_t dsp = 0, ddp = 0, dp_host = 0, out_bytes = 0;
    int nf = 0;

    // preload_dest: the operation may leave destination bytes alone (alpha of YUV -> 32-bit RGB), so the host destination
    // is uploaded first and survives the round trip
    // device_in_place: when src == dst the operation may run in place on the device copy (flips); otherwise the device
    // copy gets its own destination region and only the download lands on the shared host buffer
    HostStage(const uint8_t *src, size_t in_bytes, size_t spitch, uint8_t *dst, size_t outb, size_t dpitch, int nframes,
              acgpu_stream_t caller_stream, bool preload_dest = false, bool device_in_place = true)
    {
        if (nframes <= 0 || tls.device_only > 0) return;
        const int src_kind = pointer_kind(src);
        dst_kind = src == dst ? src_kind : pointer_kind(dst);
        const bool src_host = src_kind != 2;
        dst_host = dst_kind != 2;
        if (!src_host && !dst_host) return;
        staged = true;
        busy = new StagedCall();
        lone = busy->others == 0;
        c = ctx();
        if (!c) { ok = false; return; }
        // the staged sequence runs on the thread's private stream: a device-resident side may still be in flight on the
        // stream the caller named
        if (!order_after(c, caller_stream)) { ok = false; return; }
        dest = dst; out_bytes = outb; nf = nframes;
        const size_t sp_host = spitch ? spitch : in_bytes;
        dp_host = dpitch ? dpitch : outb;
        const bool in_place = device_in_place && static_cast<const void *>(src) == static_cast<const void *>(dst);
        dsp = src_host ? align_up(in_place && outb > in_bytes ? outb : in_bytes, 256) : sp_host;    // in place: room for the larger side
        ddp = in_place ? dsp : dst_host ? align_up(outb, 256) : dp_host;
        const size_t src_region = src_host ? dsp * (size_t)nframes : 0, dst_region = (dst_host && !in_place) ? ddp * (size_t)nframes : 0;
        // (always used on the thread's own stream and every staged call ends with a sync: stream order is enough)
        if (c->plane_stage_cap < src_region + dst_region + 256) {
            if (c->plane_stage) { cudaStreamSynchronize(c->stream); cudaFree(c->plane_stage); c->plane_stage = nullptr; c->plane_stage_cap = 0; }
            const size_t cap = align_up(src_region + dst_region + 256, 1u << 20);
            if (!check(cudaMalloc(&c->plane_stage, cap), "cudaMalloc(plane staging)")) { ok = false; return; }
            c->plane_stage_cap = cap;
        }
        uint8_t *as = c->plane_stage, *ad = c->plane_stage + src_region;
        if (src_host) {
            ok = staged_h2d(c, as, dsp, src, sp_host, in_bytes, (size_t)nframes, c->stream, src_kind, lone);
            dsrc = as;
        } else {
            dsrc = src;
        }
        ddst = in_place ? const_cast<uint8_t *>(dsrc) : dst_host ? ad : dst;
        if (ok && preload_dest && dst_host && !in_place && outb)
            ok = staged_h2d(c, ad, ddp, dst, dp_host, outb, (size_t)nframes, c->stream, dst_kind, lone);
    }
    acgpu_stream_t stream() const { return reinterpret_cast<acgpu_stream_t>(c->stream); }
    int finish(int launched)
    {
        bool good = ok && launched;
        if (good && dst_host && out_bytes) good = staged_d2h(c, dest, dp_host, ddst, ddp, out_bytes, (size_t)nf, c->stream, dst_kind, lone);
        if (c) good = check(cudaStreamSynchronize(c->stream), "frame operation") && good;
        return good ? 1 : 0;
    }
    ~HostStage() { delete busy; }
    HostStage(const HostStage &) = delete;
    HostStage &operator=(const HostStage &) = delete;
};

// One pitch rule for every entry point: 0 means tightly packed frames, and consecutive frames must not overlap.
bool norm_pitch(const char *who, size_t *pitch, size_t frame_bytes, int nframes)
{
    if (*pitch == 0) *pitch = frame_bytes;
    if (nframes > 1 && *pitch < frame_bytes) {
        set_error("%s: frame pitch %zu is smaller than a frame (%zu bytes)", who, *pitch, frame_bytes);
        return false;
    }
    return true;
}

// src/dest relation over the whole batch: 0 = disjoint, 1 = the same buffer (src == dest), -1 = partial overlap (rejected;
// the reference's precondition is "src and dest do not overlap" unless they are equal, tcvideo.c:180).
int alias_state(const char *who, const uint8_t *src, size_t spitch, size_t inb, const uint8_t *dst, size_t dpitch, size_t outb, int nframes)
{
    if (src == dst) return 1;
    const size_t sspan = spitch * (size_t)(nframes - 1) + inb, dspan = dpitch * (size_t)(nframes - 1) + outb;
    if (src < dst + dspan && dst < src + sspan) {
        set_error("%s: src and dest overlap", who);
        return -1;
    }
    return 0;
}

// src == dest on the device for an operation whose rows read other rows: the result is produced in the thread's temporary and
// copied back, which equals the reference's sequential in-place result whenever every destination byte lies at or before
// the source bytes it depends on (the callers check that; other in-place uses are rejected).
template <class Op>
int through_temp(DevCtx *c, cudaStream_t st, uint8_t *dest, size_t dpitch, size_t outb, int nframes, Op op)
{
    const size_t tp = align_up(outb, 256);
    if (!ensure_arena(c, tp * (size_t)nframes) || !arena_acquire(c, st)) return 0;
    if (!op(c->arena, tp)) return 0;
    return check(cudaMemcpy2DAsync(dest, dpitch, c->arena, tp, outb, (size_t)nframes, cudaMemcpyDeviceToDevice, st), "in-place copy back")
        && arena_release(c, st) ? 1 : 0;
}

}  // namespace
}  // namespace acgpu

using namespace acgpu;

extern "C" {

// ---- libtcvideo shapes ----------------------------------------------------------------------------------
int acgpu_deinterlace_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int mode,
                            size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:290-311 argument checks
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("acgpu_deinterlace_batch: invalid frame parameters"); return 0; }
    if (mode < ACGPU_DEINT_INTERPOLATE || mode > ACGPU_DEINT_DROP_FIELD_BOTTOM) { set_error("acgpu_deinterlace_batch: invalid mode %d", mode); return 0; }
    const int64_t Bpl = (int64_t)width * Bpp;
    if (nframes <= 0) return 1;
    const bool drop = mode == ACGPU_DEINT_DROP_FIELD_TOP || mode == ACGPU_DEINT_DROP_FIELD_BOTTOM;
    const size_t inb = (size_t)Bpl * height, outb = (size_t)Bpl * (drop ? height / 2 : height);
    if (!norm_pitch("acgpu_deinterlace_batch", &spitch, inb, nframes) || !norm_pitch("acgpu_deinterlace_batch", &dpitch, outb, nframes)) return 0;
    const int alias = alias_state("acgpu_deinterlace_batch", src, spitch, inb, dest, dpitch, outb, nframes);
    if (alias < 0) return 0;
    if (alias && mode == ACGPU_DEINT_LINEAR_BLEND) {
        // tcvideo.c:368-389 in place averages rows that were already blended: not a result worth reproducing
        set_error("acgpu_deinterlace_batch: linear blend needs src != dest");
        return 0;
    }
    {
        HostStage hs(src, inb, spitch, dest, outb, dpitch, nframes, stream, false, /*device_in_place=*/false);
        if (hs.staged)
            return hs.finish(hs.ok && acgpu_deinterlace_batch(hs.dsrc, hs.ddst, width, height, Bpp, mode, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    if (alias) {    // interpolate / drop field in place: every row depends on rows at or after itself (tcvideo.c:326-364)
        DevCtx *c = ctx();
        if (!c) return 0;
        return through_temp(c, pick_stream(c, stream), dest, dpitch, outb, nframes, [&](uint8_t *t, size_t tp) {
            return acgpu_deinterlace_batch(src, t, width, height, Bpp, mode, spitch, tp, nframes, stream) == 1; });
    }
    if (mode == ACGPU_DEINT_DROP_FIELD_TOP || mode == ACGPU_DEINT_DROP_FIELD_BOTTOM) {
        // tcvideo.c:326-338: keep every other line, starting at line 1 when the top field is dropped
        std::vector<acgpu_rowop> ops((size_t)(height / 2));
        for (int y = 0; y < height / 2; y++) {
            acgpu_rowop o{};
            o.op = ACGPU_ROW_COPY;
            o.src1_off = (int64_t)(2 * y + (mode == ACGPU_DEINT_DROP_FIELD_TOP ? 1 : 0)) * Bpl;
            o.dest_off = y * Bpl;
            ops[(size_t)y] = o;
        }
        return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), height / 2, (int)Bpl, nframes, stream);
    }
    std::vector<acgpu_rowop> ops((size_t)height);
    for (int y = 0; y < height; y++) {
        acgpu_rowop o{};
        o.dest_off = y * Bpl;
        if (mode == ACGPU_DEINT_INTERPOLATE || height < 2) {
            // tcvideo.c:353-364: even rows copied, odd rows = mean of neighbours, odd last row = copy of y-1
            if (y % 2 == 0)            { o.op = ACGPU_ROW_COPY;    o.src1_off = y * Bpl; }
            else if (y == height - 1)  { o.op = ACGPU_ROW_COPY;    o.src1_off = (y - 1) * Bpl; }
            else                       { o.op = ACGPU_ROW_AVERAGE; o.src1_off = (y - 1) * Bpl; o.src2_off = (y + 1) * Bpl; }
        } else {
            // tcvideo.c:368-389 fused: out = avg(A, B) where one of A/B is the source row and the other the
            // mean of its neighbours -- the same expression for odd and even interior rows; the first and
            // last rows average with their single neighbour (the copies made at :377 and :381).
            if (y == 0)               { o.op = ACGPU_ROW_AVERAGE; o.src1_off = Bpl;           o.src2_off = 0; }
            else if (y == height - 1) { o.op = ACGPU_ROW_AVERAGE; o.src1_off = (y - 1) * Bpl; o.src2_off = y * Bpl; }
            else { o.op = ACGPU_ROW_AVERAGE3; o.src1_off = (y - 1) * Bpl; o.src2_off = (y + 1) * Bpl; o.src3_off = y * Bpl; }
        }
        ops[(size_t)y] = o;
    }
    return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), height, (int)Bpl, nframes, stream);
}

int acgpu_resize_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int resize_w, int resize_h,
                       int scale_w, int scale_h, size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:436-457 argument checks
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("acgpu_resize_batch: invalid frame parameters"); return 0; }
    auto ok_scale = [](int s) { return s == 1 || s == 2 || s == 4 || s == 8; };
    if (!ok_scale(scale_w) || !ok_scale(scale_h)) { set_error("acgpu_resize_batch: invalid scale parameters"); return 0; }
    if (width % scale_w != 0 || height % scale_h != 0) { set_error("acgpu_resize_batch: scale does not divide the frame"); return 0; }
    const int new_w = width + resize_w * scale_w, new_h = height + resize_h * scale_h;
    if (new_w <= 0 || new_h <= 0) { set_error("acgpu_resize_batch: resulting size is not positive"); return 0; }
    if (resize_w && resize_h && new_h > height) {
        // tcvideo.c:459-512 with both set: the horizontal pass reads rows 0..new_h-1 of SRC (not the vertically resized
        // rows) and overwrites whatever the vertical pass left, so the result is defined only while those rows exist
        set_error("acgpu_resize_batch: resize_w and resize_h both set with a taller result reads past the source frame");
        return 0;
    }
    if (nframes <= 0) return 1;
    const size_t inb = (size_t)width * height * Bpp, outb = (size_t)new_w * new_h * Bpp;
    if (!norm_pitch("acgpu_resize_batch", &spitch, inb, nframes) || !norm_pitch("acgpu_resize_batch", &dpitch, outb, nframes)) return 0;
    const int alias = alias_state("acgpu_resize_batch", src, spitch, inb, dest, dpitch, outb, nframes);
    if (alias < 0) return 0;
    if (alias && (resize_w > 0 || resize_h > 0 || (resize_w && resize_h))) {
        set_error("acgpu_resize_batch: enlarging in place would read bytes already overwritten; use src != dest");
        return 0;
    }
    {
        HostStage hs(src, inb, spitch, dest, outb, dpitch, nframes, stream, false, /*device_in_place=*/false);
        if (hs.staged)
            return hs.finish(hs.ok && acgpu_resize_batch(hs.dsrc, hs.ddst, width, height, Bpp, resize_w, resize_h, scale_w, scale_h,
                                                         hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    if (alias && (resize_w || resize_h)) {   // shrinking in place: destination bytes never lie after the source bytes they read
        DevCtx *c = ctx();
        if (!c) return 0;
        return through_temp(c, pick_stream(c, stream), dest, dpitch, outb, nframes, [&](uint8_t *t, size_t tp) {
            return acgpu_resize_batch(src, t, width, height, Bpp, resize_w, resize_h, scale_w, scale_h, spitch, tp, nframes, stream) == 1; });
    }
    DevCtx *c = ctx();
    if (!c) return 0;
    cudaStream_t st = pick_stream(c, stream);
    std::vector<int32_t> ts;
    std::vector<uint32_t> w1, w2;
    if (resize_h && !resize_w) {
        const int64_t Bpl = (int64_t)width * Bpp;
        build_resize_table(height * 8 / scale_h, new_h * 8 / scale_h, ts, w1, w2);
        const int rows = new_h / scale_h;
        std::vector<acgpu_rowop> ops((size_t)rows * scale_h);
        for (int i = 0; i < scale_h; i++)
            for (int y = 0; y < rows; y++) {
                acgpu_rowop o{};
                o.op = ACGPU_ROW_RESCALE;
                o.src1_off = ((int64_t)i * (height / scale_h) + ts[y]) * Bpl;
                o.src2_off = o.src1_off + Bpl;
                o.dest_off = ((int64_t)i * rows + y) * Bpl;
                o.weight1 = w1[y];
                o.weight2 = w2[y];
                ops[(size_t)i * rows + y] = o;
            }
        return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), (int)ops.size(), (int)Bpl, nframes, stream);
    }
    if (resize_w) {
        build_resize_table(width * 8 / scale_w, new_w * 8 / scale_w, ts, w1, w2);
        const int n = (int)ts.size();
        const size_t sp_ = spitch, dp_ = dpitch;
        if (resize_h_vectorisable(src, sp_, dest, dp_, width, new_w, Bpp)) {
            // per-row byte tables: source byte offset (first tap) and packed weights for every output byte of a row
            const int src_block = width / scale_w, dst_block = new_w / scale_w;
            std::vector<uint16_t> off((size_t)new_w * Bpp);
            std::vector<uint32_t> wgt((size_t)new_w * Bpp);
            for (int b = 0; b < scale_w; b++)
                for (int x = 0; x < dst_block; x++)
                    for (int k = 0; k < Bpp; k++) {
                        const size_t o = ((size_t)b * dst_block + x) * Bpp + k;
                        size_t so = ((size_t)b * src_block + ts[x]) * Bpp + k;
                        uint32_t wp;
                        if (w1[x] >= 0x10000u) wp = 0x0000FFFFu;                          // tap 1 untouched
                        else if (w2[x] >= 0x10000u) { wp = 0x0000FFFFu; so += Bpp; }      // w1 == 0: tap 2 untouched
                        else wp = (w1[x] & 0xFFFFu) | (w2[x] << 16);
                        off[o] = (uint16_t)so;
                        wgt[o] = wp;
                    }
            const uint32_t *dwgt = static_cast<const uint32_t *>(device_blob(c, wgt.data(), wgt.size() * sizeof(uint32_t), st));
            if (!dwgt) return 0;
            // window form: the four first taps of every output word within 8 source bytes (any ratio up to ~2:1);
            // narrow: the second taps that carry weight lie inside those 8 bytes too (a second tap of weight 0 may select
            // any byte: its selector is clamped into the window)
            std::vector<uint2> meta(off.size() / 4);
            bool windowed = tls.force_tier != 1, narrow = true;
            for (size_t ow = 0; ow < meta.size() && windowed; ow++) {
                uint32_t lo = off[4 * ow], hi = lo;
                for (int j = 1; j < 4; j++) { lo = std::min<uint32_t>(lo, off[4 * ow + j]); hi = std::max<uint32_t>(hi, off[4 * ow + j]); }
                if (hi - lo > 7) { windowed = false; break; }
                uint32_t selA = 0, selB = 0;
                for (int j = 0; j < 4; j++) {
                    const uint32_t t = off[4 * ow + j] - lo, t2 = t + (uint32_t)Bpp;
                    selA |= t << (4 * j);
                    if (t2 > 7 && (wgt[4 * ow + j] >> 16) != 0) narrow = false;
                    selB |= (t2 > 7 ? 7u : t2) << (4 * j);
                }
                meta[ow] = make_uint2(lo | (selA << 16), selB);
            }
            if (windowed) {
                const uint2 *dmeta = static_cast<const uint2 *>(device_blob(c, meta.data(), meta.size() * sizeof(uint2), st));
                if (!dmeta) return 0;
                for (int f0 = 0; f0 < nframes; f0 += 32768) {
                    const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
                    if (!resize_h_win_launch(src + (size_t)f0 * sp_, sp_, dest + (size_t)f0 * dp_, dp_, dmeta, dwgt, width, new_w,
                                             new_h, Bpp, narrow, nf, st))
                        return 0;
                }
                return 1;
            }
            const uint16_t *doff = static_cast<const uint16_t *>(device_blob(c, off.data(), off.size() * sizeof(uint16_t), st));
            if (!doff) return 0;
            for (int f0 = 0; f0 < nframes; f0 += 32768) {
                const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
                if (!resize_h_row_launch(src + (size_t)f0 * sp_, sp_, dest + (size_t)f0 * dp_, dp_, doff, dwgt, width, new_w,
                                         new_h, Bpp, nf, st))
                    return 0;
            }
            return 1;
        }
        const int32_t *dts = static_cast<const int32_t *>(device_blob(c, ts.data(), sizeof(int32_t) * n, st));
        const uint32_t *dw1 = static_cast<const uint32_t *>(device_blob(c, w1.data(), sizeof(uint32_t) * n, st));
        const uint32_t *dw2 = static_cast<const uint32_t *>(device_blob(c, w2.data(), sizeof(uint32_t) * n, st));
        if (!dts || !dw1 || !dw2) return 0;
        for (int f0 = 0; f0 < nframes; f0 += 32768) {
            const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
            if (!resize_h_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, dts, dw1, dw2,
                                 width, new_w, new_h, Bpp, scale_w, nf, st))
                return 0;
        }
        return 1;
    }
    // no resize requested: the reference leaves dest untouched (tcvideo.c:459,481)
    return 1;
}

int acgpu_convert_batch(uint8_t *src, uint8_t *dest, int width, int height, ImageFormat srcfmt, ImageFormat destfmt,
                        size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:1001-1067
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!src || !dest || width <= 0 || height <= 0 || !srcfmt || !destfmt) { set_error("acgpu_convert_batch: invalid image parameters"); return 0; }
    const int sf = srcfmt == IMG_YV12 ? IMG_YUV420P : (int)srcfmt, df = destfmt == IMG_YV12 ? IMG_YUV420P : (int)destfmt;
    if (describe(sf).kind == K_NONE || describe(df).kind == K_NONE) { set_error("acgpu_convert_batch: unknown format"); return 0; }
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    const size_t sfb = frame_bytes(sf, width, height), dfb = frame_bytes(df, width, height);
    if (!norm_pitch("acgpu_convert_batch", &spitch, sfb, nframes) || !norm_pitch("acgpu_convert_batch", &dpitch, dfb, nframes)) return 0;
    if (src == dest && nframes > 1 && spitch != dpitch) { set_error("acgpu_convert_batch: in place needs one pitch"); return 0; }
    if (alias_state("acgpu_convert_batch", src, spitch, sfb, dest, dpitch, dfb, nframes) < 0) return 0;
    {
        // (src == dest on the host: the device copy converts into a separate region, no device-side temporary needed)
        HostStage hs(src, sfb, spitch, dest, dfb, dpitch, nframes, stream, /*preload_dest=*/src != dest, /*device_in_place=*/false);
        if (hs.staged)
            return hs.finish(hs.ok && acgpu_convert_batch(const_cast<uint8_t *>(hs.dsrc), hs.ddst, width, height, srcfmt, destfmt,
                                                          hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    if (srcfmt == destfmt) {
        if (src == dest) return 1;
        return check(cudaMemcpy2DAsync(dest, dpitch, src, spitch, dfb, nframes, cudaMemcpyDeviceToDevice, st), "acgpu_convert_batch copy") ? 1 : 0;
    }
    uint8_t *real = dest;
    size_t rpitch = dpitch;
    if (src == dest) {                    // in place: convert into a temporary, then copy back (tcvideo.c:1044-1064)
        rpitch = align_up(dfb, 256);
        if (!ensure_arena(c, rpitch * (size_t)nframes) || !arena_acquire(c, st)) return 0;
        real = c->arena;
    }
    uint8_t *sp[3], *dp[3];
    sp[0] = src;  sp[1] = src + (size_t)width * height;  sp[2] = sp[1] + chroma_plane_bytes(sf, width, height);
    dp[0] = real; dp[1] = real + (size_t)width * height; dp[2] = dp[1] + chroma_plane_bytes(df, width, height);
    if (!acgpu_imgconvert_batch(sp, srcfmt, spitch, dp, destfmt, rpitch, width, height, nframes, stream)) return 0;
    if (src == dest)
        return check(cudaMemcpy2DAsync(dest, dpitch, real, rpitch, dfb, nframes, cudaMemcpyDeviceToDevice, st),
                     "acgpu_convert_batch copy back") && arena_release(c, st) ? 1 : 0;
    return 1;
}

int acgpu_decolor_rgb24_batch(uint8_t *frames, int width, int height, size_t pitch, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!frames || width <= 0 || height <= 0) { set_error("acgpu_decolor_rgb24_batch: invalid frame parameters"); return 0; }
    if (nframes <= 0) return 1;
    {
        const size_t fb = (size_t)width * height * 3;
        if (!norm_pitch("acgpu_decolor_rgb24_batch", &pitch, fb, nframes)) return 0;
        HostStage hs(frames, fb, pitch, frames, fb, pitch, nframes, stream);
        if (hs.staged) return hs.finish(hs.ok && acgpu_decolor_rgb24_batch(hs.ddst, width, height, hs.ddp, nframes, hs.stream()));
    }
    cudaStream_t st = pick_stream(c, stream);
    if (tls.force_tier != 1 && per_frame_chunk(nframes, [&](int f0, int nf) {
            return decolor_rgb24_fast(frames + (size_t)f0 * pitch, pitch, width, height, nf, st); })) { tls.last_tier = 2; return 1; }
    // outside the vectorised domain: the reference's own two steps through a temporary gray plane
    const size_t gpitch = align_up((size_t)width * height, 256);
    if (!ensure_arena(c, gpitch * (size_t)nframes) || !arena_acquire(c, st)) return 0;
    uint8_t *rgb[3] = {frames, nullptr, nullptr}, *gray[3] = {c->arena, nullptr, nullptr};
    return acgpu_imgconvert_batch(rgb, IMG_RGB24, pitch, gray, IMG_GRAY8, gpitch, width, height, nframes, stream)
        && acgpu_imgconvert_batch(gray, IMG_GRAY8, gpitch, rgb, IMG_RGB24, pitch, width, height, nframes, stream)
        && arena_release(c, st);
}

// ---- the remaining element-wise libtcvideo operations ------------------------------------
static bool plane_args_ok(const char *who, const void *src, const void *dest, int width, int height, int Bpp)
{
    // the check every tcv_* function opens with (e.g. libtcvideo/tcvideo.c:192-195)
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("%s: invalid frame parameters", who); return false; }
    if ((uint64_t)width * height * Bpp >= 0x7FFFFFF0ull) { set_error("%s: plane too large", who); return false; }
    return true;
}

int acgpu_clip_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int clip_left, int clip_right,
                     int clip_top, int clip_bottom, uint8_t black_pixel, size_t spitch, size_t dpitch, int nframes,
                     acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_clip_batch", src, dest, width, height, Bpp)) return 0;
    if ((int64_t)clip_left + clip_right >= width || (int64_t)clip_top + clip_bottom >= height) {   // tcvideo.c:196-202
        set_error("acgpu_clip_batch: clipping parameters (%d,%d,%d,%d) invalid for frame size %dx%d", clip_top, clip_left,
                  clip_bottom, clip_right, width, height);
        return 0;
    }
    // tcvideo.c:204-219: a clip wider than the frame eats into the opposite (negative) border
    if (clip_left > width)    { clip_right += clip_left - width;    clip_left = width; }
    if (clip_right > width)   { clip_left += clip_right - width;    clip_right = width; }
    if (clip_top > height)    { clip_bottom += clip_top - height;   clip_top = height; }
    if (clip_bottom > height) { clip_top += clip_bottom - height;   clip_bottom = height; }
    const int64_t new_w = (int64_t)width - clip_left - clip_right, new_h = (int64_t)height - clip_top - clip_bottom;
    const int64_t copy_w = (int64_t)width - (clip_left < 0 ? 0 : clip_left) - (clip_right < 0 ? 0 : clip_right);
    if (new_w * new_h * Bpp >= 0x7FFFFFF0ll) { set_error("acgpu_clip_batch: result too large"); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    const size_t inb = (size_t)width * height * Bpp, outb = (size_t)(new_w * new_h * Bpp);
    if (!norm_pitch("acgpu_clip_batch", &spitch, inb, nframes) || !norm_pitch("acgpu_clip_batch", &dpitch, outb, nframes)) return 0;
    const int alias = alias_state("acgpu_clip_batch", src, spitch, inb, dest, dpitch, outb, nframes);
    if (alias < 0) return 0;
    if (alias && (clip_left < 0 || clip_right < 0 || clip_top < 0 || clip_bottom < 0)) {
        set_error("acgpu_clip_batch: growing a frame in place would read bytes already overwritten; use src != dest");
        return 0;
    }
    {
        HostStage hs(src, inb, spitch, dest, outb, dpitch, nframes, stream, false, /*device_in_place=*/false);
        if (hs.staged)
            return hs.finish(hs.ok && acgpu_clip_batch(hs.dsrc, hs.ddst, width, height, Bpp, clip_left, clip_right, clip_top, clip_bottom,
                                                       black_pixel, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    if (alias)      // cropping in place: every destination row lies at or before the source row it copies (tcvideo.c:229-246)
        return through_temp(c, pick_stream(c, stream), dest, dpitch, outb, nframes, [&](uint8_t *t, size_t tp) {
            return acgpu_clip_batch(src, t, width, height, Bpp, clip_left, clip_right, clip_top, clip_bottom, black_pixel, spitch, tp, nframes, stream) == 1; });
    TcvWindow p{};
    p.src = src; p.spitch = spitch; p.dst = dest; p.dpitch = dpitch;
    p.dBpl = (uint32_t)(new_w * Bpp); p.sBpl = (uint32_t)width * Bpp;
    p.drows = (int)new_h; p.srows = height;
    p.row_mul = 1; p.row_add = clip_top;
    p.cl = (uint32_t)((clip_left < 0 ? -(int64_t)clip_left : 0) * Bpp);
    p.cn = (uint32_t)((copy_w > 0 ? copy_w : 0) * Bpp);
    p.sxb = (uint32_t)((clip_left > 0 ? clip_left : 0) * Bpp);
    p.fill = 0x01010101u * black_pixel;
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        TcvWindow q = p;
        q.src += (size_t)f0 * spitch; q.dst += (size_t)f0 * dpitch;
        return tcv_window_launch(q, nf, st);
    });
}

int acgpu_reduce_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h,
                       size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_reduce_batch", src, dest, width, height, Bpp)) return 0;
    if (reduce_w <= 0 || reduce_h <= 0) { set_error("acgpu_reduce_batch: invalid reduction parameters (%d,%d)", reduce_w, reduce_h); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    const size_t inb = (size_t)width * height * Bpp;
    const size_t outb = (size_t)(reduce_w != 1 ? width / reduce_w : width) * (size_t)(height / reduce_h) * Bpp;
    if (!norm_pitch("acgpu_reduce_batch", &spitch, inb, nframes) || !norm_pitch("acgpu_reduce_batch", &dpitch, outb, nframes)) return 0;
    const int alias = alias_state("acgpu_reduce_batch", src, spitch, inb, dest, dpitch, outb, nframes);
    if (alias < 0) return 0;
    {
        HostStage hs(src, inb, spitch, dest, outb, dpitch, nframes, stream, false, /*device_in_place=*/false);
        if (hs.staged)
            return hs.finish(hs.ok && acgpu_reduce_batch(hs.dsrc, hs.ddst, width, height, Bpp, reduce_w, reduce_h, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    if (alias) {    // in place: a kept pixel never moves to a higher address (tcvideo.c:694-715)
        if (reduce_w == 1 && reduce_h == 1) return 1;
        return through_temp(c, pick_stream(c, stream), dest, dpitch, outb, nframes, [&](uint8_t *t, size_t tp) {
            return acgpu_reduce_batch(src, t, width, height, Bpp, reduce_w, reduce_h, spitch, tp, nframes, stream) == 1; });
    }
    cudaStream_t st = pick_stream(c, stream);
    if (reduce_w != 1)      // tcvideo.c:694-704
        return per_frame_chunk(nframes, [&](int f0, int nf) {
            return tcv_reduce_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, width / reduce_w,
                                     height / reduce_h, reduce_w, reduce_h, Bpp, nf, st);
        });
    // :706-715 whole rows: every reduce_h-th one, or (reduce_h == 1) the plain copy
    TcvWindow p{};
    p.src = src; p.spitch = spitch; p.dst = dest; p.dpitch = dpitch;
    p.dBpl = p.sBpl = (uint32_t)width * Bpp;
    p.drows = height / reduce_h; p.srows = height;
    p.row_mul = reduce_h; p.row_add = 0;
    p.cl = 0; p.cn = p.dBpl; p.sxb = 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        TcvWindow q = p;
        q.src += (size_t)f0 * spitch; q.dst += (size_t)f0 * dpitch;
        return tcv_window_launch(q, nf, st);
    });
}

int acgpu_flip_v_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, size_t spitch, size_t dpitch,
                       int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_flip_v_batch", src, dest, width, height, Bpp)) return 0;
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    {
        const size_t fb = (size_t)width * height * Bpp;
        if (!norm_pitch("acgpu_flip_v_batch", &spitch, fb, nframes) || !norm_pitch("acgpu_flip_v_batch", &dpitch, fb, nframes)) return 0;
        const int alias = alias_state("acgpu_flip_v_batch", src, spitch, fb, dest, dpitch, fb, nframes);
        if (alias < 0 || (alias && nframes > 1 && spitch != dpitch)) { if (alias > 0) set_error("acgpu_flip_v_batch: in place needs one pitch"); return 0; }
        HostStage hs(src, fb, spitch, dest, fb, dpitch, nframes, stream);
        if (hs.staged) return hs.finish(hs.ok && acgpu_flip_v_batch(hs.dsrc, hs.ddst, width, height, Bpp, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_flip_v_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, height, Bpp, nf, st);
    });
}

int acgpu_flip_h_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, size_t spitch, size_t dpitch,
                       int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_flip_h_batch", src, dest, width, height, Bpp)) return 0;
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    {
        const size_t fb = (size_t)width * height * Bpp;
        if (!norm_pitch("acgpu_flip_h_batch", &spitch, fb, nframes) || !norm_pitch("acgpu_flip_h_batch", &dpitch, fb, nframes)) return 0;
        const int alias = alias_state("acgpu_flip_h_batch", src, spitch, fb, dest, dpitch, fb, nframes);
        if (alias < 0 || (alias && nframes > 1 && spitch != dpitch)) { if (alias > 0) set_error("acgpu_flip_h_batch: in place needs one pitch"); return 0; }
        HostStage hs(src, fb, spitch, dest, fb, dpitch, nframes, stream);
        if (hs.staged) return hs.finish(hs.ok && acgpu_flip_h_batch(hs.dsrc, hs.ddst, width, height, Bpp, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_flip_h_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, height, Bpp, nf, st);
    });
}

int acgpu_gamma_correct_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma,
                              size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_gamma_correct_batch", src, dest, width, height, Bpp)) return 0;
    if (!(gamma > 0)) { set_error("acgpu_gamma_correct_batch: invalid gamma (%.3f)", gamma); return 0; }   // tcvideo.c:848-851
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    {
        const size_t fb = (size_t)width * height * Bpp;
        if (!norm_pitch("acgpu_gamma_correct_batch", &spitch, fb, nframes) || !norm_pitch("acgpu_gamma_correct_batch", &dpitch, fb, nframes)) return 0;
        const int alias = alias_state("acgpu_gamma_correct_batch", src, spitch, fb, dest, dpitch, fb, nframes);
        if (alias < 0 || (alias && nframes > 1 && spitch != dpitch)) { if (alias > 0) set_error("acgpu_gamma_correct_batch: in place needs one pitch"); return 0; }
        HostStage hs(src, fb, spitch, dest, fb, dpitch, nframes, stream);
        if (hs.staged) return hs.finish(hs.ok && acgpu_gamma_correct_batch(hs.dsrc, hs.ddst, width, height, Bpp, gamma, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    cudaStream_t st = pick_stream(c, stream);
    uint8_t table[256];
    for (int i = 0; i < 256; i++) table[i] = (uint8_t)(pow((i / 255.0), gamma) * 255);    // tcvideo.c:1180-1189, host doubles
    const uint8_t *d_table = static_cast<const uint8_t *>(device_blob(c, table, sizeof(table), st));
    if (!d_table) return 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_lut_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_table, (size_t)width * height * Bpp, nf, st);
    });
}

int acgpu_antialias_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias,
                          size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_antialias_batch", src, dest, width, height, Bpp)) return 0;
    if (!(weight >= 0 && weight <= 1 && bias >= 0 && bias <= 1)) {                       // tcvideo.c:899-903
        set_error("acgpu_antialias_batch: invalid antialiasing parameters (weight=%.3f, bias=%.3f)", weight, bias);
        return 0;
    }
    if (src == dest) { set_error("acgpu_antialias_batch: src and dest must not overlap (a 3x3 neighbourhood is read)"); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    {
        const size_t fb = (size_t)width * height * Bpp;
        if (!norm_pitch("acgpu_antialias_batch", &spitch, fb, nframes) || !norm_pitch("acgpu_antialias_batch", &dpitch, fb, nframes)) return 0;
        const int alias = alias_state("acgpu_antialias_batch", src, spitch, fb, dest, dpitch, fb, nframes);
        if (alias < 0 || (alias && nframes > 1 && spitch != dpitch)) { if (alias > 0) set_error("acgpu_antialias_batch: in place needs one pitch"); return 0; }
        HostStage hs(src, fb, spitch, dest, fb, dpitch, nframes, stream);
        if (hs.staged) return hs.finish(hs.ok && acgpu_antialias_batch(hs.dsrc, hs.ddst, width, height, Bpp, weight, bias, hs.dsp, hs.ddp, nframes, hs.stream()));
    }
    cudaStream_t st = pick_stream(c, stream);
    uint32_t t[1024];     // c | x | y | d, tcvideo.c:1209-1224 (double -> uint32 truncation, evaluation order kept)
    for (int i = 0; i < 256; i++) {
        t[i] = i * weight * 65536;
        t[256 + i] = i * bias * (1 - weight) / 4 * 65536;
        t[512 + i] = i * (1 - bias) * (1 - weight) / 4 * 65536;
        t[768 + i] = (t[256 + i] + t[512 + i] + 1) / 2;
    }
    const uint32_t *d_t = static_cast<const uint32_t *>(device_blob(c, t, sizeof(t), st));
    if (!d_t) return 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_antialias_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_t, width, height, Bpp, nf, st);
    });
}

}  // extern "C"
