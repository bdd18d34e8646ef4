// kernels_fast_rgb.cu -- vectorised tier, RGB family (aclib/img_rgb_packed.c): byte-order permutes between the
// six RGB layouts, 24<->32-bit repacking (alpha 0 when created, dropped when removed), RGB->GRAY8 luma,
// GRAY8->RGB replication, and Y8->RGB (range-mapped luma replicated, alpha untouched; img_yuv_rgb.c:354-379).
// A thread owns 16 pixels; pixel words are normalised by load_rgb16, permuted with one PRMT each, and written
// back through the per-warp staging buffer so global stores are 512-byte contiguous.
#include "fast_common.cuh"

namespace acgpu {

bool launch_wordperm(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, uint32_t sel, size_t bytes,
                     int nframes, cudaStream_t st);   // kernels_fast_yuv.cu

namespace {

using namespace fast;

template <int L> struct RgbPos {
    using RI = RgbInfo<L>;
    static constexpr int bpp = RI::bpp, r = RI::rpos, g = RI::gpos, b = RI::bpos;
    static constexpr int a = (L == L_RGBA || L == L_BGRA) ? 3 : (L == L_ARGB || L == L_ABGR) ? 0 : -1;
};

// PRMT selector that turns a normalised source pixel word into a destination-ordered one
// (selector index 4 = byte 0 of the second PRMT operand, which is zero).
template <int SL, int DL>
__host__ __device__ constexpr uint32_t perm_selector()
{
    using S = RgbPos<SL>;
    using D = RgbPos<DL>;
    uint32_t sel = 0;
    for (int k = 0; k < 4; k++) {
        int idx = 4;
        if (k == D::r) idx = S::r;
        else if (k == D::g) idx = S::g;
        else if (k == D::b) idx = S::b;
        else if (k == D::a) idx = S::a >= 0 ? S::a : 4;
        sel |= (uint32_t)idx << (4 * k);
    }
    return sel;
}

template <class OP>
__global__ void __launch_bounds__(256, 4) k_rgb_linear(FastParams p)
{
    extern __shared__ uint4 s_stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage = s_stage + warp * 32 * OP::kStage;
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < p.nunits; base += stride) {
        const uint32_t warp_u0 = base + warp * 32;
        if (warp_u0 >= p.nunits) break;
        const uint32_t u = base + threadIdx.x;
        OP::run(p, soff, doff, u, u < p.nunits, warp_u0, (int)min(32u, p.nunits - warp_u0), stage, lane);
    }
}

template <class OP>
bool launch_rgb(const FastParams &p, int nframes, cudaStream_t st, const char *name)
{
    const LaunchShape s = shape_linear(p.nunits, nframes);
    const size_t smem = (size_t)(s.block.x / 32) * 32 * OP::kStage * sizeof(uint4);
    k_rgb_linear<OP><<<s.grid, s.block, smem, st>>>(p);
    note_launch();
    ACGPU_CHECK_LAUNCH(name);
    return true;
}

// 16 destination-ordered pixel words -> 12 words of 24-bit pixels
__device__ __forceinline__ void pack24(const uint32_t *d, uint32_t *ow)
{
#pragma unroll
    for (int g = 0; g < 4; g++) {
        ow[3 * g + 0] = __byte_perm(d[4 * g + 0], d[4 * g + 1], 0x4210);
        ow[3 * g + 1] = __byte_perm(d[4 * g + 1], d[4 * g + 2], 0x5421);
        ow[3 * g + 2] = __byte_perm(d[4 * g + 2], d[4 * g + 3], 0x6542);
    }
}

// RGB (any layout) -> RGB (any other layout) where at least one side is 24-bit
template <int SL, int DL>
struct RgbToRgb {
    static constexpr int kStage = RgbPos<DL>::bpp;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t warp_u0, int nvalid, uint4 *stage, int lane)
    {
        constexpr uint32_t SEL = perm_selector<SL, DL>();
        uint32_t px[16];
        load_rgb16<SL>(p.s0 + soff, u, valid, px);
#pragma unroll
        for (int k = 0; k < 16; k++) px[k] = __byte_perm(px[k], 0, SEL);
        if (RgbPos<DL>::bpp == 4) {
            store_chunks<4>(stage, lane, px, p.d0 + doff + (size_t)warp_u0 * 64, nvalid);
        } else {
            uint32_t ow[12];
            pack24(px, ow);
            store_chunks<3>(stage, lane, ow, p.d0 + doff + (size_t)warp_u0 * 48, nvalid);
        }
    }
};

// RGB -> GRAY8: (19595 R + 38470 G + 7471 B + 32768) >> 16      img_rgb_packed.c:179-303
template <int SL>
struct RgbToGray {
    static constexpr int kStage = 0;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t, int, uint4 *, int)
    {
        using RI = RgbInfo<SL>;
        constexpr uint32_t lo = RI::half(19595, 38470, 7471, 0), hi = RI::half(19595, 38470, 7471, 2);
        uint32_t px[16], a[16];
        load_rgb16<SL>(p.s0 + soff, u, valid, px);
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = dp2a_hi_uu(hi, px[k], dp2a_lo_uu(lo, px[k], 32768u));
        if (valid)
            stg128(p.d0 + doff + (size_t)u * 16,
                   make_uint4(pack_b2x4(a[0], a[1], a[2], a[3]), pack_b2x4(a[4], a[5], a[6], a[7]),
                              pack_b2x4(a[8], a[9], a[10], a[11]), pack_b2x4(a[12], a[13], a[14], a[15])));
    }
};

// GRAY8 -> RGB (alpha written as 0, img_rgb_packed.c:307-340) and Y8 -> RGB (luma range-mapped first, alpha
// left untouched, img_yuv_rgb.c:354-379).
template <bool FROM_Y8, int DL>
struct GrayToRgb {
    static constexpr int kStage = RgbPos<DL>::bpp;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t warp_u0, int nvalid, uint4 *stage, int lane)
    {
        using D = RgbPos<DL>;
        if (FROM_Y8 && D::bpp == 4) {       // destination is read-modify-written: ask L2 for it now
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = j * 32 + lane;
                if (c < nvalid * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d0 + doff + (size_t)warp_u0 * 64 + (size_t)c * 16));
            }
        }
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid) v = ldg128(p.s0 + soff + (size_t)u * 16);
        uint32_t gw[4] = {v.x, v.y, v.z, v.w};
        if (FROM_Y8) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t x[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    x[k] = (uint32_t)__vimin_s32_relu((int)__dp4a(gw[i], 1u << (8 * k), (uint32_t)-16), 219) * pixmath::kY2GrayMul;
                gw[i] = pack_top4(x[0], x[1], x[2], x[3]);
            }
        }
        if (D::bpp == 3) {
            uint32_t ow[12];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                ow[3 * g + 0] = __byte_perm(gw[g], 0, 0x1000);
                ow[3 * g + 1] = __byte_perm(gw[g], 0, 0x2211);
                ow[3 * g + 2] = __byte_perm(gw[g], 0, 0x3332);
            }
            store_chunks<3>(stage, lane, ow, p.d0 + doff + (size_t)warp_u0 * 48, nvalid);
        } else {
            uint32_t ow[16];
#pragma unroll
            for (int g = 0; g < 4; g++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    ow[4 * g + k] = D::a == 3 ? __byte_perm(gw[g], 0, 0x4000u | (uint32_t)(k * 0x111))
                                              : __byte_perm(gw[g], 0, 0x0004u | (uint32_t)(k * 0x1110));
            if (FROM_Y8) store_row_rgb<4, D::a == 0>(stage, lane, ow, p.d0 + doff + (size_t)warp_u0 * 64, nvalid);
            else store_chunks<4>(stage, lane, ow, p.d0 + doff + (size_t)warp_u0 * 64, nvalid);
        }
    }
};

// -K on RGB24 (src/video_trans.c:381-388): gray = (19595 R + 38470 G + 7471 B + 32768) >> 16 written to all three
// channels.  One pass, in place: a warp reads its own 1536-byte tile completely before it writes it back.
struct DecolorRgb24 {
    static constexpr int kStage = 3;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t warp_u0, int nvalid, uint4 *stage, int lane)
    {
        using RI = RgbInfo<L_RGB24>;
        constexpr uint32_t lo = RI::half(19595, 38470, 7471, 0), hi = RI::half(19595, 38470, 7471, 2);
        uint32_t px[16], ow[12];
        {
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 3; k++) {     // plain loads: the source is also the destination
                uint4 v = make_uint4(0, 0, 0, 0);
                if (valid) v = *reinterpret_cast<const uint4 *>(p.s0 + soff + (size_t)u * 48 + k * 16);
                w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {
                px[4 * g + 0] = w[3 * g];
                px[4 * g + 1] = __byte_perm(w[3 * g], w[3 * g + 1], 0x0543);
                px[4 * g + 2] = __byte_perm(w[3 * g + 1], w[3 * g + 2], 0x0432);
                px[4 * g + 3] = w[3 * g + 2] >> 8;
            }
        }
        __syncwarp();      // every lane has its pixels before any lane's tile bytes are overwritten
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t a[4];
#pragma unroll
            for (int k = 0; k < 4; k++) a[k] = dp2a_hi_uu(hi, px[4 * g + k], dp2a_lo_uu(lo, px[4 * g + k], 32768u));
            const uint32_t gw = pack_b2x4(a[0], a[1], a[2], a[3]);
            ow[3 * g + 0] = __byte_perm(gw, 0, 0x1000);
            ow[3 * g + 1] = __byte_perm(gw, 0, 0x2211);
            ow[3 * g + 2] = __byte_perm(gw, 0, 0x3332);
        }
        store_chunks<3>(stage, lane, ow, p.d0 + doff + (size_t)warp_u0 * 48, nvalid);
    }
};

int layout_of(int fmt)
{
    switch (fmt) {
    case IMG_RGB24:  return L_RGB24;
    case IMG_BGR24:  return L_BGR24;
    case IMG_RGBA32: return L_RGBA;
    case IMG_BGRA32: return L_BGRA;
    case IMG_ARGB32: return L_ARGB;
    case IMG_ABGR32: return L_ABGR;
    default: return -1;
    }
}

uint32_t runtime_selector(int sl, int dl)
{
    static const int pos[6][4] = {  // r, g, b, a (-1 none)
        {0, 1, 2, -1}, {2, 1, 0, -1}, {0, 1, 2, 3}, {2, 1, 0, 3}, {1, 2, 3, 0}, {3, 2, 1, 0}};
    uint32_t sel = 0;
    for (int k = 0; k < 4; k++) {
        int idx = 4;
        for (int c = 0; c < 4; c++)
            if (pos[dl][c] == k) idx = pos[sl][c] >= 0 ? pos[sl][c] : 4;
        sel |= (uint32_t)idx << (4 * k);
    }
    return sel;
}

template <int SL>
bool dispatch_rgb_dst(int dl, const FastParams &p, int nf, cudaStream_t st)
{
    switch (dl) {
    case L_RGB24: return launch_rgb<RgbToRgb<SL, L_RGB24>>(p, nf, st, "rgb->rgb24");
    case L_BGR24: return launch_rgb<RgbToRgb<SL, L_BGR24>>(p, nf, st, "rgb->bgr24");
    case L_RGBA:  return launch_rgb<RgbToRgb<SL, L_RGBA>>(p, nf, st, "rgb->rgba32");
    case L_BGRA:  return launch_rgb<RgbToRgb<SL, L_BGRA>>(p, nf, st, "rgb->bgra32");
    case L_ARGB:  return launch_rgb<RgbToRgb<SL, L_ARGB>>(p, nf, st, "rgb->argb32");
    case L_ABGR:  return launch_rgb<RgbToRgb<SL, L_ABGR>>(p, nf, st, "rgb->abgr32");
    default: return false;
    }
}

template <bool FROM_Y8>
bool dispatch_gray_dst(int dl, const FastParams &p, int nf, cudaStream_t st)
{
    switch (dl) {
    case L_RGB24: return launch_rgb<GrayToRgb<FROM_Y8, L_RGB24>>(p, nf, st, "gray->rgb24");
    case L_BGR24: return launch_rgb<GrayToRgb<FROM_Y8, L_BGR24>>(p, nf, st, "gray->bgr24");
    case L_RGBA:  return launch_rgb<GrayToRgb<FROM_Y8, L_RGBA>>(p, nf, st, "gray->rgba32");
    case L_BGRA:  return launch_rgb<GrayToRgb<FROM_Y8, L_BGRA>>(p, nf, st, "gray->bgra32");
    case L_ARGB:  return launch_rgb<GrayToRgb<FROM_Y8, L_ARGB>>(p, nf, st, "gray->argb32");
    case L_ABGR:  return launch_rgb<GrayToRgb<FROM_Y8, L_ABGR>>(p, nf, st, "gray->abgr32");
    default: return false;
    }
}

}  // namespace

bool decolor_rgb24_fast(uint8_t *frames, size_t pitch, int w, int h, int nframes, cudaStream_t st)
{
    const size_t P = (size_t)w * h;
    if (!al16(frames) || (nframes > 1 && pitch % 16) || P % 16 || P / 16 > 0x7FFFFFFFu) return false;
    FastParams p{};
    p.s0 = frames; p.d0 = frames; p.spitch = p.dpitch = pitch;
    p.w = w; p.h = h; p.nunits = (uint32_t)(P / 16);
    return launch_rgb<DecolorRgb24>(p, nframes, st, "decolor_rgb24");
}

bool fast_rgb_family(const ConvertArgs &a, const fast::FastParams &p)
{
    const int sl = layout_of(a.srcfmt), dl = layout_of(a.dstfmt);
    const int nf = a.nframes;
    cudaStream_t st = a.stream;
    const size_t P = (size_t)a.w * a.h;
    if (sl >= 0 && dl >= 0) {
        const int sb = sl <= L_BGR24 ? 3 : 4, db = dl <= L_BGR24 ? 3 : 4;
        if (sl == dl)          // rgb_copy / rgba_copy
            return launch_wordperm(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, 0x3210, P * sb, nf, st);
        if (sb == 4 && db == 4)   // rgba_swapall / swap02 / swap13 / alpha30 / alpha03: one PRMT per pixel
            return launch_wordperm(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, runtime_selector(sl, dl), P * 4, nf, st);
        switch (sl) {
        case L_RGB24: return dispatch_rgb_dst<L_RGB24>(dl, p, nf, st);
        case L_BGR24: return dispatch_rgb_dst<L_BGR24>(dl, p, nf, st);
        case L_RGBA:  return dispatch_rgb_dst<L_RGBA>(dl, p, nf, st);
        case L_BGRA:  return dispatch_rgb_dst<L_BGRA>(dl, p, nf, st);
        case L_ARGB:  return dispatch_rgb_dst<L_ARGB>(dl, p, nf, st);
        default:      return dispatch_rgb_dst<L_ABGR>(dl, p, nf, st);
        }
    }
    if (sl >= 0 && a.dstfmt == IMG_GRAY8) {
        switch (sl) {
        case L_RGB24: return launch_rgb<RgbToGray<L_RGB24>>(p, nf, st, "rgb24->gray8");
        case L_BGR24: return launch_rgb<RgbToGray<L_BGR24>>(p, nf, st, "bgr24->gray8");
        case L_RGBA:  return launch_rgb<RgbToGray<L_RGBA>>(p, nf, st, "rgba32->gray8");
        case L_BGRA:  return launch_rgb<RgbToGray<L_BGRA>>(p, nf, st, "bgra32->gray8");
        case L_ARGB:  return launch_rgb<RgbToGray<L_ARGB>>(p, nf, st, "argb32->gray8");
        default:      return launch_rgb<RgbToGray<L_ABGR>>(p, nf, st, "abgr32->gray8");
        }
    }
    if (dl >= 0 && a.srcfmt == IMG_GRAY8) return dispatch_gray_dst<false>(dl, p, nf, st);
    if (dl >= 0 && a.srcfmt == IMG_Y8) return dispatch_gray_dst<true>(dl, p, nf, st);
    return false;
}

}  // namespace acgpu
