// fast_common.cuh -- helpers shared by the vectorised-tier translation units (kernels_fast*.cu).
#pragma once

#include "acgpu_internal.h"
#include "pixmath.cuh"

#include <stdlib.h>

namespace acgpu {
namespace fast {

// ---------------------------------------------------------------------------------------------------
// Small device helpers

__device__ __forceinline__ uint4 ldg128(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint2 ldg64(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ uint32_t ldg32(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
// streaming stores: the output is never re-read by this kernel
__device__ __forceinline__ void stg128(uint8_t *p, uint4 v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
__device__ __forceinline__ void stg64(uint8_t *p, uint2 v) { __stcs(reinterpret_cast<uint2 *>(p), v); }
__device__ __forceinline__ void stg32(uint8_t *p, uint32_t v) { __stcs(reinterpret_cast<uint32_t *>(p), v); }

// ---- byte-aligned chroma access for ragged 4:2:0 rows (width % 16 != 0) ------------------------------------------
// nb (1..8) bytes from an address with any alignment, returned in the low bytes.  Only the aligned 32-bit words that
// hold requested bytes are touched, so the access never leaves the plane; bytes past nb are unspecified.
__device__ __forceinline__ uint2 ldg_bytes8(const uint8_t *p, uint32_t nb)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t s = (uint32_t)(a & 3), sh = s * 8;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t w0 = __ldg(q), w1 = (4 - s < nb) ? __ldg(q + 1) : 0u, w2 = (8 - s < nb) ? __ldg(q + 2) : 0u;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
// nb (1 .. 4*NW) bytes from an address with any alignment into NW words (low bytes first); same word-granular rule.
template <int NW>
__device__ __forceinline__ void ldg_bytes(const uint8_t *p, uint32_t nb, uint32_t *out)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if (NW % 4 == 0 && nb == 4u * NW && (a & 15) == 0) {          // whole and aligned after all (uniform along a row)
#pragma unroll
        for (int i = 0; i < NW / 4; i++) {
            const uint4 v = ldg128(p + 16 * i);
            out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
        }
        return;
    }
    if (NW % 2 == 0 && nb == 4u * NW && (a & 7) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 2; i++) {
            const uint2 v = ldg64(p + 8 * i);
            out[2 * i] = v.x; out[2 * i + 1] = v.y;
        }
        return;
    }
    const uint32_t s = (uint32_t)(a & 3), sh = s * 8;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = (4u * i < nb + s) ? __ldg(q + i) : 0u;
#pragma unroll
    for (int i = 0; i < NW; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
}
__device__ __forceinline__ uint64_t u64_of(uint2 v) { return (uint64_t)v.x | ((uint64_t)v.y << 32); }
__device__ __forceinline__ uint2 uint2_of(uint64_t v) { return make_uint2((uint32_t)v, (uint32_t)(v >> 32)); }

// The 8 chroma samples of the 16 flat pixels that start at column x0 (even) of row y in a 4:2:0 frame of width w.
// A unit that runs over the end of its row (at most one per row, w >= 16) takes its first k samples from chroma row
// y/2 and the rest from the start of chroma row (y+1)/2.
__device__ __forceinline__ uint2 gather420(const uint8_t *plane, uint32_t y, uint32_t x0, uint32_t w)
{
    const uint32_t cw = w >> 1, k = min(8u, (w - x0) >> 1);
    uint2 v = ldg_bytes8(plane + (size_t)(y >> 1) * cw + (x0 >> 1), k);
    if (k < 8) {
        const uint2 t = ldg_bytes8(plane + (size_t)((y + 1) >> 1) * cw, 8 - k);
        v = uint2_of((u64_of(v) & ((1ull << (8 * k)) - 1)) | (u64_of(t) << (8 * k)));
    }
    return v;
}

// 8 bytes to an address with any alignment, in the widest pieces the address allows (uniform along a row).
__device__ __forceinline__ void stg64_unaligned(uint8_t *p, uint2 v)
{
    const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(p) & 7u;
    if (a == 0) {
        stg64(p, v);
    } else if (a == 4) {
        stg32(p, v.x);
        stg32(p + 4, v.y);
    } else if (!(a & 1)) {              // 2 mod 4: 2 + 4 + 2 bytes
        *reinterpret_cast<uint16_t *>(p) = (uint16_t)v.x;
        stg32(p + 2, __funnelshift_r(v.x, v.y, 16));
        *reinterpret_cast<uint16_t *>(p + 6) = (uint16_t)(v.y >> 16);
    } else if ((a & 3) == 1) {          // 1 mod 4: 1 + 2 + 4 + 1 bytes
        p[0] = (uint8_t)v.x;
        *reinterpret_cast<uint16_t *>(p + 1) = (uint16_t)(v.x >> 8);
        stg32(p + 3, __funnelshift_r(v.x, v.y, 24));
        p[7] = (uint8_t)(v.y >> 24);
    } else {                            // 3 mod 4: 1 + 4 + 2 + 1 bytes
        p[0] = (uint8_t)v.x;
        stg32(p + 1, __funnelshift_r(v.x, v.y, 8));
        *reinterpret_cast<uint16_t *>(p + 5) = (uint16_t)(v.y >> 8);
        p[7] = (uint8_t)(v.y >> 24);
    }
}
// bytes j0 .. j0+nb-1 of v to p[0 .. nb)
__device__ __forceinline__ void stg_bytes8(uint8_t *p, uint2 v, uint32_t j0, uint32_t nb)
{
    if (nb == 8) {
        stg64_unaligned(p, v);
    } else {
        const uint64_t x = u64_of(v) >> (8 * j0);
#pragma unroll
        for (uint32_t i = 0; i < 7; i++)
            if (i < nb) p[i] = (uint8_t)(x >> (8 * i));
    }
}

__device__ __forceinline__ uint32_t word_of(const uint4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int b) { return (w >> (8 * b)) & 0xFFu; }

// One colour channel: j = 16*Y + c (dp4a picks the Y byte and scales it), clamp j to [0, 3498], and the
// answer is the TOP byte of j*1220944 + 2^23 (pixmath::ylut_word_fast).  3 instructions.
// BIAS: a constant the caller left inside c (biased table entries); it is removed inside the clamp instruction.
template <int BIAS>
__device__ __forceinline__ uint32_t channel_word(uint32_t yword, uint32_t ysel, int c)
{
    const int j = (int)__dp4a(yword, ysel, (uint32_t)c);
    const int jc = BIAS ? __viaddmin_s32_relu(j, -BIAS, pixmath::kJMax) : __vimin_s32_relu(j, pixmath::kJMax);
    return (uint32_t)jc * pixmath::kJMul + pixmath::kJAdd;
}
// the last two steps of channel_word on a ready j
template <int BIAS>
__device__ __forceinline__ uint32_t clamp_scale(int j)
{
    const int jc = BIAS ? __viaddmin_s32_relu(j, -BIAS, pixmath::kJMax) : __vimin_s32_relu(j, pixmath::kJMax);
    return (uint32_t)jc * pixmath::kJMul + pixmath::kJAdd;
}
// (a.b3, b.b3, c.b3, d.b3) -> one word
__device__ __forceinline__ uint32_t pack_top4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
}

// Per-warp staging: lane-owned chunks (each lane owns K consecutive 16-byte chunks of the warp's contiguous
// output) are written to shared memory and read back in global order so every STG covers 512 contiguous bytes.
// K = 3 needs no swizzle (48-byte lane stride is conflict-free); K = 4 / K = 2 XOR-swizzle the chunk slot.
template <int K>
__device__ __forceinline__ int stage_slot(int lane, int k)
{
    if (K == 4) return lane * 4 + (k ^ ((lane >> 1) & 3));
    if (K == 2) return lane * 2 + (k ^ ((lane >> 2) & 1));
    return lane * K + k;
}
template <int K>
__device__ __forceinline__ int stage_slot_linear(int c)   // c = global chunk index inside the warp tile
{
    return stage_slot<K>(c / K, c % K);
}

// S420R: 4:2:0 whose width is not a multiple of 16 ("ragged"): walked as flat 16-pixel units like the other layouts,
// luma and RGB stay 16-byte aligned, only the chroma rows are reached through byte-aligned accesses.
enum SrcKind { S420 = 0, S422 = 1, S411 = 2, S444 = 3, SYUY2 = 4, SUYVY = 5, SYVYU = 6, S420R = 7 };

struct FastParams {
    const uint8_t *s0, *s1, *s2;
    uint8_t *d0, *d1, *d2;
    size_t spitch, dpitch;
    int w, h;
    int upr;           // 4:2:0 mode: 16-pixel units per row
    int nrp;           // 4:2:0 mode: row pairs
    uint32_t nunits;   // linear mode: units per frame
    int flat420;       // 4:2:0 YUV->RGB: warps are packed across row-pair boundaries (no idle lanes when w/16 % 32 != 0)
    int ragged420;     // 4:2:0 with width % 16 != 0: S420R / D420R kernels (flat units, byte-aligned chroma rows)
};

// (a.b2, b.b2, c.b2, d.b2) -> one word
__device__ __forceinline__ uint32_t pack_b2x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
}

// Stores K lane-owned 16-byte chunks (words ow[0..4K)) through the staging buffer; plain coalesced write.
template <int K>
__device__ __forceinline__ void store_chunks(uint4 *stage, int lane, const uint32_t *ow, uint8_t *warpbase, int nvalid)
{
#pragma unroll
    for (int k = 0; k < K; k++)
        stage[stage_slot<K>(lane, k)] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
    __syncwarp();
    const int nchunks = nvalid * K;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int c = j * 32 + lane;
        if (c < nchunks) stg128(warpbase + (size_t)c * 16, stage[stage_slot_linear<K>(c)]);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// launch helpers

// How many "waves" of resident blocks a launch is cut into.  Measured on B200 (tools/membench, profiles/): a grid of
// exactly-resident persistent blocks loses ~10 % of HBM bandwidth to static load imbalance between SMs (two dies,
// near/far L2); the block scheduler evens that out when there are many more blocks than slots.  Kernels without a
// per-block prologue want ~32 waves, the YUV->RGB kernels (2 KB table fill per block) peak at ~8.
inline int waves(int dflt)
{
    static int v = -2;
    if (v == -2) { const char *e = getenv("ACGPU_WAVES"); v = e ? atoi(e) : -1; }
    return v > 0 ? v : dflt;
}

inline bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

struct LaunchShape {
    dim3 grid, block;
};

// 4:2:0 mode: one block spans a row pair (ceil(upr/32) warps), grid.x strides over row pairs.
inline LaunchShape shape_420(int upr, int nrp, int nframes, int nwaves = 32)
{
    // block = up to 256 threads along the row; wider rows (> 4096 pixels) are cut into column segments (grid.z)
    LaunchShape s;
    const int threads = upr >= 256 ? 256 : ((upr + 31) / 32) * 32;
    const int segs = (upr + threads - 1) / threads;
    s.block = dim3(threads);
    const int per_sm = 2048 / threads;
    long want = (long)sm_count() * per_sm * waves(nwaves);
    long gx = (want + (long)nframes * segs - 1) / ((long)nframes * segs);
    if (gx < 1) gx = 1;
    if (gx > nrp) gx = nrp;
    s.grid = dim3((unsigned)gx, (unsigned)nframes, (unsigned)segs);
    return s;
}

inline LaunchShape shape_linear(uint32_t nunits, int nframes, int nwaves = 32)
{
    LaunchShape s;
    s.block = dim3(256);
    long want = (long)sm_count() * 8 * waves(nwaves);
    long gx = (want + nframes - 1) / nframes;
    const long maxgx = (nunits + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > maxgx) gx = maxgx;
    s.grid = dim3((unsigned)gx, (unsigned)nframes);
    return s;
}


// Writes one row's 16 pixels per lane (BPP*4 words) through the warp staging buffer to `rowbase`
// (the address of the warp's first unit).  nvalid = number of lanes holding real units (warp-uniform).
// Two-segment form: the first `ksplit` lanes' pixels continue at rowbase, the remaining lanes' pixels start at
// `rowbase2` (a warp that straddles the end of an image row); the one-segment form passes ksplit = 32.
template <int BPP, bool AFIRST>
__device__ __forceinline__ void store_row_rgb2(uint4 *stage, int lane, const uint32_t *ow, uint8_t *rowbase, uint8_t *rowbase2,
                                               int ksplit, int nvalid)
{
    constexpr int K = BPP;   // 16-byte chunks per lane
#pragma unroll
    for (int k = 0; k < K; k++)
        stage[stage_slot<K>(lane, k)] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
    __syncwarp();
    const int nchunks = nvalid * K, csplit = ksplit * K;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int c = j * 32 + lane;
        if (c < nchunks) {
            uint4 v = stage[stage_slot_linear<K>(c)];
            uint8_t *g = c < csplit ? rowbase + (size_t)c * 16 : rowbase2 + (size_t)(c - csplit) * 16;
            if (BPP == 4) {     // keep the destination's alpha bytes (img_yuv_rgb.c:62-64 never stores them)
                const uint4 old = *reinterpret_cast<const uint4 *>(g);
                const uint32_t m = AFIRST ? 0x000000FFu : 0xFF000000u;
                v.x = (v.x & ~m) | (old.x & m);
                v.y = (v.y & ~m) | (old.y & m);
                v.z = (v.z & ~m) | (old.z & m);
                v.w = (v.w & ~m) | (old.w & m);
            }
            stg128(g, v);
        }
    }
    __syncwarp();
}

template <int BPP, bool AFIRST>
__device__ __forceinline__ void store_row_rgb(uint4 *stage, int lane, const uint32_t *ow, uint8_t *rowbase, int nvalid)
{
    store_row_rgb2<BPP, AFIRST>(stage, lane, ow, rowbase, rowbase, 32, nvalid);
}

enum RgbLayout { L_RGB24 = 0, L_BGR24 = 1, L_RGBA = 2, L_BGRA = 3, L_ARGB = 4, L_ABGR = 5 };
enum YuvDst { D420 = 0, D422 = 1, D411 = 2, D444 = 3, DYUY2 = 4, DUYVY = 5, DYVYU = 6, DY8 = 7, D420R = 8 };

template <int SL> struct RgbInfo {
    static constexpr int bpp = SL <= L_BGR24 ? 3 : 4;
    // byte position of each channel inside the (normalised) pixel word
    static constexpr int rpos = (SL == L_RGB24 || SL == L_RGBA) ? 0 : (SL == L_BGR24 || SL == L_BGRA) ? 2 : SL == L_ARGB ? 1 : 3;
    static constexpr int gpos = (SL == L_ARGB || SL == L_ABGR) ? 2 : 1;
    static constexpr int bpos = (SL == L_RGB24 || SL == L_RGBA) ? 2 : (SL == L_BGR24 || SL == L_BGRA) ? 0 : SL == L_ARGB ? 3 : 1;
    static constexpr uint32_t half(int kr, int kg, int kb, int p0)   // coefficients for byte positions p0, p0+1
    {
        const int c0 = rpos == p0 ? kr : gpos == p0 ? kg : bpos == p0 ? kb : 0;
        const int c1 = rpos == p0 + 1 ? kr : gpos == p0 + 1 ? kg : bpos == p0 + 1 ? kb : 0;
        return (uint32_t)(c0 & 0xFFFF) | ((uint32_t)(c1 & 0xFFFF) << 16);
    }
};

__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_lo_su(uint32_t a, uint32_t b, uint32_t c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((int)c));
    return (uint32_t)d;
}
__device__ __forceinline__ uint32_t dp2a_hi_su(uint32_t a, uint32_t b, uint32_t c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((int)c));
    return (uint32_t)d;
}

// Loads the 16 pixels of unit `u` of a row that starts at `row` into normalised pixel words.
template <int SL>
__device__ __forceinline__ void load_rgb16(const uint8_t *row, uint32_t u, bool valid, uint32_t *px)
{
    if (RgbInfo<SL>::bpp == 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (valid) v = ldg128(row + (size_t)u * 64 + k * 16);
            px[4 * k] = v.x; px[4 * k + 1] = v.y; px[4 * k + 2] = v.z; px[4 * k + 3] = v.w;
        }
    } else {
        uint32_t w[12];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (valid) v = ldg128(row + (size_t)u * 48 + k * 16);
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {   // 4 pixels = 3 words
            px[4 * g + 0] = w[3 * g];
            px[4 * g + 1] = __byte_perm(w[3 * g], w[3 * g + 1], 0x0543);
            px[4 * g + 2] = __byte_perm(w[3 * g + 1], w[3 * g + 2], 0x0432);
            px[4 * g + 3] = w[3 * g + 2] >> 8;
        }
    }
}

// per-byte rounded-up / truncated means on four packed bytes (aclib (a+b+1)/2 and (a+b)/2)
__device__ __forceinline__ uint32_t avg_up4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu); }
__device__ __forceinline__ uint32_t avg_dn4(uint32_t a, uint32_t b) { return (a & b) + (((a ^ b) >> 1) & 0x7F7F7F7Fu); }

}  // namespace fast

// family entry points (each returns false without launching when the pair is not one of its own)
bool fast_yuv_family(const ConvertArgs &a, const fast::FastParams &p);   // kernels_fast_yuv.cu: YUV<->YUV, Y8, gray maps
bool fast_rgb_family(const ConvertArgs &a, const fast::FastParams &p);   // kernels_fast_rgb.cu: RGB<->RGB, gray<->RGB, Y8->RGB

}  // namespace acgpu
