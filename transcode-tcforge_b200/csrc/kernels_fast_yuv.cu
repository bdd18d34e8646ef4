// kernels_fast_yuv.cu -- vectorised tier, YUV<->YUV family: planar chroma resampling (img_yuv_planar.c),
// packed permutes (img_yuv_packed.c), planar<->packed (img_yuv_mixed.c) and the luma/gray range maps
// (img_yuv_rgb.c:254-348).  Pure byte shuffling plus per-byte means, all done on packed 32-bit words with
// PRMT; every kernel reads and writes each byte exactly once with 16/8/4-byte coalesced accesses.
#include "fast_common.cuh"

namespace acgpu {
namespace {

using namespace fast;

enum { MODE_LINEAR = 0, MODE_ROWPAIR = 1 };
enum PlanarKind { P420 = 0, P422 = 1, P411 = 2, P444 = 3, PY8 = 4, PGRAY = 5 };   // "planar-like" operands
enum PackedKind { QYUY2 = 0, QUYVY = 1, QYVYU = 2 };

// ---- generic drivers: OP::run is called once per unit (16 pixels; two rows of them in row-pair mode) ----
template <class OP>
__global__ void __launch_bounds__(256, 4) k_linear(FastParams p)
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
__global__ void __launch_bounds__(256, 4) k_rowpair(FastParams p)
{
    extern __shared__ uint4 s_stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage = s_stage + warp * 32 * OP::kStage;
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    const int unit = blockIdx.z * blockDim.x + threadIdx.x;              // blockIdx.z: column segment of wide rows
    const int wu0 = blockIdx.z * blockDim.x + warp * 32;                 // first unit of this warp
    const int nvalid = min(32, p.upr - wu0);
    if (nvalid <= 0) return;
    const bool valid = unit < p.upr;
    for (int rp = blockIdx.x; rp < p.nrp; rp += gridDim.x)
        OP::run(p, soff, doff, rp, unit, valid, wu0, nvalid, stage, lane);
}

template <class OP>
bool launch_op(const FastParams &p, int nframes, cudaStream_t st, const char *name)
{
    const LaunchShape s = OP::kMode == MODE_ROWPAIR ? shape_420(p.upr, p.nrp, nframes) : shape_linear(p.nunits, nframes);
    const size_t smem = (size_t)(s.block.x / 32) * 32 * OP::kStage * sizeof(uint4);
    if constexpr (OP::kMode == MODE_ROWPAIR) k_rowpair<OP><<<s.grid, s.block, smem, st>>>(p);
    else k_linear<OP><<<s.grid, s.block, smem, st>>>(p);
    note_launch();
    ACGPU_CHECK_LAUNCH(name);
    return true;
}

// ---- word-level helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const uint8_t *p, bool valid, uint32_t *w)
{
    uint4 v = make_uint4(0, 0, 0, 0);
    if (valid) v = ldg128(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
}
__device__ __forceinline__ void ld2(const uint8_t *p, bool valid, uint32_t *w)
{
    uint2 v = make_uint2(0, 0);
    if (valid) v = ldg64(p);
    w[0] = v.x; w[1] = v.y;
}
__device__ __forceinline__ uint32_t ld1(const uint8_t *p, bool valid) { return valid ? ldg32(p) : 0u; }
__device__ __forceinline__ void st4(uint8_t *p, bool valid, const uint32_t *w) { if (valid) stg128(p, make_uint4(w[0], w[1], w[2], w[3])); }
__device__ __forceinline__ void st2(uint8_t *p, bool valid, const uint32_t *w) { if (valid) stg64(p, make_uint2(w[0], w[1])); }
__device__ __forceinline__ void st1(uint8_t *p, bool valid, uint32_t w) { if (valid) stg32(p, w); }

// luma range maps on four packed bytes (pixmath::y2gray_word_fast / gray2y_word_fast)
__device__ __forceinline__ uint32_t map_y2gray4(uint32_t w)
{
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        x[k] = (uint32_t)__vimin_s32_relu((int)__dp4a(w, 1u << (8 * k), (uint32_t)-16), 219) * pixmath::kY2GrayMul;
    return pack_top4(x[0], x[1], x[2], x[3]);
}
__device__ __forceinline__ uint32_t map_gray2y4(uint32_t w)
{
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = __dp4a(w, 1u << (8 * k), 0u) * pixmath::kGray2YMul + (16u << 24);
    return pack_top4(x[0], x[1], x[2], x[3]);
}

// chroma bytes owned by one 16-pixel unit of a row
__host__ __device__ constexpr int chroma_bytes(int pk) { return pk == P444 ? 16 : pk == P411 ? 4 : (pk == P420 || pk == P422) ? 8 : 0; }

template <int N> __device__ __forceinline__ void ldn(const uint8_t *p, bool valid, uint32_t *w)
{
    if (N == 16) ld4(p, valid, w);
    else if (N == 8) ld2(p, valid, w);
    else if (N == 4) w[0] = ld1(p, valid);
}
template <int N> __device__ __forceinline__ void stn(uint8_t *p, bool valid, const uint32_t *w)
{
    if (N == 16) st4(p, valid, w);
    else if (N == 8) st2(p, valid, w);
    else if (N == 4) st1(p, valid, w[0]);
}

// Converts the chroma bytes of one unit-row between horizontal subsamplings (no vertical step here).
//   NS source bytes -> ND destination bytes, NS,ND in {4 (4:1:1), 8 (4:2:x), 16 (4:4:4)}
template <int NS, int ND>
__device__ __forceinline__ void chroma_h(const uint32_t *s, uint32_t *d)
{
    if (NS == ND) {
#pragma unroll
        for (int i = 0; i < NS / 4; i++) d[i] = s[i];
    } else if (NS == 8 && ND == 16) {           // replicate x2      img_yuv_planar.c:198-211
        d[0] = __byte_perm(s[0], 0, 0x1100); d[1] = __byte_perm(s[0], 0, 0x3322);
        d[2] = __byte_perm(s[1], 0, 0x1100); d[3] = __byte_perm(s[1], 0, 0x3322);
    } else if (NS == 4 && ND == 8) {            // replicate x2      :133-146
        d[0] = __byte_perm(s[0], 0, 0x1100); d[1] = __byte_perm(s[0], 0, 0x3322);
    } else if (NS == 4 && ND == 16) {           // replicate x4      :148-164
        d[0] = __byte_perm(s[0], 0, 0x0000); d[1] = __byte_perm(s[0], 0, 0x1111);
        d[2] = __byte_perm(s[0], 0, 0x2222); d[3] = __byte_perm(s[0], 0, 0x3333);
    } else if (NS == 16 && ND == 8) {           // (a+b+1)/2 pairs   :253-266
        d[0] = avg_up4(__byte_perm(s[0], s[1], 0x6420), __byte_perm(s[0], s[1], 0x7531));
        d[1] = avg_up4(__byte_perm(s[2], s[3], 0x6420), __byte_perm(s[2], s[3], 0x7531));
    } else if (NS == 8 && ND == 4) {            // (a+b+1)/2 pairs   :183-196, :66-81
        d[0] = avg_up4(__byte_perm(s[0], s[1], 0x6420), __byte_perm(s[0], s[1], 0x7531));
    } else if (NS == 16 && ND == 4) {           // (a+b+c+d+2)/4     :234-251
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = __dp4a(s[i], 0x01010101u, 2u) >> 2;
        d[0] = __byte_perm(__byte_perm(r[0], r[1], 0x0040), __byte_perm(r[2], r[3], 0x0040), 0x5410);
    }
}

// ---- planar-like -> planar-like (incl. Y8 / GRAY8 operands and the luma range maps) -------------------------
// Linear mode: neither side is 4:2:0.  SP/DP in PlanarKind.
template <int SP, int DP>
struct PlanarLinear {
    static constexpr int kMode = MODE_LINEAR, kStage = 0;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t, int, uint4 *, int)
    {
        uint32_t y[4];
        ld4(p.s0 + soff + (size_t)u * 16, valid, y);
        if (SP == PGRAY && DP != PGRAY) {                 // GRAY8 -> Y: 16 + g*219/255      img_yuv_rgb.c:283-326
#pragma unroll
            for (int i = 0; i < 4; i++) y[i] = map_gray2y4(y[i]);
        } else if (SP != PGRAY && DP == PGRAY) {          // Y -> GRAY8 range map             img_yuv_rgb.c:254-261
#pragma unroll
            for (int i = 0; i < 4; i++) y[i] = map_y2gray4(y[i]);
        }
        st4(p.d0 + doff + (size_t)u * 16, valid, y);
        constexpr int ND = chroma_bytes(DP == P420 ? P411 : DP);   // 4:2:0 fill: P/4 bytes per plane = 4 per unit
        constexpr int NS = chroma_bytes(SP);
        if (ND == 0) return;
        if (NS == 0) {                                    // Y8 / GRAY8 source: neutral chroma   img_yuv_planar.c:278-308
            const uint32_t f[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};
            stn<ND>(p.d1 + doff + (size_t)u * ND, valid, f);
            stn<ND>(p.d2 + doff + (size_t)u * ND, valid, f);
        } else {
            uint32_t s[4], d[4];
            ldn<NS>(p.s1 + soff + (size_t)u * NS, valid, s);
            chroma_h<NS, ND>(s, d);
            stn<ND>(p.d1 + doff + (size_t)u * ND, valid, d);
            ldn<NS>(p.s2 + soff + (size_t)u * NS, valid, s);
            chroma_h<NS, ND>(s, d);
            stn<ND>(p.d2 + doff + (size_t)u * ND, valid, d);
        }
    }
};

// Row-pair mode: one side is 4:2:0 (the other is 4:2:0, 4:2:2, 4:1:1 or 4:4:4).
template <int SP, int DP>
struct PlanarRowPair {
    static constexpr int kMode = MODE_ROWPAIR, kStage = 0;
    static __device__ __forceinline__ void plane(const FastParams &p, const uint8_t *s, uint8_t *d, int rp, int unit, bool valid)
    {
        constexpr int NS = chroma_bytes(SP), ND = chroma_bytes(DP);
        const int w = p.w;
        if (SP == P420 && DP == P420) {
            uint32_t a[2];
            ld2(s + (size_t)rp * (w >> 1) + unit * 8, valid, a);
            st2(d + (size_t)rp * (w >> 1) + unit * 8, valid, a);
        } else if (SP == P420) {                          // up-sample vertically by row duplication :66-111
            uint32_t a[2], o[4];
            ld2(s + (size_t)rp * (w >> 1) + unit * 8, valid, a);
            chroma_h<8, ND>(a, o);
            const size_t pitch = DP == P444 ? w : DP == P422 ? (w >> 1) : (w >> 2);
            uint8_t *q = d + (size_t)(2 * rp) * pitch + (size_t)unit * ND;
            stn<ND>(q, valid, o);
            stn<ND>(q + pitch, valid, o);
        } else {                                          // down-sample to 4:2:0 :115-131, :168-181, :215-232
            const size_t pitch = SP == P444 ? w : SP == P422 ? (w >> 1) : (w >> 2);
            const uint8_t *q = s + (size_t)(2 * rp) * pitch + (size_t)unit * NS;
            uint32_t a[4], b[4], o[2];
            ldn<NS>(q, valid, a);
            ldn<NS>(q + pitch, valid, b);
            if (SP == P422) {
                o[0] = avg_up4(a[0], b[0]); o[1] = avg_up4(a[1], b[1]);
            } else if (SP == P411) {                      // vertical mean, then replicate x2
                const uint32_t m = avg_up4(a[0], b[0]);
                o[0] = __byte_perm(m, 0, 0x1100); o[1] = __byte_perm(m, 0, 0x3322);
            } else {                                      // 2x2 box (a+b+c+d+2)/4 in 16-bit lanes
                uint32_t r[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t sa = (a[i] & 0x00FF00FFu) + ((a[i] >> 8) & 0x00FF00FFu);
                    const uint32_t sb = (b[i] & 0x00FF00FFu) + ((b[i] >> 8) & 0x00FF00FFu);
                    r[i] = ((sa + sb + 0x00020002u) >> 2) & 0x00FF00FFu;
                }
                o[0] = __byte_perm(r[0], r[1], 0x6420); o[1] = __byte_perm(r[2], r[3], 0x6420);
            }
            st2(d + (size_t)rp * (w >> 1) + unit * 8, valid, o);
        }
    }
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, int rp, int unit, bool valid,
                                               int, int, uint4 *, int)
    {
        uint32_t y[4];
        const size_t yo = (size_t)(2 * rp) * p.w + unit * 16;
        ld4(p.s0 + soff + yo, valid, y);
        st4(p.d0 + doff + yo, valid, y);
        ld4(p.s0 + soff + yo + p.w, valid, y);
        st4(p.d0 + doff + yo + p.w, valid, y);
        plane(p, p.s1 + soff, p.d1 + doff, rp, unit, valid);
        plane(p, p.s2 + soff, p.d2 + doff, rp, unit, valid);
    }
};

// ---- packed <-> planar-like (img_yuv_mixed.c) ----------------------------------------------------------------
template <int Q> struct PackedInfo {
    static constexpr bool y_even = Q != QUYVY;           // luma in bytes 0,2 (YUY2, YVYU) or 1,3 (UYVY)
    static constexpr bool u_first = Q != QYVYU;          // chroma bytes in memory order are U,V (else V,U)
};

// 8 packed words (16 pixels) -> 4 luma words + 2 U words + 2 V words
template <int Q>
__device__ __forceinline__ void split_packed(const uint32_t *w, uint32_t *yw, uint32_t *uw, uint32_t *vw)
{
    constexpr uint32_t YSEL = PackedInfo<Q>::y_even ? 0x6420 : 0x7531, CSEL = PackedInfo<Q>::y_even ? 0x7531 : 0x6420;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        yw[2 * h] = __byte_perm(w[4 * h], w[4 * h + 1], YSEL);
        yw[2 * h + 1] = __byte_perm(w[4 * h + 2], w[4 * h + 3], YSEL);
        const uint32_t ca = __byte_perm(w[4 * h], w[4 * h + 1], CSEL), cb = __byte_perm(w[4 * h + 2], w[4 * h + 3], CSEL);
        const uint32_t first = __byte_perm(ca, cb, 0x6420), second = __byte_perm(ca, cb, 0x7531);
        uw[h] = PackedInfo<Q>::u_first ? first : second;
        vw[h] = PackedInfo<Q>::u_first ? second : first;
    }
}

// 4 luma words + 2 U words + 2 V words -> 8 packed words
template <int Q>
__device__ __forceinline__ void join_packed(const uint32_t *yw, const uint32_t *uw, const uint32_t *vw, uint32_t *w)
{
    constexpr uint32_t A = Q == QYUY2 ? 0x5140 : Q == QUYVY ? 0x1504 : 0x4150;
    constexpr uint32_t B = Q == QYUY2 ? 0x7362 : Q == QUYVY ? 0x3726 : 0x6372;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t tlo = __byte_perm(uw[h], vw[h], 0x5140), thi = __byte_perm(uw[h], vw[h], 0x7362);
        w[4 * h + 0] = __byte_perm(yw[2 * h], tlo, A);
        w[4 * h + 1] = __byte_perm(yw[2 * h], tlo, B);
        w[4 * h + 2] = __byte_perm(yw[2 * h + 1], thi, A);
        w[4 * h + 3] = __byte_perm(yw[2 * h + 1], thi, B);
    }
}

__device__ __forceinline__ void load_packed16(const uint8_t *p, bool valid, uint32_t *w)
{
    ld4(p, valid, w);
    ld4(p + 16, valid, w + 4);
}

// packed -> 4:2:2 / 4:1:1 / 4:4:4 / Y8 / GRAY8 (linear)
template <int Q, int DP>
struct PackedToPlanarLinear {
    static constexpr int kMode = MODE_LINEAR, kStage = 0;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t, int, uint4 *, int)
    {
        uint32_t w[8], yw[4], uw[2], vw[2];
        load_packed16(p.s0 + soff + (size_t)u * 32, valid, w);
        split_packed<Q>(w, yw, uw, vw);
        if (DP == PGRAY) {
#pragma unroll
            for (int i = 0; i < 4; i++) yw[i] = map_y2gray4(yw[i]);
        }
        st4(p.d0 + doff + (size_t)u * 16, valid, yw);
        constexpr int ND = chroma_bytes(DP);
        if (ND == 0) return;
        uint32_t d[4];
        chroma_h<8, ND>(uw, d);
        stn<ND>(p.d1 + doff + (size_t)u * ND, valid, d);
        chroma_h<8, ND>(vw, d);
        stn<ND>(p.d2 + doff + (size_t)u * ND, valid, d);
    }
};

// packed -> 4:2:0: even row's chroma averaged with the odd row's, (a+b+1)/2 (img_yuv_mixed.c:144-164)
template <int Q>
struct PackedTo420 {
    static constexpr int kMode = MODE_ROWPAIR, kStage = 0;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, int rp, int unit, bool valid,
                                               int, int, uint4 *, int)
    {
        uint32_t w[8], y0[4], y1[4], u0[2], v0[2], u1[2], v1[2];
        const uint8_t *row = p.s0 + soff + ((size_t)(2 * rp) * p.w + unit * 16) * 2;
        load_packed16(row, valid, w);
        split_packed<Q>(w, y0, u0, v0);
        load_packed16(row + (size_t)p.w * 2, valid, w);
        split_packed<Q>(w, y1, u1, v1);
        uint8_t *Y = p.d0 + doff + (size_t)(2 * rp) * p.w + unit * 16;
        st4(Y, valid, y0);
        st4(Y + p.w, valid, y1);
        const uint32_t uo[2] = {avg_up4(u0[0], u1[0]), avg_up4(u0[1], u1[1])};
        const uint32_t vo[2] = {avg_up4(v0[0], v1[0]), avg_up4(v0[1], v1[1])};
        const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
        st2(p.d1 + doff + co, valid, uo);
        st2(p.d2 + doff + co, valid, vo);
    }
};

// Chroma of one unit-row (8 samples per plane) from a planar-like source.
template <int SP>
__device__ __forceinline__ void chroma8_from(const uint8_t *plane, size_t unit_index, bool valid, uint32_t *c)
{
    if (SP == P422 || SP == P420) {
        ld2(plane + unit_index * 8, valid, c);
    } else if (SP == P411) {
        const uint32_t s = ld1(plane + unit_index * 4, valid);
        chroma_h<4, 8>(&s, c);
    } else if (SP == P444) {                              // (a+b)/2 TRUNCATED (img_yuv_mixed.c:130-140)
        uint32_t s[4];
        ld4(plane + unit_index * 16, valid, s);
        c[0] = avg_dn4(__byte_perm(s[0], s[1], 0x6420), __byte_perm(s[0], s[1], 0x7531));
        c[1] = avg_dn4(__byte_perm(s[2], s[3], 0x6420), __byte_perm(s[2], s[3], 0x7531));
    } else {                                              // Y8 / GRAY8: neutral chroma
        c[0] = c[1] = 0x80808080u;
    }
}

// 4:2:2 / 4:1:1 / 4:4:4 / Y8 / GRAY8 -> packed (linear)
template <int SP, int Q>
struct PlanarToPackedLinear {
    static constexpr int kMode = MODE_LINEAR, kStage = 2;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t warp_u0, int nvalid, uint4 *stage, int lane)
    {
        uint32_t yw[4], uw[2], vw[2], w[8];
        ld4(p.s0 + soff + (size_t)u * 16, valid, yw);
        if (SP == PGRAY) {
#pragma unroll
            for (int i = 0; i < 4; i++) yw[i] = map_gray2y4(yw[i]);
        }
        chroma8_from<SP>(p.s1 + soff, u, valid, uw);
        chroma8_from<SP>(p.s2 + soff, u, valid, vw);
        join_packed<Q>(yw, uw, vw, w);
        store_chunks<2>(stage, lane, w, p.d0 + doff + (size_t)warp_u0 * 32, nvalid);
    }
};

// 4:2:0 -> packed: both rows of a pair share the chroma row (img_yuv_mixed.c:88-101)
template <int Q>
struct P420ToPacked {
    static constexpr int kMode = MODE_ROWPAIR, kStage = 2;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, int rp, int unit, bool valid,
                                               int wu0, int nvalid, uint4 *stage, int lane)
    {
        uint32_t yw[4], uw[2], vw[2], w[8];
        const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
        ld2(p.s1 + soff + co, valid, uw);
        ld2(p.s2 + soff + co, valid, vw);
        const uint8_t *Y = p.s0 + soff + (size_t)(2 * rp) * p.w + unit * 16;
        uint8_t *out = p.d0 + doff + ((size_t)(2 * rp) * p.w + (size_t)wu0 * 16) * 2;
        ld4(Y, valid, yw);
        join_packed<Q>(yw, uw, vw, w);
        store_chunks<2>(stage, lane, w, out, nvalid);
        ld4(Y + p.w, valid, yw);
        join_packed<Q>(yw, uw, vw, w);
        store_chunks<2>(stage, lane, w, out + (size_t)p.w * 2, nvalid);
    }
};

// ---- ragged 4:2:0 (width % 16 != 0): flat 16-pixel units, byte-aligned chroma rows -----------------------------
// Rows of such a frame are not 16-byte aligned, so the row-pair kernels above do not apply.  Luma (and every operand
// of the OTHER format) is still walked as flat, aligned 16-pixel units; only the 4:2:0 chroma planes are reached
// through byte-aligned accesses.  OTHER: a PlanarKind (P422 / P444 / P411) or 16 + PackedKind.
template <int OTHER> struct OtherInfo {
    static constexpr bool packed = OTHER >= 16;
    static constexpr int q = OTHER - 16;
    static constexpr int nc = packed ? 8 : chroma_bytes(OTHER);     // chroma bytes per plane per unit-row (planar)
};

// 4:2:0 -> OTHER: the unit's 8 chroma samples come from gather420 (first part of chroma row y/2, continued on the next
// row's when the unit wraps); every other row re-reads them (L1/L2 hits), exactly as the C loops do.
template <int OTHER>
struct Ragged420From {
    static constexpr int kMode = MODE_LINEAR, kStage = OtherInfo<OTHER>::packed ? 2 : 0;
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t warp_u0, int nvalid, uint4 *stage, int lane)
    {
        using OI = OtherInfo<OTHER>;
        uint32_t yw[4], uw[2] = {0, 0}, vw[2] = {0, 0};
        ld4(p.s0 + soff + (size_t)u * 16, valid, yw);
        if (valid) {
            const uint32_t w = (uint32_t)p.w, n = u * 16u, y = n / w, x0 = n - y * w;
            const uint2 uu = gather420(p.s1 + soff, y, x0, w), vv = gather420(p.s2 + soff, y, x0, w);
            uw[0] = uu.x; uw[1] = uu.y; vw[0] = vv.x; vw[1] = vv.y;
        }
        if (OI::packed) {
            uint32_t o[8];
            join_packed<OI::q>(yw, uw, vw, o);
            store_chunks<2>(stage, lane, o, p.d0 + doff + (size_t)warp_u0 * 32, nvalid);
        } else {
            st4(p.d0 + doff + (size_t)u * 16, valid, yw);
            uint32_t d[4];
            chroma_h<8, OI::nc>(uw, d);
            stn<OI::nc>(p.d1 + doff + (size_t)u * OI::nc, valid, d);
            chroma_h<8, OI::nc>(vw, d);
            stn<OI::nc>(p.d2 + doff + (size_t)u * OI::nc, valid, d);
        }
    }
};

// OTHER -> 4:2:0: luma is a flat copy; chroma row r is made from OTHER's rows 2r and 2r+1, so only the part of a unit
// that lies on an EVEN row produces samples (one contiguous run: the whole unit, its head, or -- for a unit that
// wraps from an odd row into an even one -- its tail).  In OTHER's flat layout the sample below is always one row
// pitch further, wrapped or not.
template <int OTHER>
struct Ragged420To {
    static constexpr int kMode = MODE_LINEAR, kStage = 0;
    // the run's samples [j0, j0+cnt) of the given plane (planar) / of both planes (packed), vertically combined
    static __device__ __forceinline__ void combine(const FastParams &p, size_t soff, uint32_t u, uint32_t j0, uint32_t cnt,
                                                   uint32_t *uo, uint32_t *vo)
    {
        using OI = OtherInfo<OTHER>;
        const uint32_t w = (uint32_t)p.w;
        if (OI::packed) {                                  // even row copies, odd row (prev+cur+1)/2   img_yuv_mixed.c:144-164
            uint32_t a[8], b[8], ya[4], ua[2], va[2], ub[2], vb[2];
            const uint8_t *q = p.s0 + soff + (size_t)u * 32 + 4 * j0;
            ldg_bytes<8>(q, 4 * cnt, a);
            ldg_bytes<8>(q + (size_t)w * 2, 4 * cnt, b);
            split_packed<OI::q>(a, ya, ua, va);
            split_packed<OI::q>(b, ya, ub, vb);
            uo[0] = avg_up4(ua[0], ub[0]); uo[1] = avg_up4(ua[1], ub[1]);
            vo[0] = avg_up4(va[0], vb[0]); vo[1] = avg_up4(va[1], vb[1]);
            return;
        }
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t *plane = (pl ? p.s2 : p.s1) + soff;
            uint32_t *o = pl ? vo : uo;
            if (OTHER == P422) {                           // (a+b+1)/2 vertically                         img_yuv_planar.c:168-181
                uint32_t a[2], b[2];
                ldg_bytes<2>(plane + (size_t)u * 8 + j0, cnt, a);
                ldg_bytes<2>(plane + (size_t)u * 8 + j0 + (w >> 1), cnt, b);
                o[0] = avg_up4(a[0], b[0]); o[1] = avg_up4(a[1], b[1]);
            } else if (OTHER == P411) {                    // vertical mean, then replicate x2              :115-131
                uint32_t a[1], b[1];
                ldg_bytes<1>(plane + (size_t)u * 4 + (j0 >> 1), cnt >> 1, a);
                ldg_bytes<1>(plane + (size_t)u * 4 + (j0 >> 1) + (w >> 2), cnt >> 1, b);
                const uint32_t m = avg_up4(a[0], b[0]);
                o[0] = __byte_perm(m, 0, 0x1100); o[1] = __byte_perm(m, 0, 0x3322);
            } else {                                       // 4:4:4: 2x2 box (a+b+c+d+2)/4                   :215-232
                uint32_t a[4], b[4], r[4];
                ldg_bytes<4>(plane + (size_t)u * 16 + 2 * j0, 2 * cnt, a);
                ldg_bytes<4>(plane + (size_t)u * 16 + 2 * j0 + w, 2 * cnt, b);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t sa = (a[i] & 0x00FF00FFu) + ((a[i] >> 8) & 0x00FF00FFu);
                    const uint32_t sb = (b[i] & 0x00FF00FFu) + ((b[i] >> 8) & 0x00FF00FFu);
                    r[i] = ((sa + sb + 0x00020002u) >> 2) & 0x00FF00FFu;
                }
                o[0] = __byte_perm(r[0], r[1], 0x6420); o[1] = __byte_perm(r[2], r[3], 0x6420);
            }
        }
    }
    static __device__ __forceinline__ void run(const FastParams &p, size_t soff, size_t doff, uint32_t u, bool valid,
                                               uint32_t, int, uint4 *, int)
    {
        using OI = OtherInfo<OTHER>;
        uint32_t yw[4];
        if (OI::packed) {
            uint32_t a[8], uw[2], vw[2];
            load_packed16(p.s0 + soff + (size_t)u * 32, valid, a);
            split_packed<OI::q>(a, yw, uw, vw);
        } else {
            ld4(p.s0 + soff + (size_t)u * 16, valid, yw);
        }
        st4(p.d0 + doff + (size_t)u * 16, valid, yw);
        if (!valid) return;
        const uint32_t w = (uint32_t)p.w, cw = w >> 1, n = u * 16u, y = n / w, x0 = n - y * w;
        const uint32_t k = min(8u, (w - x0) >> 1);
        uint32_t ye, j0, cnt, xc;
        if (!(y & 1)) { ye = y; j0 = 0; cnt = k; xc = x0 >> 1; }
        else if (k < 8) { ye = y + 1; j0 = k; cnt = 8 - k; xc = 0; }
        else return;
        uint32_t uo[2], vo[2];
        combine(p, soff, u, j0, cnt, uo, vo);
        const size_t co = (size_t)(ye >> 1) * cw + xc;
        stg_bytes8(p.d1 + doff + co, make_uint2(uo[0], uo[1]), 0, cnt);
        stg_bytes8(p.d2 + doff + co, make_uint2(vo[0], vo[1]), 0, cnt);
    }
};

bool ragged420_family(const ConvertArgs &a, const FastParams &p)
{
    const int nf = a.nframes;
    cudaStream_t st = a.stream;
    const bool from = a.srcfmt == IMG_YUV420P;
    const int other = from ? a.dstfmt : a.srcfmt;
    switch (other) {
    case IMG_YUV420P: return launch_op<PlanarLinear<P411, P411>>(p, nf, st, "ragged 420 copy");   // same plane sizes, flat copy
    case IMG_Y8:      return from ? launch_op<PlanarLinear<P420, PY8>>(p, nf, st, "ragged 420->y8") : launch_op<PlanarLinear<PY8, P420>>(p, nf, st, "ragged y8->420");
    case IMG_GRAY8:   return from ? launch_op<PlanarLinear<P420, PGRAY>>(p, nf, st, "ragged 420->gray8") : launch_op<PlanarLinear<PGRAY, P420>>(p, nf, st, "ragged gray8->420");
    case IMG_YUV422P: return from ? launch_op<Ragged420From<P422>>(p, nf, st, "ragged 420->422") : launch_op<Ragged420To<P422>>(p, nf, st, "ragged 422->420");
    case IMG_YUV444P: return from ? launch_op<Ragged420From<P444>>(p, nf, st, "ragged 420->444") : launch_op<Ragged420To<P444>>(p, nf, st, "ragged 444->420");
    case IMG_YUV411P: return from ? launch_op<Ragged420From<P411>>(p, nf, st, "ragged 420->411") : launch_op<Ragged420To<P411>>(p, nf, st, "ragged 411->420");
    case IMG_YUY2:    return from ? launch_op<Ragged420From<16 + QYUY2>>(p, nf, st, "ragged 420->yuy2") : launch_op<Ragged420To<16 + QYUY2>>(p, nf, st, "ragged yuy2->420");
    case IMG_UYVY:    return from ? launch_op<Ragged420From<16 + QUYVY>>(p, nf, st, "ragged 420->uyvy") : launch_op<Ragged420To<16 + QUYVY>>(p, nf, st, "ragged uyvy->420");
    case IMG_YVYU:    return from ? launch_op<Ragged420From<16 + QYVYU>>(p, nf, st, "ragged 420->yvyu") : launch_op<Ragged420To<16 + QYVYU>>(p, nf, st, "ragged yvyu->420");
    default: return false;
    }
}

// ---- packed <-> packed: one PRMT per 4-byte group, fully linear (img_yuv_packed.c:30-78) ---------------------
__global__ void __launch_bounds__(256) k_wordperm(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                                                 uint32_t sel, uint32_t nchunks)
{
    const uint8_t *s = src + (size_t)blockIdx.y * spitch;
    uint8_t *d = dst + (size_t)blockIdx.y * dpitch;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nchunks; i += stride) {
        const uint4 v = *reinterpret_cast<const uint4 *>(s + (size_t)i * 16);   // plain load: src may alias dst
        stg128(d + (size_t)i * 16, make_uint4(__byte_perm(v.x, 0, sel), __byte_perm(v.y, 0, sel),
                                              __byte_perm(v.z, 0, sel), __byte_perm(v.w, 0, sel)));
    }
}

template <int SP>
bool dispatch_planar_dst(int dp, const FastParams &p, int nf, cudaStream_t st)
{
    switch (dp) {
    case P420:
        if constexpr (SP <= P444) return launch_op<PlanarRowPair<SP, P420>>(p, nf, st, "planar->420");
        else return launch_op<PlanarLinear<SP, P420>>(p, nf, st, "luma->420");
    case P422:
        if constexpr (SP == P420) return launch_op<PlanarRowPair<P420, P422>>(p, nf, st, "420->422");
        else return launch_op<PlanarLinear<SP, P422>>(p, nf, st, "planar->422");
    case P411:
        if constexpr (SP == P420) return launch_op<PlanarRowPair<P420, P411>>(p, nf, st, "420->411");
        else return launch_op<PlanarLinear<SP, P411>>(p, nf, st, "planar->411");
    case P444:
        if constexpr (SP == P420) return launch_op<PlanarRowPair<P420, P444>>(p, nf, st, "420->444");
        else return launch_op<PlanarLinear<SP, P444>>(p, nf, st, "planar->444");
    case PY8:   return launch_op<PlanarLinear<SP, PY8>>(p, nf, st, "planar->y8");
    case PGRAY: return launch_op<PlanarLinear<SP, PGRAY>>(p, nf, st, "planar->gray8");
    default: return false;
    }
}

int planar_kind(int fmt)
{
    switch (fmt) {
    case IMG_YUV420P: return P420;
    case IMG_YUV422P: return P422;
    case IMG_YUV411P: return P411;
    case IMG_YUV444P: return P444;
    case IMG_Y8:      return PY8;
    case IMG_GRAY8:   return PGRAY;
    default:          return -1;
    }
}
int packed_kind(int fmt) { return fmt == IMG_YUY2 ? QYUY2 : fmt == IMG_UYVY ? QUYVY : fmt == IMG_YVYU ? QYVYU : -1; }

template <int Q>
bool dispatch_packed_to(int dp, const FastParams &p, int nf, cudaStream_t st)
{
    switch (dp) {
    case P420:  return launch_op<PackedTo420<Q>>(p, nf, st, "packed->420");
    case P422:  return launch_op<PackedToPlanarLinear<Q, P422>>(p, nf, st, "packed->422");
    case P411:  return launch_op<PackedToPlanarLinear<Q, P411>>(p, nf, st, "packed->411");
    case P444:  return launch_op<PackedToPlanarLinear<Q, P444>>(p, nf, st, "packed->444");
    case PY8:   return launch_op<PackedToPlanarLinear<Q, PY8>>(p, nf, st, "packed->y8");
    case PGRAY: return launch_op<PackedToPlanarLinear<Q, PGRAY>>(p, nf, st, "packed->gray8");
    default: return false;
    }
}

template <int Q>
bool dispatch_to_packed(int sp, const FastParams &p, int nf, cudaStream_t st)
{
    switch (sp) {
    case P420:  return launch_op<P420ToPacked<Q>>(p, nf, st, "420->packed");
    case P422:  return launch_op<PlanarToPackedLinear<P422, Q>>(p, nf, st, "422->packed");
    case P411:  return launch_op<PlanarToPackedLinear<P411, Q>>(p, nf, st, "411->packed");
    case P444:  return launch_op<PlanarToPackedLinear<P444, Q>>(p, nf, st, "444->packed");
    case PY8:   return launch_op<PlanarToPackedLinear<PY8, Q>>(p, nf, st, "y8->packed");
    case PGRAY: return launch_op<PlanarToPackedLinear<PGRAY, Q>>(p, nf, st, "gray8->packed");
    default: return false;
    }
}

}  // namespace

bool launch_wordperm(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, uint32_t sel, size_t bytes,
                     int nframes, cudaStream_t st)
{
    const uint32_t nchunks = (uint32_t)(bytes / 16);
    if (nchunks == 0) return true;
    long gx = ((long)sm_count() * 8 * fast::waves(32) + nframes - 1) / nframes;
    const long maxgx = (nchunks + 255) / 256;
    if (gx < 1) gx = 1;
    if (gx > maxgx) gx = maxgx;
    k_wordperm<<<dim3((unsigned)gx, (unsigned)nframes), 256, 0, st>>>(src, spitch, dst, dpitch, sel, nchunks);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_wordperm");
    return true;
}

bool fast_yuv_family(const ConvertArgs &a, const fast::FastParams &p)
{
    const int sp = planar_kind(a.srcfmt), dp = planar_kind(a.dstfmt);
    const int sq = packed_kind(a.srcfmt), dq = packed_kind(a.dstfmt);
    const int nf = a.nframes;
    cudaStream_t st = a.stream;
    if (p.ragged420) return ragged420_family(a, p);
    if (sp >= 0 && dp >= 0) {
        if (a.srcfmt == IMG_GRAY8 && a.dstfmt == IMG_GRAY8) return launch_op<PlanarLinear<PY8, PY8>>(p, nf, st, "gray8 copy");
        switch (sp) {
        case P420:  return dispatch_planar_dst<P420>(dp, p, nf, st);
        case P422:  return dispatch_planar_dst<P422>(dp, p, nf, st);
        case P411:  return dispatch_planar_dst<P411>(dp, p, nf, st);
        case P444:  return dispatch_planar_dst<P444>(dp, p, nf, st);
        case PY8:   return dispatch_planar_dst<PY8>(dp, p, nf, st);
        case PGRAY: return dispatch_planar_dst<PGRAY>(dp, p, nf, st);
        }
        return false;
    }
    if (sq >= 0 && dp >= 0) {
        switch (sq) {
        case QYUY2: return dispatch_packed_to<QYUY2>(dp, p, nf, st);
        case QUYVY: return dispatch_packed_to<QUYVY>(dp, p, nf, st);
        default:    return dispatch_packed_to<QYVYU>(dp, p, nf, st);
        }
    }
    if (sp >= 0 && dq >= 0) {
        switch (dq) {
        case QYUY2: return dispatch_to_packed<QYUY2>(sp, p, nf, st);
        case QUYVY: return dispatch_to_packed<QUYVY>(sp, p, nf, st);
        default:    return dispatch_to_packed<QYVYU>(sp, p, nf, st);
        }
    }
    if (sq >= 0 && dq >= 0) {
        // per 4-byte group b0 b1 b2 b3 (img_yuv_packed.c): identity / swap16 (b1 b0 b3 b2) / swapuv (b0 b3 b2 b1) /
        // UYVY->YVYU (b1 b2 b3 b0) / YVYU->UYVY (b3 b0 b1 b2)
        uint32_t sel;
        if (sq == dq) sel = 0x3210;
        else if ((sq == QYUY2 && dq == QUYVY) || (sq == QUYVY && dq == QYUY2)) sel = 0x2301;
        else if ((sq == QYUY2 && dq == QYVYU) || (sq == QYVYU && dq == QYUY2)) sel = 0x1230;
        else if (sq == QUYVY) sel = 0x0321;
        else sel = 0x2103;
        return launch_wordperm(a.src.p[0], a.src.pitch, a.dst.p[0], a.dst.pitch, sel, (size_t)a.w * a.h * 2, nf, st);
    }
    return false;
}

}  // namespace acgpu
