// kernels_fast.cu -- tier 2: bandwidth-oriented conversions (16-byte vector accesses, one launch per batch).
//
// Shape shared by every kernel here: a thread owns a *unit* of 16 horizontally adjacent pixels (two rows
// of them when a 4:2:0 plane is involved, so each chroma sample is loaded and decoded once), the 32 units
// of a warp are contiguous in memory, planar operands are read/written as one 128-bit (luma) or 64-bit
// (chroma) access per lane, and byte-interleaved operands (24/32-bit RGB, packed YUV) go through a per-warp
// shared-memory transpose so that every global store instruction writes 512 contiguous bytes.
// Grids are persistent-ish: (blocks_x, frames) with blocks striding over rows / units of their frame.
//
// Arithmetic is the reduced integer form of pixmath.cuh (proved equal to the C path exhaustively).
// Tensor cores are not used: a per-pixel 3x3 integer map with truncation rules is not a dense contraction,
// and the path is HBM-bound (DESIGN.md section 4 has the instruction budget that makes it so).
//
// Domain: every plane pointer and frame pitch 16-byte aligned, sizes on the formats' unit grid, and
// width % 16 == 0 when a 4:2:0 plane is involved (otherwise (width*height) % 16 == 0).  4:2:0 frames of any other even
// width >= 16 with (width*height) % 16 == 0 go through the ragged-row kernels (S420R / D420R here, Ragged420From / To in
// kernels_fast_yuv.cu; their chroma planes only need 4-byte alignment).  Outside it
// convert_fast() returns false and the generic tier runs.
#include "fast_common.cuh"

#include <cuda.h>      // CUtensorMap types only: the encoder is reached through cudaGetDriverEntryPoint, libcuda is not linked

#include <stdlib.h>

namespace acgpu {
namespace {

using namespace fast;

// ---------------------------------------------------------------------------------------------------
// Chroma offset tables (aclib/img_yuv_rgb.c:50-55), built at compile time.  Entry layout is chosen so one
// 64-bit shared load yields two ready-to-use terms:  v[i] = {rV[i] - 256, gV[i] - 256},  u[i] = {bU[i] - 256, gU[i]}
// (the -256 is the Ylut origin shift, see pixmath::ylut_word_fast).
struct ChromaTabs {
    int v[512];
    int u[512];
};
constexpr int chroma_term(int coef, int c) { return (coef * (c - 128) * 16 + 76309 / 2) / 76309; }
constexpr ChromaTabs make_tabs()
{
    ChromaTabs t{};
    for (int i = 0; i < 256; i++) {
        t.v[2 * i] = chroma_term(pixmath::kCRV, i) - 256;
        t.v[2 * i + 1] = chroma_term(pixmath::kCGV, i) - 256;
        t.u[2 * i] = chroma_term(pixmath::kCBU, i) - 256;
        t.u[2 * i + 1] = chroma_term(pixmath::kCGU, i);
    }
    return t;
}
__device__ const ChromaTabs g_tabs = make_tabs();

// Same terms packed two per 32-bit word and biased by +4096 so both halves are non-negative:
//   v16[i] = (rV-256+4096) | (gV-256+4096) << 16,  u16[i] = (bU-256+4096) | (gU+4096) << 16.
// One LDS.32 per table (32 banks -> ~1.7x fewer conflict replays on random chroma than LDS.64); the biases are
// removed for free inside the clamp instruction (VIADDMNMX: relu(min(j - bias, 3498))).
struct ChromaTabs16 {
    uint32_t v[256];
    uint32_t u[256];
};
constexpr ChromaTabs16 make_tabs16()
{
    ChromaTabs16 t{};
    for (int i = 0; i < 256; i++) {
        t.v[i] = (uint32_t)(chroma_term(pixmath::kCRV, i) - 256 + 4096) | ((uint32_t)(chroma_term(pixmath::kCGV, i) - 256 + 4096) << 16);
        t.u[i] = (uint32_t)(chroma_term(pixmath::kCBU, i) - 256 + 4096) | ((uint32_t)(chroma_term(pixmath::kCGU, i) + 4096) << 16);
    }
    return t;
}
__device__ const ChromaTabs16 g_tabs16 = make_tabs16();
#ifndef ACGPU_LUT16
#define ACGPU_LUT16 1
#endif
constexpr int kBiasRB = ACGPU_LUT16 ? 4096 : 0, kBiasG = ACGPU_LUT16 ? 8192 : 0;
#ifndef ACGPU_REPL444
#define ACGPU_REPL444 0
#endif
constexpr int kReplCopies = ACGPU_REPL444 > 0 ? ACGPU_REPL444 : 1;     // private copies of the chroma tables for 4:4:4 sources
constexpr bool kReplOn = ACGPU_REPL444 > 0;

// ---------------------------------------------------------------------------------------------------
// K1: YUV (7 layouts) -> RGB (6 layouts).  aclib/img_yuv_rgb.c:58-136.
//   SWAP   : first colour byte is B instead of R (BGR24, BGRA32, ABGR32)
//   BPP    : 3 or 4;  AFIRST: alpha is byte 0 (ARGB32, ABGR32) -- alpha is never written: for BPP 4 the
//            destination chunk is read, merged with a byte mask, and written back (read-modify-write).

template <int SRC> struct SrcInfo {
    static constexpr bool packed = SRC >= SYUY2 && SRC <= SYVYU;
    static constexpr int nchroma = SRC == S444 ? 16 : SRC == S411 ? 4 : 8;   // chroma samples per 16 pixels
    static constexpr int yo = SRC == SUYVY ? 1 : 0;
    static constexpr int uo = SRC == SYUY2 ? 1 : SRC == SUYVY ? 0 : 3;
    static constexpr int vo = SRC == SYUY2 ? 3 : SRC == SUYVY ? 2 : 1;
};

// Converts one row of 16 pixels.  yw: 4 words (planar luma) or 8 words (packed groups).
// cr/cg/cb: per-chroma-sample offsets.  Emits BPP*4 words.
template <int SRC, bool SWAP, int BPP, bool AFIRST>
__device__ __forceinline__ void convert_row(const uint32_t *yw, const int *cr, const int *cg, const int *cb, uint32_t *out)
{
    using SI = SrcInfo<SRC>;
#pragma unroll
    for (int g = 0; g < 4; g++) {          // 4 pixels per group
        uint32_t xa[4], xg[4], xc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int px = g * 4 + k;
            const int s = SI::nchroma == 16 ? px : SI::nchroma == 8 ? px / 2 : px / 4;
            uint32_t word, sel;
            if (SI::packed) {
                word = yw[px / 2];
                sel = 0x10u << (8 * ((px & 1) * 2 + SI::yo));
            } else {
                word = yw[g];
                sel = 0x10u << (8 * k);
            }
            if (ACGPU_LUT16 && SI::nchroma == 16) {
                // 4:4:4: every pixel has its own chroma sample, so unpacking the table words (5 instructions per sample)
                // costs more than feeding them to dp2a as they are: cr[s] / cb[s] hold the packed V / U table words
                // (low half: R or B term, high half: G term), Y16 = 16*Y is shared by the three channels.
                const uint32_t y16 = __dp4a(word, sel, 0u), tv = (uint32_t)cr[s], tu = (uint32_t)cb[s];
                const uint32_t jr = dp2a_lo_uu(tv, 0x0001u, y16), jb = dp2a_lo_uu(tu, 0x0001u, y16);
                const uint32_t jg = dp2a_lo_uu(tu, 0x0100u, dp2a_lo_uu(tv, 0x0100u, y16));
                xa[k] = clamp_scale<kBiasRB>((int)(SWAP ? jb : jr));
                xg[k] = clamp_scale<kBiasG>((int)jg);
                xc[k] = clamp_scale<kBiasRB>((int)(SWAP ? jr : jb));
            } else {
                xa[k] = channel_word<kBiasRB>(word, sel, SWAP ? cb[s] : cr[s]);
                xg[k] = channel_word<kBiasG>(word, sel, cg[s]);
                xc[k] = channel_word<kBiasRB>(word, sel, SWAP ? cr[s] : cb[s]);
            }
        }
        if (BPP == 3) {
            out[g * 3 + 0] = pack_top4(xa[0], xg[0], xc[0], xa[1]);
            out[g * 3 + 1] = pack_top4(xg[1], xc[1], xa[2], xg[2]);
            out[g * 3 + 2] = pack_top4(xc[2], xa[3], xg[3], xc[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t t = __byte_perm(xa[k], xg[k], 0x0073);
                out[g * 4 + k] = AFIRST ? __byte_perm(t, xc[k], 0x7100) : __byte_perm(t, xc[k], 0x0710);
            }
        }
    }
}

// ---- bulk (TMA) store of a warp tile: shared -> global in one instruction, issued by lane 0 ------------------
// The lane-owned chunks are written to shared memory in plain global order (a 48-byte lane stride is bank-conflict
// free), made visible to the async proxy, and the whole 1536-byte tile leaves with one cp.async.bulk (SASS UBLKCP):
// no LDS read-back, no per-lane STG, no store address arithmetic.  Two buffers per warp alternate so the copy engine
// drains one tile while the warp computes the next.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

template <int K>
__device__ __forceinline__ void store_row_bulk(uint4 *buf, int lane, const uint32_t *ow, uint8_t *gdst, int nvalid)
{
    if (lane == 0) bulk_wait_read<1>();      // this buffer's previous tile has been read out (the other may be in flight)
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) buf[lane * K + k] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) bulk_store(gdst, buf, (uint32_t)nvalid * K * 16);
}

// 32-bit destinations are read-modify-written (alpha is kept).  Asking L2 for the destination tile at the top of
// the iteration takes the DRAM read latency off the critical path between "pixels ready" and "store".
template <int BPP>
__device__ __forceinline__ void prefetch_dest_row2(const uint8_t *rowbase, const uint8_t *rowbase2, int ksplit, int lane, int nvalid)
{
    if (BPP != 4) return;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = j * 32 + lane;
        if (c < nvalid * 4)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(c < ksplit * 4 ? rowbase + (size_t)c * 16 : rowbase2 + (size_t)(c - ksplit * 4) * 16));
    }
}
template <int BPP>
__device__ __forceinline__ void prefetch_dest_row(const uint8_t *rowbase, int lane, int nvalid)
{
    prefetch_dest_row2<BPP>(rowbase, rowbase, 32, lane, nvalid);
}

template <int SRC>
__device__ __forceinline__ void chroma_terms(const int2 *tab, uint32_t U, uint32_t V, int &cr, int &cg, int &cb)
{
#if ACGPU_LUT16
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tab);
    const uint32_t tv = t32[V], tu = t32[256 + U];
    if (SRC == S444) { cr = (int)tv; cb = (int)tu; cg = 0; return; }      // consumed packed, see convert_row
    cr = (int)(tv & 0xFFFFu);
    cb = (int)(tu & 0xFFFFu);
    cg = (int)((tv >> 16) + (tu >> 16));
#else
    const int2 tv = tab[V], tu = tab[256 + U];
    cr = tv.x;
    cg = tv.y + tu.y;
    cb = tu.x;
#endif
}

template <int SRC, bool SWAP, int BPP, bool AFIRST, bool BULK, bool FLAT>
__global__ void __launch_bounds__(256, FLAT ? 4 : 5) k_yuv2rgb(FastParams p)   // 48 regs: measured best of 3..6 blocks (profiles/r1_experiments.md)
{
    static_assert(!BULK || BPP == 3, "bulk stores cannot merge the untouched alpha byte");
    using SI = SrcInfo<SRC>;
    __shared__ int2 s_tab[512];
    extern __shared__ uint4 s_stage[];
#if ACGPU_LUT16
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = reinterpret_cast<const int2 *>(&g_tabs16)[i];
#else
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tab[i] = reinterpret_cast<const int2 *>(&g_tabs)[i];
#endif
    // 4:4:4 looks two table words up per PIXEL with uncorrelated indices: on one copy of the tables 55 % of the shared
    // wavefronts were bank-conflict replays and the L1/shared pipe ran at 95 % (profiles/r1c_ncu_secondary_kernels.md).
    // That layout gets kReplCopies private copies, interleaved (entry e of copy c = word e * kReplCopies + c) so that a lane only
    // shares banks with the lanes of its own copy.  MEASURED AND REJECTED (ACGPU_REPL444 = 0 is the default build): 32 copies
    // (64 KB: two blocks per SM) are conflict-free but lose to the occupancy they cost, 0.835 -> 0.625 of peak; 16 copies
    // keep five blocks per SM and reach 0.82 on random bytes but also on a smooth picture, where the single copy runs at
    // 0.98 (neighbouring lanes hit the same words and the hardware broadcasts) -- profiles/r2_experiments.md.
    constexpr bool kRepl = SRC == S444 && !BULK && ACGPU_LUT16 && kReplOn;
    uint32_t *const rep = reinterpret_cast<uint32_t *>(s_stage + (blockDim.x / 32) * 32 * BPP * (BULK ? 2 : 1));
    if (kRepl)
        for (int i = threadIdx.x; i < 512 * kReplCopies; i += blockDim.x) rep[i] = reinterpret_cast<const uint32_t *>(&g_tabs16)[i / kReplCopies];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage = s_stage + warp * (32 * BPP) * (BULK ? 2 : 1);
    uint4 *stage2 = stage + 32 * BPP;      // BULK: second tile buffer
    (void)stage2;
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    uint8_t *dst = p.d0 + doff;

    if (SRC == S420 && FLAT) {
        // Flat 4:2:0 mode: the (row pair, unit) grid is walked as one linear sequence, so every warp carries 32 real
        // units even when the row does not fill whole warps (1080p: 120 units = 3.75 warps).  A warp may then straddle
        // the end of a row pair: its first `k` lanes finish row pair A, the rest start row pair A+1, and the transposed
        // store writes two segments.  No per-lane division: the warp's (rpA, uA) advance by constants each trip.
        const uint8_t *Y = p.s0 + soff, *U = p.s1 + soff, *V = p.s2 + soff;
        const uint32_t upr = (uint32_t)p.upr, total = (uint32_t)p.nrp * upr;
        const uint32_t stride = gridDim.x * blockDim.x, qs = stride / upr, rs = stride - qs * upr;
        uint32_t g0 = blockIdx.x * blockDim.x + warp * 32;
        uint32_t rpA = g0 / upr, uA = g0 - rpA * upr;
        const size_t rowb = (size_t)p.w * BPP;
        for (; g0 < total; g0 += stride) {
            const int nvalid = (int)min(32u, total - g0);
            const int k = (int)min((uint32_t)nvalid, upr - uA);           // lanes that still belong to row pair A
            uint32_t rp = rpA, unit = uA + lane;
            if (unit >= upr) { unit -= upr; rp++; }
            const bool valid = lane < nvalid;
            uint32_t y0[4] = {0, 0, 0, 0}, y1[4] = {0, 0, 0, 0};
            uint2 uu = make_uint2(0, 0), vv = make_uint2(0, 0);
            if (valid) {
                const uint8_t *yp = Y + (size_t)(2 * rp) * p.w + unit * 16;
                const uint4 a = ldg128(yp), b = ldg128(yp + p.w);
                y0[0] = a.x; y0[1] = a.y; y0[2] = a.z; y0[3] = a.w;
                y1[0] = b.x; y1[1] = b.y; y1[2] = b.z; y1[3] = b.w;
                const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
                uu = ldg64(U + co);
                vv = ldg64(V + co);
            }
            uint8_t *segA = dst + ((size_t)(2 * rpA) * p.w + (size_t)uA * 16) * BPP;
            uint8_t *segB = dst + (size_t)(2 * (rpA + 1)) * rowb;
            prefetch_dest_row2<BPP>(segA, segB, k, lane, nvalid);
            prefetch_dest_row2<BPP>(segA + rowb, segB + rowb, k, lane, nvalid);
            int cr[8], cg[8], cb[8];
#pragma unroll
            for (int s = 0; s < 8; s++)
                chroma_terms<SRC>(s_tab, byte_of(s < 4 ? uu.x : uu.y, s & 3), byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
            uint32_t ow[BPP * 4];
            convert_row<SRC, SWAP, BPP, AFIRST>(y0, cr, cg, cb, ow);
            store_row_rgb2<BPP, AFIRST>(stage, lane, ow, segA, segB, k, nvalid);
            convert_row<SRC, SWAP, BPP, AFIRST>(y1, cr, cg, cb, ow);
            store_row_rgb2<BPP, AFIRST>(stage, lane, ow, segA + rowb, segB + rowb, k, nvalid);
            rpA += qs;
            uA += rs;
            if (uA >= upr) { uA -= upr; rpA++; }
        }
    } else if (SRC == S420) {
        const uint8_t *Y = p.s0 + soff, *U = p.s1 + soff, *V = p.s2 + soff;
        const int unit = blockIdx.z * blockDim.x + threadIdx.x;          // blockIdx.z: column segment of wide rows
        const int wu0 = blockIdx.z * blockDim.x + warp * 32;             // first unit of this warp
        const bool valid = unit < p.upr;
        const int nvalid = min(32, p.upr - wu0);
        if (nvalid <= 0) return;
        for (int rp = blockIdx.x; rp < p.nrp; rp += gridDim.x) {
            uint32_t y0[4] = {0, 0, 0, 0}, y1[4] = {0, 0, 0, 0};
            uint2 uu = make_uint2(0, 0), vv = make_uint2(0, 0);
            if (valid) {
                const uint8_t *yp = Y + (size_t)(2 * rp) * p.w + unit * 16;
                const uint4 a = ldg128(yp), b = ldg128(yp + p.w);
                y0[0] = a.x; y0[1] = a.y; y0[2] = a.z; y0[3] = a.w;
                y1[0] = b.x; y1[1] = b.y; y1[2] = b.z; y1[3] = b.w;
                const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
                uu = ldg64(U + co);
                vv = ldg64(V + co);
            }
            uint8_t *row0 = dst + ((size_t)(2 * rp) * p.w + (size_t)wu0 * 16) * BPP;
            prefetch_dest_row<BPP>(row0, lane, nvalid);
            prefetch_dest_row<BPP>(row0 + (size_t)p.w * BPP, lane, nvalid);
            int cr[8], cg[8], cb[8];
#pragma unroll
            for (int s = 0; s < 8; s++)
                chroma_terms<SRC>(s_tab, byte_of(s < 4 ? uu.x : uu.y, s & 3), byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
            uint32_t ow[BPP * 4];
            convert_row<SRC, SWAP, BPP, AFIRST>(y0, cr, cg, cb, ow);
            if (BULK) store_row_bulk<BPP>(stage, lane, ow, row0, nvalid);
            else store_row_rgb<BPP, AFIRST>(stage, lane, ow, row0, nvalid);
            convert_row<SRC, SWAP, BPP, AFIRST>(y1, cr, cg, cb, ow);
            if (BULK) store_row_bulk<BPP>(stage2, lane, ow, row0 + (size_t)p.w * BPP, nvalid);
            else store_row_rgb<BPP, AFIRST>(stage, lane, ow, row0 + (size_t)p.w * BPP, nvalid);
        }
        if (BULK && lane == 0) bulk_wait_all<0>();
    } else {
        const uint8_t *S0 = p.s0 + soff, *S1 = p.s1 + soff, *S2 = p.s2 + soff;
        const uint32_t stride = gridDim.x * blockDim.x;
        bool flip = false;
        for (uint32_t base = blockIdx.x * blockDim.x; base < p.nunits; base += stride) {
            const uint32_t u = base + threadIdx.x;
            const uint32_t warp_u0 = base + warp * 32;
            if (warp_u0 >= p.nunits) break;                  // warp-uniform
            const int nvalid = (int)min(32u, p.nunits - warp_u0);
            const bool valid = u < p.nunits;
            prefetch_dest_row<BPP>(dst + (size_t)warp_u0 * 16 * BPP, lane, nvalid);
            uint32_t yw[SI::packed ? 8 : 4];
            int cr[SI::nchroma], cg[SI::nchroma], cb[SI::nchroma];
            if (SI::packed) {
                uint4 a = make_uint4(0, 0, 0, 0), b = a;
                if (valid) { a = ldg128(S0 + (size_t)u * 32); b = ldg128(S0 + (size_t)u * 32 + 16); }
                yw[0] = a.x; yw[1] = a.y; yw[2] = a.z; yw[3] = a.w;
                yw[4] = b.x; yw[5] = b.y; yw[6] = b.z; yw[7] = b.w;
#pragma unroll
                for (int s = 0; s < 8; s++)
                    chroma_terms<SRC>(s_tab, byte_of(yw[s], SI::uo), byte_of(yw[s], SI::vo), cr[s], cg[s], cb[s]);
            } else {
                uint4 a = make_uint4(0, 0, 0, 0);
                if (valid) a = ldg128(S0 + (size_t)u * 16);
                yw[0] = a.x; yw[1] = a.y; yw[2] = a.z; yw[3] = a.w;
                if (SRC == S444) {
                    uint4 uu = make_uint4(0, 0, 0, 0), vv = uu;
                    if (valid) { uu = ldg128(S1 + (size_t)u * 16); vv = ldg128(S2 + (size_t)u * 16); }
#pragma unroll
                    for (int s = 0; s < 16; s++) {
                        if (kRepl) {        // packed table words, consumed as they are by convert_row
                            const uint32_t *rt = rep + (lane & (kReplCopies - 1));
                            cr[s] = (int)rt[byte_of(word_of(vv, s >> 2), s & 3) * kReplCopies];
                            cb[s] = (int)rt[(256u + byte_of(word_of(uu, s >> 2), s & 3)) * kReplCopies];
                            cg[s] = 0;
                        } else {
                            chroma_terms<SRC>(s_tab, byte_of(word_of(uu, s >> 2), s & 3), byte_of(word_of(vv, s >> 2), s & 3), cr[s], cg[s], cb[s]);
                        }
                    }
                } else if (SRC == S422) {
                    uint2 uu = make_uint2(0, 0), vv = uu;
                    if (valid) { uu = ldg64(S1 + (size_t)u * 8); vv = ldg64(S2 + (size_t)u * 8); }
#pragma unroll
                    for (int s = 0; s < 8; s++)
                        chroma_terms<SRC>(s_tab, byte_of(s < 4 ? uu.x : uu.y, s & 3), byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
                } else if (SRC == S420R) {
                    // ragged 4:2:0: the unit's 16 flat pixels start at column x0 (even) of row y; their chroma samples sit
                    // at an arbitrary byte of chroma row y/2 (and continue on the next row's when the unit wraps)
                    uint2 uu = make_uint2(0, 0), vv = uu;
                    if (valid) {
                        const uint32_t w = (uint32_t)p.w, n = u * 16u, y = n / w, x0 = n - y * w;
                        uu = gather420(S1, y, x0, w);
                        vv = gather420(S2, y, x0, w);
                    }
#pragma unroll
                    for (int s = 0; s < 8; s++)
                        chroma_terms<SRC>(s_tab, byte_of(s < 4 ? uu.x : uu.y, s & 3), byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
                } else {   // S411
                    uint32_t uu = 0, vv = 0;
                    if (valid) { uu = ldg32(S1 + (size_t)u * 4); vv = ldg32(S2 + (size_t)u * 4); }
#pragma unroll
                    for (int s = 0; s < 4; s++) chroma_terms<SRC>(s_tab, byte_of(uu, s), byte_of(vv, s), cr[s], cg[s], cb[s]);
                }
            }
            uint32_t ow[BPP * 4];
            convert_row<SRC, SWAP, BPP, AFIRST>(yw, cr, cg, cb, ow);
            if (BULK) {
                store_row_bulk<BPP>(flip ? stage2 : stage, lane, ow, dst + (size_t)warp_u0 * 16 * BPP, nvalid);
                flip = !flip;
            } else {
                store_row_rgb<BPP, AFIRST>(stage, lane, ow, dst + (size_t)warp_u0 * 16 * BPP, nvalid);
            }
        }
        if (BULK && lane == 0) bulk_wait_all<0>();
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: RGB (6 layouts) -> YUV (7 layouts + Y8).  aclib/img_yuv_rgb.c:142-221.
// Every pixel yields Y; chroma is POINT-sampled, never averaged: 4:2:0 takes U at (even x, even y) and V at
// (odd x, odd y); 4:2:2 / packed take U at even x and V at odd x (YVYU: V at even x, U at odd x); 4:1:1 takes
// U at x%4==0 and V at x%4==2.  Each component is two dp2a instructions on the pixel word (16-bit
// coefficients x 8-bit samples) with the rounding constant and the +16/+128 offset folded into the
// accumulator, so the answer is simply byte 2 of the accumulator.

// accumulators whose byte 2 is the component value
template <int SL> __device__ __forceinline__ uint32_t acc_y(uint32_t px)
{
    using RI = RgbInfo<SL>;
    constexpr uint32_t lo = RI::half(16829, 33039, 6416, 0), hi = RI::half(16829, 33039, 6416, 2);
    return dp2a_hi_uu(hi, px, dp2a_lo_uu(lo, px, 32768u + (16u << 16)));
}
template <int SL> __device__ __forceinline__ uint32_t acc_u(uint32_t px)
{
    using RI = RgbInfo<SL>;
    constexpr uint32_t lo = RI::half(-9714, -19070, 28784, 0), hi = RI::half(-9714, -19070, 28784, 2);
    return dp2a_hi_su(hi, px, dp2a_lo_su(lo, px, 32768u + (128u << 16)));
}
template <int SL> __device__ __forceinline__ uint32_t acc_v(uint32_t px)
{
    using RI = RgbInfo<SL>;
    constexpr uint32_t lo = RI::half(28784, -24103, -4681, 0), hi = RI::half(28784, -24103, -4681, 2);
    return dp2a_hi_su(hi, px, dp2a_lo_su(lo, px, 32768u + (128u << 16)));
}
template <int SL, int DST>
__global__ void __launch_bounds__(256, 4) k_rgb2yuv(FastParams p)
{
    using RI = RgbInfo<SL>;
    extern __shared__ uint4 s_stage[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *stage = s_stage + warp * 64;
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    const uint8_t *src = p.s0 + soff;
    uint8_t *Y = p.d0 + doff, *U = p.d1 + doff, *V = p.d2 + doff;

    if (DST == D420) {
        const uint32_t unit = blockIdx.z * blockDim.x + threadIdx.x;
        const bool valid = (int)unit < p.upr;
        if (p.upr - (int)(blockIdx.z * blockDim.x + warp * 32) <= 0) return;
        for (int rp = blockIdx.x; rp < p.nrp; rp += gridDim.x) {
            uint32_t px[16], ay[16];
            const uint8_t *row0 = src + (size_t)(2 * rp) * p.w * RI::bpp;
            // row 0: Y everywhere, U from even pixels
            load_rgb16<SL>(row0, unit, valid, px);
#pragma unroll
            for (int k = 0; k < 16; k++) ay[k] = acc_y<SL>(px[k]);
            uint32_t uacc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) uacc[k] = acc_u<SL>(px[2 * k]);
            if (valid) {
                stg128(Y + (size_t)(2 * rp) * p.w + unit * 16,
                       make_uint4(pack_b2x4(ay[0], ay[1], ay[2], ay[3]), pack_b2x4(ay[4], ay[5], ay[6], ay[7]),
                                  pack_b2x4(ay[8], ay[9], ay[10], ay[11]), pack_b2x4(ay[12], ay[13], ay[14], ay[15])));
                stg64(U + (size_t)rp * (p.w >> 1) + unit * 8,
                      make_uint2(pack_b2x4(uacc[0], uacc[1], uacc[2], uacc[3]), pack_b2x4(uacc[4], uacc[5], uacc[6], uacc[7])));
            }
            // row 1: Y everywhere, V from odd pixels
            load_rgb16<SL>(row0 + (size_t)p.w * RI::bpp, unit, valid, px);
#pragma unroll
            for (int k = 0; k < 16; k++) ay[k] = acc_y<SL>(px[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) uacc[k] = acc_v<SL>(px[2 * k + 1]);
            if (valid) {
                stg128(Y + (size_t)(2 * rp + 1) * p.w + unit * 16,
                       make_uint4(pack_b2x4(ay[0], ay[1], ay[2], ay[3]), pack_b2x4(ay[4], ay[5], ay[6], ay[7]),
                                  pack_b2x4(ay[8], ay[9], ay[10], ay[11]), pack_b2x4(ay[12], ay[13], ay[14], ay[15])));
                stg64(V + (size_t)rp * (p.w >> 1) + unit * 8,
                      make_uint2(pack_b2x4(uacc[0], uacc[1], uacc[2], uacc[3]), pack_b2x4(uacc[4], uacc[5], uacc[6], uacc[7])));
            }
        }
    } else {
        const uint32_t stride = gridDim.x * blockDim.x;
        for (uint32_t base = blockIdx.x * blockDim.x; base < p.nunits; base += stride) {
            const uint32_t u = base + threadIdx.x;
            const uint32_t warp_u0 = base + warp * 32;
            if (warp_u0 >= p.nunits) break;
            const int nvalid = (int)min(32u, p.nunits - warp_u0);
            const bool valid = u < p.nunits;
            uint32_t px[16], ay[16];
            load_rgb16<SL>(src, u, valid, px);
#pragma unroll
            for (int k = 0; k < 16; k++) ay[k] = acc_y<SL>(px[k]);
            if (DST >= DYUY2 && DST <= DYVYU) {
                // cells (Y, C): C = U of even pixels / V of odd pixels (YVYU: the other way round)
                uint32_t ow[8];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint32_t c0 = DST == DYVYU ? acc_v<SL>(px[2 * g]) : acc_u<SL>(px[2 * g]);
                    const uint32_t c1 = DST == DYVYU ? acc_u<SL>(px[2 * g + 1]) : acc_v<SL>(px[2 * g + 1]);
                    ow[g] = DST == DUYVY ? pack_b2x4(c0, ay[2 * g], c1, ay[2 * g + 1]) : pack_b2x4(ay[2 * g], c0, ay[2 * g + 1], c1);
                }
                store_chunks<2>(stage, lane, ow, Y + (size_t)warp_u0 * 32, nvalid);
                continue;
            }
            if (valid)
                stg128(Y + (size_t)u * 16,
                       make_uint4(pack_b2x4(ay[0], ay[1], ay[2], ay[3]), pack_b2x4(ay[4], ay[5], ay[6], ay[7]),
                                  pack_b2x4(ay[8], ay[9], ay[10], ay[11]), pack_b2x4(ay[12], ay[13], ay[14], ay[15])));
            if (DST == D422) {
                uint32_t ua[8], va[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { ua[k] = acc_u<SL>(px[2 * k]); va[k] = acc_v<SL>(px[2 * k + 1]); }
                if (valid) {
                    stg64(U + (size_t)u * 8, make_uint2(pack_b2x4(ua[0], ua[1], ua[2], ua[3]), pack_b2x4(ua[4], ua[5], ua[6], ua[7])));
                    stg64(V + (size_t)u * 8, make_uint2(pack_b2x4(va[0], va[1], va[2], va[3]), pack_b2x4(va[4], va[5], va[6], va[7])));
                }
            } else if (DST == D411) {
                uint32_t ua[4], va[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { ua[k] = acc_u<SL>(px[4 * k]); va[k] = acc_v<SL>(px[4 * k + 2]); }
                if (valid) {
                    stg32(U + (size_t)u * 4, pack_b2x4(ua[0], ua[1], ua[2], ua[3]));
                    stg32(V + (size_t)u * 4, pack_b2x4(va[0], va[1], va[2], va[3]));
                }
            } else if (DST == D420R) {
                // ragged 4:2:0: the row decides what is sampled -- U from the even pixels of even rows, V from the odd
                // pixels of odd rows -- and the samples land at an arbitrary byte of chroma row y/2.  A unit that wraps
                // into row y+1 after k samples switches plane (and sampling phase) there.
                if (valid) {
                    using RI2 = RgbInfo<SL>;
                    constexpr uint32_t ulo = RI2::half(-9714, -19070, 28784, 0), uhi = RI2::half(-9714, -19070, 28784, 2);
                    constexpr uint32_t vlo = RI2::half(28784, -24103, -4681, 0), vhi = RI2::half(28784, -24103, -4681, 2);
                    const uint32_t w = (uint32_t)p.w, cw = w >> 1, n = u * 16u, y = n / w, x0 = n - y * w;
                    const uint32_t k = min(8u, (w - x0) >> 1);
                    const bool oddA = y & 1;
                    uint32_t c[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const bool odd = oddA != ((uint32_t)j >= k);
                        const uint32_t s = odd ? px[2 * j + 1] : px[2 * j];
                        c[j] = dp2a_hi_su(odd ? vhi : uhi, s, dp2a_lo_su(odd ? vlo : ulo, s, 32768u + (128u << 16)));
                    }
                    const uint2 v = make_uint2(pack_b2x4(c[0], c[1], c[2], c[3]), pack_b2x4(c[4], c[5], c[6], c[7]));
                    stg_bytes8((oddA ? V : U) + (size_t)(y >> 1) * cw + (x0 >> 1), v, 0, k);
                    if (k < 8) stg_bytes8((oddA ? U : V) + (size_t)((y + 1) >> 1) * cw, v, k, 8 - k);
                }
            } else if (DST == D444) {
                uint32_t ca[16];
#pragma unroll
                for (int k = 0; k < 16; k++) ca[k] = acc_u<SL>(px[k]);
                if (valid)
                    stg128(U + (size_t)u * 16,
                           make_uint4(pack_b2x4(ca[0], ca[1], ca[2], ca[3]), pack_b2x4(ca[4], ca[5], ca[6], ca[7]),
                                      pack_b2x4(ca[8], ca[9], ca[10], ca[11]), pack_b2x4(ca[12], ca[13], ca[14], ca[15])));
#pragma unroll
                for (int k = 0; k < 16; k++) ca[k] = acc_v<SL>(px[k]);
                if (valid)
                    stg128(V + (size_t)u * 16,
                           make_uint4(pack_b2x4(ca[0], ca[1], ca[2], ca[3]), pack_b2x4(ca[4], ca[5], ca[6], ca[7]),
                                      pack_b2x4(ca[8], ca[9], ca[10], ca[11]), pack_b2x4(ca[12], ca[13], ca[14], ca[15])));
            }
        }
    }
}

inline bool flat420_enabled()
{
    static const bool on = [] { const char *e = getenv("ACGPU_FLAT420"); return !e || atoi(e) != 0; }();   // profiling knob
    return on;
}

template <int SRC, bool SWAP, int BPP, bool AFIRST, bool BULK = false>
bool launch_yuv2rgb(const FastParams &p, int nframes, cudaStream_t st)
{
    FastParams q = p;
    // Flat 4:2:0 mode packs warps across row-pair boundaries.  It costs registers (64 vs 48) and a two-segment store,
    // so it only pays when the row-per-block shape would idle >20 % of the lanes: PAL 720 (45 units -> 64 lanes)
    // 0.74 -> 0.84 of peak; 1280 (80 -> 96) break-even; 1920 / 3840 (120 -> 128, 240 -> 256) lose 6-8 % (measured).
    // It needs at least a warp of units per row (a warp then straddles at most one row-pair boundary).
    const int lanes_row = ((p.upr + 31) / 32) * 32;
    q.flat420 = SRC == S420 && !BULK && p.upr >= 32 && p.upr * 10 < lanes_row * 8
             && (uint64_t)p.upr * p.nrp < 0x7FFFFFFFu && flat420_enabled();
    constexpr bool kRepl = SRC == S444 && !BULK && ACGPU_LUT16 && kReplOn;       // private table copies (see the kernel)
    LaunchShape s = SRC != S420 ? shape_linear(p.nunits, nframes, kRepl ? 4 : 8)
                  : q.flat420   ? shape_linear((uint32_t)(p.upr * p.nrp), nframes, 8)
                                : shape_420(p.upr, p.nrp, nframes, 8);
    const size_t smem = (size_t)(s.block.x / 32) * 32 * BPP * sizeof(uint4) * (BULK ? 2 : 1) + (kRepl ? 512 * kReplCopies * sizeof(uint32_t) : 0);
    if (kRepl && !check(cudaFuncSetAttribute(k_yuv2rgb<SRC, SWAP, BPP, AFIRST, BULK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "k_yuv2rgb smem"))
        return false;
    if (SRC == S420 && !BULK && q.flat420) k_yuv2rgb<SRC, SWAP, BPP, AFIRST, false, (SRC == S420 && !BULK)><<<s.grid, s.block, smem, st>>>(q);
    else k_yuv2rgb<SRC, SWAP, BPP, AFIRST, BULK, false><<<s.grid, s.block, smem, st>>>(q);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuv2rgb");
    return true;
}

template <int SRC>
bool dispatch_yuv2rgb_dst(int dstfmt, const FastParams &p, int nframes, cudaStream_t st)
{
    switch (dstfmt) {
    case IMG_RGB24:  return launch_yuv2rgb<SRC, false, 3, false>(p, nframes, st);
    case IMG_BGR24:  return launch_yuv2rgb<SRC, true, 3, false>(p, nframes, st);
    case IMG_RGBA32: return launch_yuv2rgb<SRC, false, 4, false>(p, nframes, st);
    case IMG_BGRA32: return launch_yuv2rgb<SRC, true, 4, false>(p, nframes, st);
    case IMG_ARGB32: return launch_yuv2rgb<SRC, false, 4, true>(p, nframes, st);
    case IMG_ABGR32: return launch_yuv2rgb<SRC, true, 4, true>(p, nframes, st);
    default: return false;
    }
}

// ---------------------------------------------------------------------------------------------------
// Tier 3b (experiment, selectable): YUV420P -> RGB24/BGR24 with the planar sources staged by bulk-async copies.
// Each warp runs its own two-stage pipeline: lane 0 issues four cp.async.bulk (two luma row segments, one U and one V
// segment: 512 + 512 + 256 + 256 bytes) for the NEXT row pair into shared memory, completion is signalled on a
// per-stage mbarrier (expect_tx), and the warp converts the CURRENT row pair out of shared memory meanwhile.  The
// loads therefore overlap the arithmetic without holding registers (the register-prefetch variant lost to
// occupancy).  Output goes through the tier-2 transpose (LDS+STG) or, with BULKOUT, a bulk store.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool SWAP, bool BULKOUT>
__global__ void __launch_bounds__(256, 4) k_yuv420_rgb24_tma(FastParams p)
{
    constexpr int BPP = 3;
    __shared__ uint32_t s_tab[512];
    extern __shared__ uint4 s_dyn[];
    // per warp: 2 stages x 96 uint4 (y0: 32, y1: 32, u: 16, v: 16) + output staging (96 or 2 x 96 uint4) + 2 mbarriers
    constexpr int kIn = 96, kOut = BULKOUT ? 192 : 96, kPerWarp = 2 * kIn + kOut + 1;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tab[i] = reinterpret_cast<const uint32_t *>(&g_tabs16)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *wbase = s_dyn + warp * kPerWarp;
    uint4 *in0 = wbase, *in1 = wbase + kIn, *stage = wbase + 2 * kIn;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + 2 * kIn + kOut);
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    const uint8_t *Y = p.s0 + soff, *U = p.s1 + soff, *V = p.s2 + soff;
    uint8_t *dst = p.d0 + doff;
    const int wu0 = blockIdx.z * blockDim.x + warp * 32;
    const int nvalid = min(32, p.upr - wu0);
    if (nvalid <= 0) return;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t ybytes = (uint32_t)nvalid * 16, cbytes = (uint32_t)nvalid * 8;
    auto issue = [&](int rp, int st) {       // lane 0 only
        uint4 *b = st ? in1 : in0;
        mbar_expect_tx(&bars[st], 2 * ybytes + 2 * cbytes);
        const uint8_t *yp = Y + (size_t)(2 * rp) * p.w + (size_t)wu0 * 16;
        bulk_load(b, yp, ybytes, &bars[st]);
        bulk_load(b + 32, yp + p.w, ybytes, &bars[st]);
        const size_t co = (size_t)rp * (p.w >> 1) + (size_t)wu0 * 8;
        bulk_load(b + 64, U + co, cbytes, &bars[st]);
        bulk_load(b + 80, V + co, cbytes, &bars[st]);
    };
    int rp = blockIdx.x;
    if (rp < p.nrp && lane == 0) issue(rp, 0);
    for (int it = 0; rp < p.nrp; rp += gridDim.x, it++) {
        const int st = it & 1;
        if (lane == 0 && rp + (int)gridDim.x < p.nrp) issue(rp + gridDim.x, st ^ 1);
        mbar_wait(&bars[st], (it >> 1) & 1);
        const uint4 *b = st ? in1 : in0;
        const uint4 a = b[lane], c = b[32 + lane];
        const uint2 uu = reinterpret_cast<const uint2 *>(b + 64)[lane], vv = reinterpret_cast<const uint2 *>(b + 80)[lane];
        const uint32_t y0[4] = {a.x, a.y, a.z, a.w}, y1[4] = {c.x, c.y, c.z, c.w};
        int cr[8], cg[8], cb[8];
#pragma unroll
        for (int s = 0; s < 8; s++)
            chroma_terms<S420>(reinterpret_cast<const int2 *>(s_tab), byte_of(s < 4 ? uu.x : uu.y, s & 3),
                               byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
        uint32_t ow[12];
        uint8_t *row0 = dst + ((size_t)(2 * rp) * p.w + (size_t)wu0 * 16) * BPP;
        convert_row<S420, SWAP, BPP, false>(y0, cr, cg, cb, ow);
        if (BULKOUT) store_row_bulk<BPP>(stage, lane, ow, row0, nvalid);
        else store_row_rgb<BPP, false>(stage, lane, ow, row0, nvalid);
        convert_row<S420, SWAP, BPP, false>(y1, cr, cg, cb, ow);
        if (BULKOUT) store_row_bulk<BPP>(stage + 96, lane, ow, row0 + (size_t)p.w * BPP, nvalid);
        else store_row_rgb<BPP, false>(stage, lane, ow, row0 + (size_t)p.w * BPP, nvalid);
        __syncwarp();        // every lane is done with this input stage before lane 0 refills it
    }
    if (BULKOUT && lane == 0) bulk_wait_all<0>();
}

template <bool SWAP, bool BULKOUT>
bool launch_yuv420_rgb24_tma(const FastParams &p, int nframes, cudaStream_t st)
{
    const LaunchShape s = shape_420(p.upr, p.nrp, nframes, 8);
    const size_t smem = (size_t)(s.block.x / 32) * (2 * 96 + (BULKOUT ? 192 : 96) + 1) * sizeof(uint4);
    if (smem > 48 * 1024
        && !check(cudaFuncSetAttribute(k_yuv420_rgb24_tma<SWAP, BULKOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr"))
        return false;
    k_yuv420_rgb24_tma<SWAP, BULKOUT><<<s.grid, s.block, smem, st>>>(p);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuv420_rgb24_tma");
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Tier 3c (experiment, selectable: $ACGPU_TMA = 3 / 4): YUV420P -> RGB24/BGR24 with real tensor maps.
//   stores ($ACGPU_TMA=3 and 4): the RGB24 batch is a 3-D tensor [frame][row][w*3/8 x uint64]; a warp's two output rows
//     (2 x 1536 bytes) leave shared memory with ONE cp.async.bulk.tensor.3d store of a {192, 2, 1} box (SASS UTMASTG) --
//     no LDS read-back, no per-lane STG, no store address arithmetic, partial warps clipped by the tensor bounds;
//   loads ($ACGPU_TMA=4): Y is [frame][row][w/4 x uint32] (box {128, 2, 1}), U and V are [frame][row/2][w/8 x uint32]
//     (box {64, 1, 1}); each warp runs a two-stage pipeline of three tensor loads (SASS UTMALDG) per row pair onto an
//     mbarrier, as tier 3b does with four 1-D bulk copies.
// Tensor maps are encoded on the host per launch (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: the library
// links no libcuda) and passed as __grid_constant__ parameters.
__device__ __forceinline__ void tensor_store_3d(const CUtensorMap *map, const void *ssrc, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(ssrc)),
                 "r"(x), "r"(y), "r"(z) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tensor_load_3d(void *sdst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(sdst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

template <bool SWAP, int BPP, bool AFIRST, bool LOADS, bool TSTORE, int STAGES>
__global__ void __launch_bounds__(256, 3) k_yuv420_rgb24_tma2d(FastParams p, const __grid_constant__ CUtensorMap mapY,
                                                               const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapV,
                                                               const __grid_constant__ CUtensorMap mapO)
{
    static_assert(!TSTORE || BPP == 3, "tensor stores cannot merge the untouched alpha byte");
    __shared__ uint32_t s_tab[512];
    extern __shared__ __align__(128) uint8_t s_raw[];
    // per warp (bytes): output tiles (two of 2 x 1536 for tensor stores, one 1536-byte transpose buffer otherwise),
    // [STAGES input stages of 1024 (Y) + 256 (U) + 256 (V)], the mbarriers
    constexpr int kOutTile = 3072, kOut = TSTORE ? 2 * kOutTile : 512 * BPP, kInStage = 1536, kPerWarp = kOut + (LOADS ? STAGES * kInStage : 0) + 128;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tab[i] = reinterpret_cast<const uint32_t *>(&g_tabs16)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = s_raw + (size_t)warp * kPerWarp;
    uint8_t *out0 = wbase, *in0 = wbase + kOut;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + kPerWarp - 128);
    const int frame = blockIdx.y;
    const size_t soff = (size_t)frame * p.spitch, doff = (size_t)frame * p.dpitch;
    const uint8_t *Y = p.s0 + soff, *U = p.s1 + soff, *V = p.s2 + soff;
    const int unit = blockIdx.z * blockDim.x + threadIdx.x, wu0 = blockIdx.z * blockDim.x + warp * 32;
    const bool valid = unit < p.upr;
    const int nvalid = min(32, p.upr - wu0);
    if (nvalid <= 0) return;
    if (LOADS && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int rp, int st) {       // lane 0 only
        uint8_t *b = in0 + st * kInStage;
        mbar_expect_tx(&bars[st], kInStage);
        tensor_load_3d(b, &mapY, wu0 * 4, 2 * rp, frame, &bars[st]);            // x in uint32 elements: 16 pixels = 4
        tensor_load_3d(b + 1024, &mapU, wu0 * 2, rp, frame, &bars[st]);
        tensor_load_3d(b + 1280, &mapV, wu0 * 2, rp, frame, &bars[st]);
    };
    const int step = (int)gridDim.x;
    int rp = blockIdx.x;
    if (LOADS && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++)
            if (rp + s * step < p.nrp) issue(rp + s * step, s);
    }
    for (int it = 0; rp < p.nrp; rp += step, it++) {
        uint32_t y0[4] = {0, 0, 0, 0}, y1[4] = {0, 0, 0, 0};
        uint2 uu = make_uint2(0, 0), vv = make_uint2(0, 0);
        if (LOADS) {
            const int st = it % STAGES;
            if (lane == 0 && rp + (STAGES - 1) * step < p.nrp) issue(rp + (STAGES - 1) * step, (it + STAGES - 1) % STAGES);
            mbar_wait(&bars[st], (it / STAGES) & 1);
            const uint8_t *b = in0 + st * kInStage;
            const uint4 a = reinterpret_cast<const uint4 *>(b)[lane], c = reinterpret_cast<const uint4 *>(b + 512)[lane];
            y0[0] = a.x; y0[1] = a.y; y0[2] = a.z; y0[3] = a.w;
            y1[0] = c.x; y1[1] = c.y; y1[2] = c.z; y1[3] = c.w;
            uu = reinterpret_cast<const uint2 *>(b + 1024)[lane];
            vv = reinterpret_cast<const uint2 *>(b + 1280)[lane];
            __syncwarp();        // every lane has read this stage before lane 0 may refill it (STAGES - 1 trips from now)
        } else if (valid) {
            const uint8_t *yp = Y + (size_t)(2 * rp) * p.w + unit * 16;
            const uint4 a = ldg128(yp), c = ldg128(yp + p.w);
            y0[0] = a.x; y0[1] = a.y; y0[2] = a.z; y0[3] = a.w;
            y1[0] = c.x; y1[1] = c.y; y1[2] = c.z; y1[3] = c.w;
            const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
            uu = ldg64(U + co);
            vv = ldg64(V + co);
        }
        int cr[8], cg[8], cb[8];
#pragma unroll
        for (int s = 0; s < 8; s++)
            chroma_terms<S420>(reinterpret_cast<const int2 *>(s_tab), byte_of(s < 4 ? uu.x : uu.y, s & 3),
                               byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
        uint32_t ow[BPP * 4];
        if (TSTORE) {
            uint8_t *tile = out0 + (it & 1) * kOutTile;
            if (lane == 0) bulk_wait_read<1>();      // this tile's previous store has been read out of shared memory
            __syncwarp();
            convert_row<S420, SWAP, BPP, false>(y0, cr, cg, cb, ow);
#pragma unroll
            for (int k = 0; k < 3; k++) reinterpret_cast<uint4 *>(tile + lane * 48)[k] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
            convert_row<S420, SWAP, BPP, false>(y1, cr, cg, cb, ow);
#pragma unroll
            for (int k = 0; k < 3; k++) reinterpret_cast<uint4 *>(tile + 1536 + lane * 48)[k] = make_uint4(ow[4 * k], ow[4 * k + 1], ow[4 * k + 2], ow[4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tensor_store_3d(&mapO, tile, wu0 * 6, 2 * rp, frame);     // x in uint64 elements: 16 pixels = 48 bytes = 6
        } else {
            uint8_t *row0 = p.d0 + doff + ((size_t)(2 * rp) * p.w + (size_t)wu0 * 16) * BPP;
            prefetch_dest_row<BPP>(row0, lane, nvalid);
            prefetch_dest_row<BPP>(row0 + (size_t)p.w * BPP, lane, nvalid);
            convert_row<S420, SWAP, BPP, AFIRST>(y0, cr, cg, cb, ow);
            store_row_rgb<BPP, AFIRST>(reinterpret_cast<uint4 *>(out0), lane, ow, row0, nvalid);
            convert_row<S420, SWAP, BPP, AFIRST>(y1, cr, cg, cb, ow);
            store_row_rgb<BPP, AFIRST>(reinterpret_cast<uint4 *>(out0), lane, ow, row0 + (size_t)p.w * BPP, nvalid);
        }
    }
    if (TSTORE && lane == 0) bulk_wait_all<0>();
}

// The same front end for the other YUV sources (4:2:2, 4:4:4, 4:1:1 planar; YUY2 / UYVY / YVYU packed) -> RGB24 / BGR24:
// a warp takes 512 pixels of TWO consecutive rows per trip (every plane's box is two rows high, so a trip is three tensor
// loads for planar sources and one for packed ones), converts each row as tier 2's linear walk does and stores through
// the tier-2 transpose.  Stage layout: luma 2 x 512 bytes, then U and V with 2 x kCRow bytes each; packed: 2 x 1024 bytes.
template <int SRC> struct RowStage {
    static constexpr bool packed = SrcInfo<SRC>::packed;
    static constexpr int kCRow = SRC == S444 ? 512 : SRC == S411 ? 128 : 256;       // chroma bytes of one plane, one row, one warp
    static constexpr int kBytes = packed ? 2048 : 1024 + 4 * kCRow;
};

template <int SRC, bool SWAP, int STAGES>
__global__ void __launch_bounds__(256, 2) k_yuvrows_rgb24_tma(FastParams p, const __grid_constant__ CUtensorMap mapY,
                                                              const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapV)
{
    constexpr int BPP = 3;
    using SI = SrcInfo<SRC>;
    using RS = RowStage<SRC>;
    __shared__ int2 s_tab[512];
    extern __shared__ __align__(128) uint8_t s_raw[];
    constexpr int kPerWarp = 1536 + STAGES * RS::kBytes + 128;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = reinterpret_cast<const int2 *>(&g_tabs16)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = s_raw + (size_t)warp * kPerWarp;
    uint4 *stage_out = reinterpret_cast<uint4 *>(wbase);
    uint8_t *in0 = wbase + 1536;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + kPerWarp - 128);
    const int frame = blockIdx.y;
    const size_t doff = (size_t)frame * p.dpitch;
    const int wu0 = blockIdx.z * blockDim.x + warp * 32;
    const int nvalid = min(32, p.upr - wu0);
    if (nvalid <= 0) return;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int rp, int st) {       // lane 0 only
        uint8_t *b = in0 + st * RS::kBytes;
        mbar_expect_tx(&bars[st], RS::kBytes);
        if (RS::packed) {
            tensor_load_3d(b, &mapY, wu0 * 8, 2 * rp, frame, &bars[st]);               // 16 pixels = 32 bytes = 8 words
        } else {
            tensor_load_3d(b, &mapY, wu0 * 4, 2 * rp, frame, &bars[st]);
            tensor_load_3d(b + 1024, &mapU, wu0 * (RS::kCRow / 128), 2 * rp, frame, &bars[st]);
            tensor_load_3d(b + 1024 + 2 * RS::kCRow, &mapV, wu0 * (RS::kCRow / 128), 2 * rp, frame, &bars[st]);
        }
    };
    const int step = (int)gridDim.x;
    int rp = blockIdx.x;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++)
            if (rp + s * step < p.nrp) issue(rp + s * step, s);
    }
    for (int it = 0; rp < p.nrp; rp += step, it++) {
        const int st = it % STAGES;
        if (lane == 0 && rp + (STAGES - 1) * step < p.nrp) issue(rp + (STAGES - 1) * step, (it + STAGES - 1) % STAGES);
        mbar_wait(&bars[st], (it / STAGES) & 1);
        const uint8_t *b = in0 + st * RS::kBytes;
        uint32_t yw[2][SI::packed ? 8 : 4], cu[2][4], cv[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            if (RS::packed) {
                const uint4 a = reinterpret_cast<const uint4 *>(b + r * 1024)[2 * lane], c = reinterpret_cast<const uint4 *>(b + r * 1024)[2 * lane + 1];
                yw[r][0] = a.x; yw[r][1] = a.y; yw[r][2] = a.z; yw[r][3] = a.w;
                yw[r][4] = c.x; yw[r][5] = c.y; yw[r][6] = c.z; yw[r][7] = c.w;
            } else {
                const uint4 a = reinterpret_cast<const uint4 *>(b + r * 512)[lane];
                yw[r][0] = a.x; yw[r][1] = a.y; yw[r][2] = a.z; yw[r][3] = a.w;
                const uint8_t *ub = b + 1024 + r * RS::kCRow, *vb = ub + 2 * RS::kCRow;
                if (SRC == S444) {
                    const uint4 u4 = reinterpret_cast<const uint4 *>(ub)[lane], v4 = reinterpret_cast<const uint4 *>(vb)[lane];
                    cu[r][0] = u4.x; cu[r][1] = u4.y; cu[r][2] = u4.z; cu[r][3] = u4.w;
                    cv[r][0] = v4.x; cv[r][1] = v4.y; cv[r][2] = v4.z; cv[r][3] = v4.w;
                } else if (SRC == S422) {
                    const uint2 u2 = reinterpret_cast<const uint2 *>(ub)[lane], v2 = reinterpret_cast<const uint2 *>(vb)[lane];
                    cu[r][0] = u2.x; cu[r][1] = u2.y; cv[r][0] = v2.x; cv[r][1] = v2.y;
                } else {
                    cu[r][0] = reinterpret_cast<const uint32_t *>(ub)[lane];
                    cv[r][0] = reinterpret_cast<const uint32_t *>(vb)[lane];
                }
            }
        }
        __syncwarp();        // every lane has read this stage before lane 0 may refill it (STAGES - 1 trips from now)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            int cr[SI::nchroma], cg[SI::nchroma], cb[SI::nchroma];
#pragma unroll
            for (int s = 0; s < SI::nchroma; s++) {
                uint32_t U, V;
                if (RS::packed) { U = byte_of(yw[r][s], SI::uo); V = byte_of(yw[r][s], SI::vo); }
                else { U = byte_of(cu[r][s >> 2], s & 3); V = byte_of(cv[r][s >> 2], s & 3); }
                chroma_terms<SRC>(s_tab, U, V, cr[s], cg[s], cb[s]);
            }
            uint32_t ow[12];
            convert_row<SRC, SWAP, BPP, false>(yw[r], cr, cg, cb, ow);
            uint8_t *row = p.d0 + doff + ((size_t)(2 * rp + r) * p.w + (size_t)wu0 * 16) * BPP;
            store_row_rgb<BPP, false>(stage_out, lane, ow, row, nvalid);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// [frames][rows][row_bytes / elem] tensor of `elem`-byte unsigned elements, box {box_x, box_y, 1}
static thread_local bool t_encode_failed = false;     // the last tier-3 launch attempt stopped at a tensor-map encode (nothing was launched)
static bool make_map3(CUtensorMap *m, const void *base, int elem, uint64_t row_bytes, uint64_t rows, uint64_t frames, uint64_t pitch,
                      uint32_t box_x, uint32_t box_y)
{
    EncodeTiledFn enc = encode_tiled_fn();
    t_encode_failed = true;
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return false; }
    const CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    const cuuint64_t dims[3] = {row_bytes / (uint64_t)elem, rows, frames}, strides[2] = {row_bytes, frames > 1 ? pitch : row_bytes * rows};
    const cuuint32_t box[3] = {box_x, box_y, 1}, estr[3] = {1, 1, 1};
    const CUresult r = enc(m, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
    t_encode_failed = false;
    return true;
}

template <bool SWAP, int BPP, bool AFIRST, bool LOADS, bool TSTORE, int STAGES>
bool launch_yuv420_rgb24_tma2d(const FastParams &p, int nframes, cudaStream_t st)
{
    CUtensorMap mY, mU, mV, mO;
    const uint64_t w = (uint64_t)p.w, h = (uint64_t)p.nrp * 2, nf = (uint64_t)nframes;
    if (LOADS) {
        if (!make_map3(&mY, p.s0, 4, w, h, nf, p.spitch, 128, 2) || !make_map3(&mU, p.s1, 4, w / 2, h / 2, nf, p.spitch, 64, 1)
            || !make_map3(&mV, p.s2, 4, w / 2, h / 2, nf, p.spitch, 64, 1))
            return false;
    }
    if (TSTORE) {
        if (!make_map3(&mO, p.d0, 8, w * 3, h, nf, p.dpitch, 192, 2)) return false;
    } else {
        mO = mY;                // unused
    }
    if (!LOADS) mY = mU = mV = mO;
    const LaunchShape s = shape_420(p.upr, p.nrp, nframes, 8);
    const size_t smem = (size_t)(s.block.x / 32) * ((TSTORE ? 2 * 3072 : 512 * BPP) + (LOADS ? STAGES * 1536 : 0) + 128) + 128;
    auto kern = k_yuv420_rgb24_tma2d<SWAP, BPP, AFIRST, LOADS, TSTORE, STAGES>;
    if (!check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return false;
    kern<<<s.grid, s.block, smem, st>>>(p, mY, mU, mV, mO);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuv420_rgb24_tma2d");
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Flat form of the tensor-map staged loads (YUV420P -> RGB24 / BGR24): the (row pair, unit) grid is walked as one linear
// sequence, as in the FLAT mode of k_yuv2rgb, so every warp carries 32 real units whatever the width is -- the row-pair
// form above gives a row's last warp whatever is left (1280: 80 units on 96 lanes, 720: 45 on 64) and needs chroma rows
// whose stride is a multiple of 16 bytes (width % 32 == 0).  Per trip a warp issues
//   * the luma box {128 x uint32, 2 rows} at its first unit (clipped by the tensor's row end when the warp runs over it),
//   * for a warp that straddles the end of a row pair, a second luma box at the start of the next row pair,
//   * ONE box of 64 words per chroma plane from a FLAT view of the plane ([frame][w*h/16 x uint32]): the chroma samples of
//     consecutive flat units are consecutive in memory across the end of a chroma row, and a flat view has no row stride
//     to align (PAL: chroma rows of 360 bytes),
// onto the stage's mbarrier; lanes before the split read the first luma tile, the others the second, and the row goes out
// through the two-segment transposed store.
__device__ __forceinline__ void tensor_load_3d_flat(void *sdst, const CUtensorMap *map, int x, int z, uint64_t *bar)
{
    tensor_load_3d(sdst, map, x, 0, z, bar);
}
template <bool SWAP, int STAGES>
__global__ void __launch_bounds__(256, 3) k_yuv420_rgb24_tmaflat(FastParams p, const __grid_constant__ CUtensorMap mapY,
                                                                 const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapV)
{
    constexpr int BPP = 3;
    __shared__ uint32_t s_tab[512];
    extern __shared__ __align__(128) uint8_t s_raw[];
    // per warp (bytes): the 1536-byte transpose buffer, STAGES input stages of 1024 (luma A) + 1024 (luma B) + 256 (U) + 256 (V), the mbarriers
    constexpr int kOut = 512 * BPP, kInStage = 2560, kPerWarp = kOut + STAGES * kInStage + 128;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_tab[i] = reinterpret_cast<const uint32_t *>(&g_tabs16)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = s_raw + (size_t)warp * kPerWarp;
    uint8_t *out0 = wbase, *in0 = wbase + kOut;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + kPerWarp - 128);
    const int frame = blockIdx.y;
    uint8_t *dst = p.d0 + (size_t)frame * p.dpitch;
    const uint32_t upr = (uint32_t)p.upr, total = (uint32_t)p.nrp * upr;
    const uint32_t stride = gridDim.x * blockDim.x, qs = stride / upr, rs = stride - qs * upr;
    uint32_t g0 = blockIdx.x * blockDim.x + warp * 32;
    if (g0 >= total) return;
    uint32_t rpA = g0 / upr, uA = g0 - rpA * upr;            // the trip being converted
    uint32_t gI = g0, rpI = rpA, uI = uA;                    // the trip being requested (STAGES - 1 ahead), lane 0 only
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int st) {       // lane 0 only: requests trip (gI, rpI, uI), then advances the request cursor
        uint8_t *b = in0 + st * kInStage;
        const uint32_t nv = min(32u, total - gI), k = min(nv, upr - uI);
        mbar_expect_tx(&bars[st], k < nv ? 2560u : 1536u);
        tensor_load_3d(b, &mapY, (int)(uI * 4), (int)(2 * rpI), frame, &bars[st]);              // x in uint32 elements: 16 pixels = 4
        if (k < nv) tensor_load_3d(b + 1024, &mapY, 0, (int)(2 * (rpI + 1)), frame, &bars[st]);
        tensor_load_3d_flat(b + 2048, &mapU, (int)(gI * 2), frame, &bars[st]);
        tensor_load_3d_flat(b + 2304, &mapV, (int)(gI * 2), frame, &bars[st]);
        gI += stride; rpI += qs; uI += rs;
        if (uI >= upr) { uI -= upr; rpI++; }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++)
            if (gI < total) issue(s);
    }
    const size_t rowb = (size_t)p.w * BPP;
    for (int it = 0; g0 < total; g0 += stride, it++) {
        const int nvalid = (int)min(32u, total - g0);
        const int k = (int)min((uint32_t)nvalid, upr - uA);           // lanes that still belong to row pair A
        const int st = it % STAGES;
        if (lane == 0 && gI < total) issue((it + STAGES - 1) % STAGES);
        mbar_wait(&bars[st], (it / STAGES) & 1);
        const uint8_t *b = in0 + st * kInStage;
        const uint8_t *yt = lane < k ? b + lane * 16 : b + 1024 + (lane - k) * 16;
        const uint4 a = *reinterpret_cast<const uint4 *>(yt), c = *reinterpret_cast<const uint4 *>(yt + 512);
        const uint2 uu = reinterpret_cast<const uint2 *>(b + 2048)[lane], vv = reinterpret_cast<const uint2 *>(b + 2304)[lane];
        __syncwarp();        // every lane has read this stage before lane 0 may refill it (STAGES - 1 trips from now)
        const uint32_t y0[4] = {a.x, a.y, a.z, a.w}, y1[4] = {c.x, c.y, c.z, c.w};
        int cr[8], cg[8], cb[8];
#pragma unroll
        for (int s = 0; s < 8; s++)
            chroma_terms<S420>(reinterpret_cast<const int2 *>(s_tab), byte_of(s < 4 ? uu.x : uu.y, s & 3),
                               byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
        uint8_t *segA = dst + ((size_t)(2 * rpA) * p.w + (size_t)uA * 16) * BPP;
        uint8_t *segB = dst + (size_t)(2 * (rpA + 1)) * rowb;
        uint32_t ow[BPP * 4];
        if (k == nvalid) {          // warp-uniform: a warp inside one row pair stores one segment (no per-chunk choice of base)
            convert_row<S420, SWAP, BPP, false>(y0, cr, cg, cb, ow);
            store_row_rgb<BPP, false>(reinterpret_cast<uint4 *>(out0), lane, ow, segA, nvalid);
            convert_row<S420, SWAP, BPP, false>(y1, cr, cg, cb, ow);
            store_row_rgb<BPP, false>(reinterpret_cast<uint4 *>(out0), lane, ow, segA + rowb, nvalid);
        } else {
            convert_row<S420, SWAP, BPP, false>(y0, cr, cg, cb, ow);
            store_row_rgb2<BPP, false>(reinterpret_cast<uint4 *>(out0), lane, ow, segA, segB, k, nvalid);
            convert_row<S420, SWAP, BPP, false>(y1, cr, cg, cb, ow);
            store_row_rgb2<BPP, false>(reinterpret_cast<uint4 *>(out0), lane, ow, segA + rowb, segB + rowb, k, nvalid);
        }
        rpA += qs;
        uA += rs;
        if (uA >= upr) { uA -= upr; rpA++; }
    }
}

template <bool SWAP>
bool launch_yuv420_rgb24_tmaflat(const FastParams &p, int nframes, cudaStream_t st)
{
    constexpr int STAGES = 2;       // 2 / 3 / 4 stages: 0.92 / 0.90 / 0.88 of the copy rate (the shared memory a stage costs is occupancy)
    CUtensorMap mY, mU, mV;
    const uint64_t w = (uint64_t)p.w, h = (uint64_t)p.nrp * 2, nf = (uint64_t)nframes, cbytes = w * h / 4;
    if (!make_map3(&mY, p.s0, 4, w, h, nf, p.spitch, 128, 2) || !make_map3(&mU, p.s1, 4, cbytes, 1, nf, p.spitch, 64, 1)
        || !make_map3(&mV, p.s2, 4, cbytes, 1, nf, p.spitch, 64, 1))
        return false;
    // 6.6 KB of stages per warp; blocks of four warps measured best (96 / 128 / 160 threads: 0.90 / 0.92 / 0.88)
    static const int threads = [] { const char *e = getenv("ACGPU_TMA_FLAT_BLOCK"); const int v = e ? atoi(e) : 128; return v >= 32 && v <= 256 ? v / 32 * 32 : 128; }();
    LaunchShape s;
    s.block = dim3((unsigned)threads);
    const long total = (long)p.upr * p.nrp, maxgx = (total + threads - 1) / threads;
    long gx = ((long)sm_count() * (640 / threads) * waves(8) + nframes - 1) / nframes;
    if (gx < 1) gx = 1;
    if (gx > maxgx) gx = maxgx;
    s.grid = dim3((unsigned)gx, (unsigned)nframes);
    const size_t smem = (size_t)(s.block.x / 32) * (1536 + STAGES * 2560 + 128) + 128;
    auto kern = k_yuv420_rgb24_tmaflat<SWAP, STAGES>;
    if (!check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return false;
    kern<<<s.grid, s.block, smem, st>>>(p, mY, mU, mV);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuv420_rgb24_tmaflat");
    return true;
}

template <bool SWAP>
bool launch_tma2d_mode(int mode, const FastParams &p, int nframes, cudaStream_t st)
{
    switch (mode) {
    case 3: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, false, true, 2>(p, nframes, st);     // LDG loads, tensor stores
    case 4: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, true, true, 2>(p, nframes, st);      // tensor loads (2 stages), tensor stores
    case 5: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, true, false, 2>(p, nframes, st);     // tensor loads, LDS + STG stores
    case 6: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, true, true, 3>(p, nframes, st);      // tensor loads (3 stages), tensor stores
    case 8: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, true, false, 4>(p, nframes, st);     // tensor loads (4 stages), LDS + STG stores
    default: return launch_yuv420_rgb24_tma2d<SWAP, 3, false, true, false, 3>(p, nframes, st);    // 7: tensor loads (3 stages), LDS + STG stores
    }
}

// The automatic form: three-stage tensor-map loads in front of the tier-2 stores, for the 24-bit destinations.
bool tma_loads_dst(int dstfmt, const FastParams &p, int nframes, cudaStream_t st)
{
    switch (dstfmt) {
    case IMG_RGB24:  return launch_yuv420_rgb24_tma2d<false, 3, false, true, false, 3>(p, nframes, st);
    case IMG_BGR24:  return launch_yuv420_rgb24_tma2d<true, 3, false, true, false, 3>(p, nframes, st);
    default: return false;       // 32-bit destinations measured slower this way (read-modify-write of alpha: 0.987 -> 0.938)
    }
}

template <int SRC>
bool dispatch_yuv2rgb_bulk(int dstfmt, const FastParams &p, int nframes, cudaStream_t st)
{
    switch (dstfmt) {
    case IMG_RGB24: return launch_yuv2rgb<SRC, false, 3, false, true>(p, nframes, st);
    case IMG_BGR24: return launch_yuv2rgb<SRC, true, 3, false, true>(p, nframes, st);
    default: return false;
    }
}

bool fast_yuv2rgb(const ConvertArgs &a, const FastParams &p)
{
    switch (a.srcfmt) {
    case IMG_YUV420P: return p.ragged420 ? dispatch_yuv2rgb_dst<S420R>(a.dstfmt, p, a.nframes, a.stream)
                                         : dispatch_yuv2rgb_dst<S420>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV422P: return dispatch_yuv2rgb_dst<S422>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV411P: return dispatch_yuv2rgb_dst<S411>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV444P: return dispatch_yuv2rgb_dst<S444>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUY2:    return dispatch_yuv2rgb_dst<SYUY2>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_UYVY:    return dispatch_yuv2rgb_dst<SUYVY>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YVYU:    return dispatch_yuv2rgb_dst<SYVYU>(a.dstfmt, p, a.nframes, a.stream);
    default: return false;
    }
}

template <int SL, int DST>
bool launch_rgb2yuv(const FastParams &p, int nframes, cudaStream_t st)
{
    LaunchShape s = DST == D420 ? shape_420(p.upr, p.nrp, nframes) : shape_linear(p.nunits, nframes);
    const size_t smem = (size_t)(s.block.x / 32) * 64 * sizeof(uint4);
    k_rgb2yuv<SL, DST><<<s.grid, s.block, smem, st>>>(p);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_rgb2yuv");
    return true;
}

template <int SL>
bool dispatch_rgb2yuv_dst(int dstfmt, const FastParams &p, int nframes, cudaStream_t st)
{
    switch (dstfmt) {
    case IMG_YUV420P: return p.ragged420 ? launch_rgb2yuv<SL, D420R>(p, nframes, st) : launch_rgb2yuv<SL, D420>(p, nframes, st);
    case IMG_YUV422P: return launch_rgb2yuv<SL, D422>(p, nframes, st);
    case IMG_YUV411P: return launch_rgb2yuv<SL, D411>(p, nframes, st);
    case IMG_YUV444P: return launch_rgb2yuv<SL, D444>(p, nframes, st);
    case IMG_YUY2:    return launch_rgb2yuv<SL, DYUY2>(p, nframes, st);
    case IMG_UYVY:    return launch_rgb2yuv<SL, DUYVY>(p, nframes, st);
    case IMG_YVYU:    return launch_rgb2yuv<SL, DYVYU>(p, nframes, st);
    case IMG_Y8:      return launch_rgb2yuv<SL, DY8>(p, nframes, st);
    default: return false;
    }
}

bool fast_rgb2yuv(const ConvertArgs &a, const FastParams &p)
{
    switch (a.srcfmt) {
    case IMG_RGB24:  return dispatch_rgb2yuv_dst<L_RGB24>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_BGR24:  return dispatch_rgb2yuv_dst<L_BGR24>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_RGBA32: return dispatch_rgb2yuv_dst<L_RGBA>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_BGRA32: return dispatch_rgb2yuv_dst<L_BGRA>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_ARGB32: return dispatch_rgb2yuv_dst<L_ARGB>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_ABGR32: return dispatch_rgb2yuv_dst<L_ABGR>(a.dstfmt, p, a.nframes, a.stream);
    default: return false;
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused YUV420P -> RGB -> YUV (frame chains, BASELINE config 4: YUV420P -> RGB24 -> YUV422P).  Two conversions through an
// RGB frame that nobody looks at are one pass: the 8-bit RGB values of a pixel are formed in registers exactly as
// k_yuv2rgb forms them (same table terms, same clamp; convert_row in its 32-bit layout gives one word per pixel) and fed
// straight to the dp2a accumulators of k_rgb2yuv.  The intermediate's byte order does not matter -- the second conversion
// reads the channels it was written with -- so every RGB layout fuses to the same code.  Bit-identical to the two calls;
// 3.5 bytes of HBM traffic per pixel instead of 9.5.
// Form of the arithmetic (ACGPU_FUSED_IMAD, compile time, default 0).  0: as the two kernels do it -- the channel bytes are
// packed into one word per pixel (two PRMT) and every component is two dp2a on that word; with the yuv->rgb steps that is ~17
// of 21 instructions per pixel on the ALU pipe (IDP, VIADDMNMX, PRMT), which runs at 65 % with the multiplier pipe at 35 %
// (profiles/r2_ncu_summaries.md).  1: the channel comes out of the scale step as an integer -- one multiply-high with a 64-bit
// addend gives floor((jc * 1220944 + 2^23) / 2^24) -- and every component is three multiply-adds on those integers: no packing,
// no dp2a, 10 of 21 instructions per pixel on each pipe.  MEASURED AND REJECTED: bit-identical, but 128.8 k -> 106.9 k UHD
// frames/s (0.57 -> 0.47 of the copy rate): the integer multiplier is the narrower pipe on this part, the balanced mix is the
// slower one (profiles/r2_experiments.md section 3b).
#ifndef ACGPU_FUSED_IMAD
#define ACGPU_FUSED_IMAD 0
#endif
template <int BIAS>
__device__ __forceinline__ uint32_t channel_int(uint32_t yword, uint32_t ysel, int c)
{
    const int j = (int)__dp4a(yword, ysel, (uint32_t)c);
    const int jc = BIAS ? __viaddmin_s32_relu(j, -BIAS, pixmath::kJMax) : __vimin_s32_relu(j, pixmath::kJMax);
    return (uint32_t)(((uint64_t)(uint32_t)jc * (uint64_t)(pixmath::kJMul * 256u) + ((uint64_t)pixmath::kJAdd << 8)) >> 32);
}
// accumulators whose byte 2 is the component (img_yuv_rgb.c:142-147).  k0 = the rounding constant plus the +16 / +128, held
// in a register the compiler cannot see through (opaque_const): as an immediate it takes the addend slot of the first
// multiply-add, the coefficient then needs a register that the result overwrites, and every chain re-loads it (60 MOV per
// trip).
__device__ __forceinline__ uint32_t opaque_const(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t iacc_y(uint32_t r, uint32_t g, uint32_t b, uint32_t k0) { return r * 16829u + (g * 33039u + (b * 6416u + k0)); }
__device__ __forceinline__ uint32_t iacc_u(uint32_t r, uint32_t g, uint32_t b, uint32_t k0)
{
    return (uint32_t)((int)r * -9714 + ((int)g * -19070 + ((int)b * 28784 + (int)k0)));
}
__device__ __forceinline__ uint32_t iacc_v(uint32_t r, uint32_t g, uint32_t b, uint32_t k0)
{
    return (uint32_t)((int)r * 28784 + ((int)g * -24103 + ((int)b * -4681 + (int)k0)));
}

// One unit (16 pixels of both rows of a pair) of the fused pair: luma words a / b, chroma words uu / vv -> stores.
template <int DST>
__device__ __forceinline__ void fused_unit(const FastParams &p, const int2 *s_tab, const uint4 &a, const uint4 &b, const uint2 &uu,
                                           const uint2 &vv, uint8_t *Y, uint8_t *U, uint8_t *V, int rp, int unit, uint32_t ky, uint32_t kc)
{
    constexpr int SL = L_RGBA;          // convert_row<..., 4, false>: R, G, B in bytes 0..2 of each pixel word
    (void)SL; (void)ky; (void)kc;
    int cr[8], cg[8], cb[8];
#pragma unroll
    for (int s = 0; s < 8; s++)
        chroma_terms<S420>(s_tab, byte_of(s < 4 ? uu.x : uu.y, s & 3), byte_of(s < 4 ? vv.x : vv.y, s & 3), cr[s], cg[s], cb[s]);
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint32_t yw[4] = {r ? b.x : a.x, r ? b.y : a.y, r ? b.z : a.z, r ? b.w : a.w};
        uint32_t ay[16], ua[16], va[16];       // component accumulators; only the sampled positions are computed
        (void)ua; (void)va;
#if ACGPU_FUSED_IMAD
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t sel = 0x10u << (8 * (k & 3));
            const uint32_t R = channel_int<kBiasRB>(yw[k >> 2], sel, cr[k >> 1]);
            const uint32_t G = channel_int<kBiasG>(yw[k >> 2], sel, cg[k >> 1]);
            const uint32_t B = channel_int<kBiasRB>(yw[k >> 2], sel, cb[k >> 1]);
            ay[k] = iacc_y(R, G, B, ky);
            if (DST == D444) { ua[k] = iacc_u(R, G, B, kc); va[k] = iacc_v(R, G, B, kc); }
            else if (DST == D422) { if (k & 1) va[k] = iacc_v(R, G, B, kc); else ua[k] = iacc_u(R, G, B, kc); }
            else if (r == 0 && !(k & 1)) ua[k] = iacc_u(R, G, B, kc);        // D420: U at (even x, even y)
            else if (r == 1 && (k & 1)) va[k] = iacc_v(R, G, B, kc);         //       V at (odd x, odd y)
        }
#else
        uint32_t px[16];
        convert_row<S420, false, 4, false>(yw, cr, cg, cb, px);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            ay[k] = acc_y<SL>(px[k]);
            if (DST == D444) { ua[k] = acc_u<SL>(px[k]); va[k] = acc_v<SL>(px[k]); }
            else if (DST == D422) { if (k & 1) va[k] = acc_v<SL>(px[k]); else ua[k] = acc_u<SL>(px[k]); }
            else if (r == 0 && !(k & 1)) ua[k] = acc_u<SL>(px[k]);
            else if (r == 1 && (k & 1)) va[k] = acc_v<SL>(px[k]);
        }
#endif
        const size_t row = (size_t)(2 * rp + r);
        stg128(Y + row * p.w + unit * 16,
               make_uint4(pack_b2x4(ay[0], ay[1], ay[2], ay[3]), pack_b2x4(ay[4], ay[5], ay[6], ay[7]),
                          pack_b2x4(ay[8], ay[9], ay[10], ay[11]), pack_b2x4(ay[12], ay[13], ay[14], ay[15])));
        if (DST == D422) {          // U at even x, V at odd x of every row (img_yuv_rgb.c:166)
            stg64(U + row * (p.w >> 1) + unit * 8, make_uint2(pack_b2x4(ua[0], ua[2], ua[4], ua[6]), pack_b2x4(ua[8], ua[10], ua[12], ua[14])));
            stg64(V + row * (p.w >> 1) + unit * 8, make_uint2(pack_b2x4(va[1], va[3], va[5], va[7]), pack_b2x4(va[9], va[11], va[13], va[15])));
        } else if (DST == D420) {   // U at (even x, even y), V at (odd x, odd y) (:162)
            if (r == 0)
                stg64(U + (size_t)rp * (p.w >> 1) + unit * 8, make_uint2(pack_b2x4(ua[0], ua[2], ua[4], ua[6]), pack_b2x4(ua[8], ua[10], ua[12], ua[14])));
            else
                stg64(V + (size_t)rp * (p.w >> 1) + unit * 8, make_uint2(pack_b2x4(va[1], va[3], va[5], va[7]), pack_b2x4(va[9], va[11], va[13], va[15])));
        } else {                    // D444: every pixel
            stg128(U + row * p.w + unit * 16,
                   make_uint4(pack_b2x4(ua[0], ua[1], ua[2], ua[3]), pack_b2x4(ua[4], ua[5], ua[6], ua[7]),
                              pack_b2x4(ua[8], ua[9], ua[10], ua[11]), pack_b2x4(ua[12], ua[13], ua[14], ua[15])));
            stg128(V + row * p.w + unit * 16,
                   make_uint4(pack_b2x4(va[0], va[1], va[2], va[3]), pack_b2x4(va[4], va[5], va[6], va[7]),
                              pack_b2x4(va[8], va[9], va[10], va[11]), pack_b2x4(va[12], va[13], va[14], va[15])));
        }
    }
}

template <int DST>
__global__ void __launch_bounds__(256, 4) k_yuv420_rgb_yuv(FastParams p)
{
    __shared__ int2 s_tab[512];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = reinterpret_cast<const int2 *>(&g_tabs16)[i];
    __syncthreads();
    const size_t soff = (size_t)blockIdx.y * p.spitch, doff = (size_t)blockIdx.y * p.dpitch;
    const uint8_t *Ys = p.s0 + soff, *Us = p.s1 + soff, *Vs = p.s2 + soff;
    uint8_t *Y = p.d0 + doff, *U = p.d1 + doff, *V = p.d2 + doff;
    const int unit = blockIdx.z * blockDim.x + threadIdx.x;
    if (unit >= p.upr) return;
    const uint32_t ky = opaque_const(32768u + (16u << 16)), kc = opaque_const(32768u + (128u << 16));
    for (int rp = blockIdx.x; rp < p.nrp; rp += gridDim.x) {
        const uint8_t *yp = Ys + (size_t)(2 * rp) * p.w + unit * 16;
        const uint4 a = ldg128(yp), b = ldg128(yp + p.w);
        const size_t co = (size_t)rp * (p.w >> 1) + unit * 8;
        const uint2 uu = ldg64(Us + co), vv = ldg64(Vs + co);
        fused_unit<DST>(p, s_tab, a, b, uu, vv, Y, U, V, rp, unit, ky, kc);
    }
}

}  // namespace

// Builds the launch parameters if the call is inside the vectorised tiers' domain.
static bool fast_domain(const ConvertArgs &a, FastParams *out)
{
    const FmtDesc sd = describe(a.srcfmt), dd = describe(a.dstfmt);
    const int w = a.w, h = a.h;
    if (w <= 0 || h <= 0 || a.nframes <= 0) return false;
    if (a.nframes > 1 && (a.src.pitch % 16 || a.dst.pitch % 16)) return false;
    const bool any420 = a.srcfmt == IMG_YUV420P || a.dstfmt == IMG_YUV420P;
    const size_t P = (size_t)w * h;
    // every plane 16-byte aligned -- except the chroma planes of a ragged 4:2:0 frame, which are reached through
    // byte-aligned accesses anyway (e.g. 50x16: V starts at byte 1000 of the frame)
    // (the same kernels take an aligned width whose 4:2:0 chroma planes are not 16-byte aligned: width*height % 64 != 0)
    const bool chroma_off = (a.srcfmt == IMG_YUV420P && !(al16(a.src.p[1]) && al16(a.src.p[2])))
                         || (a.dstfmt == IMG_YUV420P && !(al16(a.dst.p[1]) && al16(a.dst.p[2])));
    const bool ragged_w = any420 && (w % 16 || chroma_off);
    for (int i = 0; i < 3; i++) {
        const bool s_free = ragged_w && i > 0 && a.srcfmt == IMG_YUV420P, d_free = ragged_w && i > 0 && a.dstfmt == IMG_YUV420P;
        if (a.src.p[i] && (s_free ? ((uintptr_t)a.src.p[i] & 3) != 0 : !al16(a.src.p[i]))) return false;
        if (a.dst.p[i] && (d_free ? ((uintptr_t)a.dst.p[i] & 3) != 0 : !al16(a.dst.p[i]))) return false;
    }
    bool ragged = false;
    if (any420) {
        if (w % 2 || h % 2) return false;
        if (ragged_w) {
            // ragged 4:2:0 rows: flat-unit kernels (S420R / D420R in this file, Ragged420From / Ragged420To in
            // kernels_fast_yuv.cu); 4:1:1 needs whole 4-pixel groups per row
            const int other = a.srcfmt == IMG_YUV420P ? a.dstfmt : a.srcfmt;
            if (w < 16 || P % 16 || P >= 0x7FFFFFFFu) return false;
            if (other == IMG_YUV411P && w % 4) return false;
            ragged = true;
        } else {
            // The flat one-row-per-unit walk also serves aligned widths.  RGB -> 4:2:0 prefers it at every size measured
            // (no idle lanes when width/16 is not a multiple of 32: PAL 0.98 -> 1.07 of peak, 720p 1.03 -> 1.06, 1080p
            // 1.06 -> 1.07); 4:2:0 -> RGB and the YUV family lose 5-25 % to the repeated chroma work and stay on the
            // row-pair kernels.  $ACGPU_FORCE_RAGGED = 1 / 0 forces it on / off for every 4:2:0 pair (profiling).
            static const int force = [] { const char *e = getenv("ACGPU_FORCE_RAGGED"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
            const bool rgb_to_420 = sd.kind == K_RGB && a.dstfmt == IMG_YUV420P;
            if (P < 0x7FFFFFFFu && (force == 1 || (force == -1 && rgb_to_420))) ragged = true;
        }
    } else {
        if (P % 16) return false;
        if ((a.srcfmt == IMG_YUV411P || a.dstfmt == IMG_YUV411P) && w % 4) return false;
        if ((sd.kind == K_PACKED || dd.kind == K_PACKED || a.srcfmt == IMG_YUV422P || a.dstfmt == IMG_YUV422P) && w % 2) return false;
    }
    if (P / 16 > 0x7FFFFFFFu) return false;
    FastParams p{};
    p.s0 = a.src.p[0]; p.s1 = a.src.p[1]; p.s2 = a.src.p[2];
    p.d0 = a.dst.p[0]; p.d1 = a.dst.p[1]; p.d2 = a.dst.p[2];
    p.spitch = a.src.pitch; p.dpitch = a.dst.pitch;
    p.w = w; p.h = h;
    p.upr = w / 16; p.nrp = h / 2;
    p.nunits = (uint32_t)(P / 16);
    p.ragged420 = ragged;
    *out = p;
    return true;
}

bool convert_fast(const ConvertArgs &a)
{
    FastParams p;
    if (!fast_domain(a, &p)) return false;
    const FmtDesc sd = describe(a.srcfmt), dd = describe(a.dstfmt);
    if ((sd.kind == K_PLANAR || sd.kind == K_PACKED) && dd.kind == K_RGB) return fast_yuv2rgb(a, p);
    if (sd.kind == K_RGB && (dd.kind == K_PLANAR || dd.kind == K_PACKED || dd.kind == K_Y8)) return fast_rgb2yuv(a, p);
    const bool s_yuvish = sd.kind == K_PLANAR || sd.kind == K_PACKED || sd.kind == K_Y8 || sd.kind == K_GRAY;
    const bool d_yuvish = dd.kind == K_PLANAR || dd.kind == K_PACKED || dd.kind == K_Y8 || dd.kind == K_GRAY;
    if (s_yuvish && d_yuvish) return fast_yuv_family(a, p);
    return fast_rgb_family(a, p);
}

// Tier 3: the same arithmetic, but 24-bit RGB tiles leave shared memory through bulk (TMA) stores.
template <int SRC, bool SWAP>
bool launch_yuvrows_rgb24_tma(const FastParams &p0, int nframes, cudaStream_t st)
{
    using RS = RowStage<SRC>;
    constexpr int STAGES = 3;
    FastParams p = p0;
    p.upr = p.w / 16;
    p.nrp = p.h / 2;
    CUtensorMap mY, mU, mV;
    const uint64_t w = (uint64_t)p.w, h = (uint64_t)p.h, nf = (uint64_t)nframes;
    if (RS::packed) {
        if (!make_map3(&mY, p.s0, 4, w * 2, h, nf, p.spitch, 256, 2)) return false;
        mU = mV = mY;
    } else {
        const uint64_t crow = SRC == S444 ? w : SRC == S411 ? w / 4 : w / 2;
        if (!make_map3(&mY, p.s0, 4, w, h, nf, p.spitch, 128, 2) || !make_map3(&mU, p.s1, 4, crow, h, nf, p.spitch, RS::kCRow / 4, 2)
            || !make_map3(&mV, p.s2, 4, crow, h, nf, p.spitch, RS::kCRow / 4, 2))
            return false;
    }
    const LaunchShape s = shape_420(p.upr, p.nrp, nframes, 8);
    const size_t smem = (size_t)(s.block.x / 32) * (1536 + STAGES * RS::kBytes + 128) + 128;
    auto kern = k_yuvrows_rgb24_tma<SRC, SWAP, STAGES>;
    if (!check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr")) return false;
    kern<<<s.grid, s.block, smem, st>>>(p, mY, mU, mV);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuvrows_rgb24_tma");
    return true;
}

template <int SRC>
bool tma_rows_dst(int dstfmt, const FastParams &p, int nframes, cudaStream_t st)
{
    return dstfmt == IMG_BGR24 ? launch_yuvrows_rgb24_tma<SRC, true>(p, nframes, st) : launch_yuvrows_rgb24_tma<SRC, false>(p, nframes, st);
}

// Automatic tier 3: YUV420P -> RGB24 / BGR24 with the planes staged by tensor-map loads (three stages per warp) in front
// of the tier-2 transposed stores.  Measured against tier 2 in the same run (profiles/r2_tma_tensor_maps.md): 1080p 0.934 ->
// 0.975 of the copy peak, UHD 0.926 -> 0.958, 720p 0.847 -> 0.879; narrow frames whose rows leave more than a fifth of a
// block's lanes idle (640 wide: 0.841 -> 0.68) stay on tier 2's flat walk.  Tensor maps need 16-byte strides: w % 32 == 0.
bool convert_tma_auto(const ConvertArgs &a)
{
    // bit 0: 4:2:0; bit 1: 4:1:1 and wide 4:2:2 (the sources the row form wins on); bit 2: every other YUV source (measured
    // slower than tier 2: selectable for experiments and parity tests only)
    static const int enabled = [] { const char *e = getenv("ACGPU_TMA_AUTO"); return e ? atoi(e) : 3; }();
    if (!enabled || (a.dstfmt != IMG_RGB24 && a.dstfmt != IMG_BGR24)) return false;
    const FmtDesc sd = describe(a.srcfmt);
    if (sd.kind != K_PLANAR && sd.kind != K_PACKED) return false;
    FastParams p;
    if (!fast_domain(a, &p) || p.ragged420 || a.h % 2) return false;
    if (!encode_tiled_fn()) return false;
    // a geometry the tensor-map encoder refuses is not an error of the call: nothing was launched, tier 2 takes it
    auto soft = [](bool launched) {
        if (!launched && t_encode_failed) { t_encode_failed = false; set_error("%s", ""); }
        return launched;
    };
    const int upr = a.w / 16, lanes_row = ((upr + 31) / 32) * 32;
    // YUV420P whose rows leave more than a tenth of the row-pair form's lanes idle, or whose chroma rows that form cannot
    // describe (width % 32 != 0): the flat form -- 0.92-0.95 of the copy rate at every width measured (PAL 0.83 -> 0.92, 1280x720
    // 0.87 -> 0.95, 640x480 / 1600x900 0.83 -> 0.93 / 0.94), where the row-pair form reaches 0.96-0.99 on full warps
    // (1920, 2560, 3840).  $ACGPU_TMA_FLAT: 0 never, 1 (default) that rule, 2 every width it can take.
    static const int flat_mode = [] { const char *e = getenv("ACGPU_TMA_FLAT"); return e ? atoi(e) : 1; }();
    if (a.srcfmt == IMG_YUV420P && (enabled & 1) && flat_mode && upr >= 32 && (uint64_t)upr * (uint64_t)(a.h / 2) < 0x7FFFFFFFu
        && (flat_mode >= 2 || a.w % 32 != 0 || upr * 10 < lanes_row * 9)) {
        return a.dstfmt == IMG_BGR24 ? soft(launch_yuv420_rgb24_tmaflat<true>(p, a.nframes, a.stream))
                                     : soft(launch_yuv420_rgb24_tmaflat<false>(p, a.nframes, a.stream));
    }
    // tensor-map strides are multiples of 16 bytes: the narrowest plane's rows decide
    const int wmod = a.srcfmt == IMG_YUV411P ? 64 : sd.kind == K_PACKED || a.srcfmt == IMG_YUV444P ? 16 : 32;
    if (a.w % wmod) return false;
    if (upr * 10 < lanes_row * 8 || upr > 256 * 65535) return false;
    if (a.srcfmt == IMG_YUV420P) return (enabled & 1) && soft(tma_loads_dst(a.dstfmt, p, a.nframes, a.stream));
    const bool wins = a.srcfmt == IMG_YUV411P || (a.srcfmt == IMG_YUV422P && a.w >= 1920);
    if (!(enabled & (wins ? 2 : 4))) return false;
    switch (a.srcfmt) {
    case IMG_YUV422P: return soft(tma_rows_dst<S422>(a.dstfmt, p, a.nframes, a.stream));
    case IMG_YUV444P: return soft(tma_rows_dst<S444>(a.dstfmt, p, a.nframes, a.stream));
    case IMG_YUV411P: return soft(tma_rows_dst<S411>(a.dstfmt, p, a.nframes, a.stream));
    case IMG_YUY2:    return soft(tma_rows_dst<SYUY2>(a.dstfmt, p, a.nframes, a.stream));
    case IMG_UYVY:    return soft(tma_rows_dst<SUYVY>(a.dstfmt, p, a.nframes, a.stream));
    case IMG_YVYU:    return soft(tma_rows_dst<SYVYU>(a.dstfmt, p, a.nframes, a.stream));
    default: return false;
    }
}

bool convert_fused_yuv420_rgb_yuv(const ConvertArgs &a)
{
    if (a.srcfmt != IMG_YUV420P || (a.dstfmt != IMG_YUV422P && a.dstfmt != IMG_YUV420P && a.dstfmt != IMG_YUV444P)) return false;
    FastParams p;
    if (!fast_domain(a, &p) || p.ragged420 || a.h % 2) return false;
    p.upr = a.w / 16;
    p.nrp = a.h / 2;
    const LaunchShape s = shape_420(p.upr, p.nrp, a.nframes, 8);
    switch (a.dstfmt) {
    case IMG_YUV422P: k_yuv420_rgb_yuv<D422><<<s.grid, s.block, 0, a.stream>>>(p); break;
    case IMG_YUV420P: k_yuv420_rgb_yuv<D420><<<s.grid, s.block, 0, a.stream>>>(p); break;
    default:          k_yuv420_rgb_yuv<D444><<<s.grid, s.block, 0, a.stream>>>(p); break;
    }
    note_launch();
    ACGPU_CHECK_LAUNCH("k_yuv420_rgb_yuv");
    return true;
}

bool convert_tma(const ConvertArgs &a)
{
    if (a.dstfmt != IMG_RGB24 && a.dstfmt != IMG_BGR24) return false;
    FastParams p;
    if (!fast_domain(a, &p) || p.ragged420) return false;
    // $ACGPU_TMA: 0 = bulk stores only (default tier 3), 1 = bulk-async staged loads + LDS/STG stores,
    //             2 = bulk-async staged loads + bulk stores, 3 = 2-D tensor-map stores, 4 = tensor-map loads and stores,
    //             5 = tensor-map loads + LDS/STG stores, 6 / 7 = as 4 / 5 with a three-stage load pipeline, 8 = as 7 with four.
    //             Loads need an even number of units per warp (w % 32 == 0).
    static const int mode = [] { const char *e = getenv("ACGPU_TMA"); return e ? atoi(e) : 0; }();
    if (mode >= 3 && a.srcfmt == IMG_YUV420P && a.w % 32 == 0 && a.h % 2 == 0 && p.upr <= 256 * 65535) {
        return a.dstfmt == IMG_BGR24 ? launch_tma2d_mode<true>(mode, p, a.nframes, a.stream) : launch_tma2d_mode<false>(mode, p, a.nframes, a.stream);
    }
    if (mode > 0 && a.srcfmt == IMG_YUV420P && a.w % 32 == 0) {
        const bool swap = a.dstfmt == IMG_BGR24;
        if (mode == 1) return swap ? launch_yuv420_rgb24_tma<true, false>(p, a.nframes, a.stream) : launch_yuv420_rgb24_tma<false, false>(p, a.nframes, a.stream);
        return swap ? launch_yuv420_rgb24_tma<true, true>(p, a.nframes, a.stream) : launch_yuv420_rgb24_tma<false, true>(p, a.nframes, a.stream);
    }
    switch (a.srcfmt) {
    case IMG_YUV420P: return dispatch_yuv2rgb_bulk<S420>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV422P: return dispatch_yuv2rgb_bulk<S422>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV411P: return dispatch_yuv2rgb_bulk<S411>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUV444P: return dispatch_yuv2rgb_bulk<S444>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YUY2:    return dispatch_yuv2rgb_bulk<SYUY2>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_UYVY:    return dispatch_yuv2rgb_bulk<SUYVY>(a.dstfmt, p, a.nframes, a.stream);
    case IMG_YVYU:    return dispatch_yuv2rgb_bulk<SYVYU>(a.dstfmt, p, a.nframes, a.stream);
    default: return false;
    }
}

}  // namespace acgpu
