// rowops.cu -- ac_average / ac_rescale and the libtcvideo row shapes built on them.
//
// Arithmetic (pixmath.cuh): average (a+b+1)/2 (aclib/average.c:37-38); rescale
// ((a*w1 + b*w2 + 32768) >> 16) & 0xFF in uint32 with the copy shortcuts of aclib/rescale.c:26-31
// (weight1 >= 65536 copies src1 and NEVER reads src2 -- libtcvideo points src2 one row past the
// frame in that case, libtcvideo/tcvideo.c:469-470).
//
// Bandwidth shape: 2 B read + 1 B written per blended byte, 1 + 1 per copied byte.  One launch covers
// every row of every frame of a batch (grid.y = frame); each thread moves one 16-byte chunk when the
// row geometry is 16-byte aligned, bytes otherwise.
#include "acgpu_internal.h"

#include <algorithm>
#include "pixmath.cuh"

#include <string.h>

#include <vector>

namespace acgpu {
namespace {

constexpr int TPB = 256;

__device__ __forceinline__ uint32_t avg4x8(uint32_t a, uint32_t b)
{
    // per-byte (a + b + 1) >> 1 without carries between lanes
    return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu);
}

// Four bytes of (a*w1 + b*w2 + 32768) >> 16, low byte kept (aclib/rescale.c:44-45).  Both weights are < 65536 on
// this path, so one dp2a (16-bit weights x 8-bit samples, SASS IDP.2A) forms a whole sum; PRMT pairs the samples
// and picks byte 2 of each accumulator: 2.25 instructions per output byte.
__device__ __forceinline__ uint32_t rescale4x8(uint32_t a, uint32_t b, uint32_t w1, uint32_t w2)
{
    const uint32_t w = (w1 & 0xFFFFu) | (w2 << 16);
    const uint32_t p01 = __byte_perm(a, b, 0x5140);   // a0 b0 a1 b1
    const uint32_t p23 = __byte_perm(a, b, 0x7362);   // a2 b2 a3 b3
    uint32_t r0, r1, r2, r3;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r0) : "r"(w), "r"(p01), "r"(32768u));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r1) : "r"(w), "r"(p01), "r"(32768u));
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r2) : "r"(w), "r"(p23), "r"(32768u));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r3) : "r"(w), "r"(p23), "r"(32768u));
    return __byte_perm(__byte_perm(r0, r1, 0x0062), __byte_perm(r2, r3, 0x0062), 0x5410);
}

__device__ __forceinline__ uint4 ld16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st16(uint8_t *p, uint4 v) { __stcs(reinterpret_cast<uint4 *>(p), v); }

__device__ __forceinline__ uint4 blend16(const acgpu_rowop &op, const uint8_t *s, size_t c)
{
    // c = byte offset of the chunk inside the row
    if (op.op == ACGPU_ROW_COPY) return ld16(s + op.src1_off + c);
    if (op.op == ACGPU_ROW_AVERAGE || op.op == ACGPU_ROW_AVERAGE3) {
        const uint4 a = ld16(s + op.src1_off + c), b = ld16(s + op.src2_off + c);
        uint4 r = make_uint4(avg4x8(a.x, b.x), avg4x8(a.y, b.y), avg4x8(a.z, b.z), avg4x8(a.w, b.w));
        if (op.op == ACGPU_ROW_AVERAGE3) {
            const uint4 e = ld16(s + op.src3_off + c);
            r = make_uint4(avg4x8(e.x, r.x), avg4x8(e.y, r.y), avg4x8(e.z, r.z), avg4x8(e.w, r.w));
        }
        return r;
    }
    if (op.weight1 >= 0x10000u) return ld16(s + op.src1_off + c);
    if (op.weight2 >= 0x10000u) return ld16(s + op.src2_off + c);
    const uint4 a = ld16(s + op.src1_off + c), b = ld16(s + op.src2_off + c);
    return make_uint4(rescale4x8(a.x, b.x, op.weight1, op.weight2), rescale4x8(a.y, b.y, op.weight1, op.weight2),
                      rescale4x8(a.z, b.z, op.weight1, op.weight2), rescale4x8(a.w, b.w, op.weight1, op.weight2));
}

__device__ __forceinline__ uint8_t blend1(const acgpu_rowop &op, const uint8_t *s, size_t c)
{
    using namespace pixmath;
    if (op.op == ACGPU_ROW_COPY) return s[op.src1_off + c];
    if (op.op == ACGPU_ROW_AVERAGE) return (uint8_t)avg2(s[op.src1_off + c], s[op.src2_off + c]);
    if (op.op == ACGPU_ROW_AVERAGE3)
        return (uint8_t)avg2(s[op.src3_off + c], avg2(s[op.src1_off + c], s[op.src2_off + c]));
    if (op.weight1 >= 0x10000u) return s[op.src1_off + c];
    if (op.weight2 >= 0x10000u) return s[op.src2_off + c];
    return (uint8_t)rescale1(s[op.src1_off + c], s[op.src2_off + c], op.weight1, op.weight2);
}

// Tiled row blends.  The host cuts the operation list into blocks of kRows consecutive operations and lists the
// DISTINCT source rows each block reads (libtcvideo's shapes reuse rows between neighbouring outputs: deinterlace
// reads y-1, y, y+1; resize reads overlapping two-tap windows).  A thread owns one 16-byte column of the tile:
// it fires one 16-byte cp.async (LDGSTS, L2-only) per distinct source row -- all of them in flight at once, no
// registers held -- waits, then computes the block's output rows from shared memory and streams them out.
// Every source byte crosses L2->SM once per block instead of once per use, and per-thread memory parallelism is
// the number of distinct rows (10 for deinterlace), which is what the one-chunk-per-thread version lacked
// (ncu, profiles/r1c: 28 % issue, 24 cycles of long-scoreboard stall per instruction).
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ uint4 blend_words(int op, uint4 a, uint4 b, uint4 e, uint32_t w1, uint32_t w2)
{
    if (op == ACGPU_ROW_COPY) return a;
    if (op == ACGPU_ROW_RESCALE)
        return make_uint4(rescale4x8(a.x, b.x, w1, w2), rescale4x8(a.y, b.y, w1, w2), rescale4x8(a.z, b.z, w1, w2),
                          rescale4x8(a.w, b.w, w1, w2));
    uint4 r = make_uint4(avg4x8(a.x, b.x), avg4x8(a.y, b.y), avg4x8(a.z, b.z), avg4x8(a.w, b.w));
    if (op == ACGPU_ROW_AVERAGE3) r = make_uint4(avg4x8(e.x, r.x), avg4x8(e.y, r.y), avg4x8(e.z, r.z), avg4x8(e.w, r.w));
    return r;
}

__global__ void __launch_bounds__(128) k_rowops_tiled(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                                                     const RowBlk *__restrict__ blks, int row_bytes, int tile_bytes, int ntiles)
{
    extern __shared__ uint4 s_tile[];          // [nsrc][tile_chunks]
    __shared__ RowBlk s_blk;
    const int bi = blockIdx.x / ntiles, tile = blockIdx.x - bi * ntiles;
    {
        const uint32_t *g = reinterpret_cast<const uint32_t *>(blks + bi);
        uint32_t *l = reinterpret_cast<uint32_t *>(&s_blk);
        for (int i = threadIdx.x; i < (int)(sizeof(RowBlk) / 4); i += blockDim.x) l[i] = g[i];
    }
    __syncthreads();
    const int tchunks = tile_bytes >> 4, c = threadIdx.x;
    const size_t col = (size_t)tile * tile_bytes + (size_t)c * 16;
    if (c >= tchunks || col >= (size_t)row_bytes) return;
    const uint8_t *s = src + (size_t)blockIdx.y * spitch + col;
    uint8_t *d = dst + (size_t)blockIdx.y * dpitch + col;
    const int nsrc = s_blk.nsrc;
    for (int r = 0; r < nsrc; r++) cp_async16(&s_tile[r * tchunks + c], s + s_blk.src_off[r]);
    cp_async_wait_all();
    const int nops = s_blk.nops;
    for (int k = 0; k < nops; k++) {
        const RowTask t = s_blk.t[k];
        const uint4 a = s_tile[t.s1 * tchunks + c];
        const uint4 b = s_tile[t.s2 * tchunks + c];
        const uint4 e = s_tile[t.s3 * tchunks + c];
        st16(d + t.dest_off, blend_words(t.op, a, b, e, t.w1, t.w2));
    }
}

__global__ void k_rowops_bytes(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                               const acgpu_rowop *__restrict__ ops, int nops, int row_bytes)
{
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= (size_t)nops * row_bytes) return;
    const int row = (int)(idx / row_bytes);
    const size_t c = idx - (size_t)row * row_bytes;
    const acgpu_rowop op = ops[row];
    dst[(size_t)blockIdx.y * dpitch + op.dest_off + c] = blend1(op, src + (size_t)blockIdx.y * spitch, c);
}

// Flat two-source blend (the kernel behind the legacy ac_average / ac_rescale calls).
__global__ void k_blend_flat(const uint8_t *s1, const uint8_t *s2, uint8_t *d, size_t bytes,
                             uint32_t w1, uint32_t w2, int op, int vec)
{
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (vec) {
        const size_t c = idx * 16;
        if (c >= bytes) return;
        const uint4 a = ld16(s1 + c), b = ld16(s2 + c);
        uint4 r;
        if (op == ACGPU_ROW_AVERAGE)
            r = make_uint4(avg4x8(a.x, b.x), avg4x8(a.y, b.y), avg4x8(a.z, b.z), avg4x8(a.w, b.w));
        else
            r = make_uint4(rescale4x8(a.x, b.x, w1, w2), rescale4x8(a.y, b.y, w1, w2),
                           rescale4x8(a.z, b.z, w1, w2), rescale4x8(a.w, b.w, w1, w2));
        st16(d + c, r);
    } else {
        if (idx >= bytes) return;
        d[idx] = op == ACGPU_ROW_AVERAGE ? (uint8_t)pixmath::avg2(s1[idx], s2[idx])
                                         : (uint8_t)pixmath::rescale1(s1[idx], s2[idx], w1, w2);
    }
}

// Horizontal half of tcv_resize (libtcvideo/tcvideo.c:481-531): the image is new_h*scale_w blocks of
// width/scale_w source pixels -> new_w/scale_w destination pixels; per destination pixel a table entry
// (source, w1, w2).  One thread per destination byte (the destination is contiguous in block order).
__global__ void k_resize_h(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                           const int32_t *__restrict__ tsrc, const uint32_t *__restrict__ tw1,
                           const uint32_t *__restrict__ tw2, int src_block, int dst_block, int Bpp, size_t nbytes)
{
    const size_t o = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (o >= nbytes) return;
    const size_t pix = o / Bpp;
    const int k = (int)(o - pix * Bpp);
    const size_t blk = pix / dst_block;
    const int x = (int)(pix - blk * dst_block);
    const uint8_t *a = src + (size_t)blockIdx.y * spitch + (blk * src_block + tsrc[x]) * Bpp + k;
    const uint32_t w1 = tw1[x], w2 = tw2[x];
    dst[(size_t)blockIdx.y * dpitch + o] =
        w1 < 0x10000u ? (uint8_t)pixmath::rescale1(a[0], a[Bpp], w1, w2) : a[0];
}

// Vectorised horizontal resize: one CUDA block per image row.  The source row is staged in shared memory with
// cp.async, then each thread produces 4-byte output words: per output byte a 16-bit source offset and a packed
// (w1, w2) weight pair come from host-built per-row tables (identical for every row and frame, so they live in L1),
// the two taps are gathered from shared memory and blended with one dp2a.  A copy tap (weight1 >= 65536, the
// source pixel must be returned untouched and the second tap must not be read: tcvideo.c:517-531) is encoded as
// weights (65535, 0) with both taps on the same byte: (a*65535 + 32768) >> 16 == a for every byte a.
constexpr int kResizeRows = 4;      // image rows per block: amortises the stage/sync and keeps 4 rows of loads in flight
__global__ void __launch_bounds__(256) k_resize_h_row(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                                                     const uint16_t *__restrict__ toff, const uint32_t *__restrict__ twgt,
                                                     int src_row_bytes, int dst_row_bytes, int Bpp, int rows)
{
    extern __shared__ uint4 s_row[];
    const int row0 = blockIdx.x * kResizeRows;
    const int nrows = min(kResizeRows, rows - row0);
    const uint8_t *s = src + (size_t)blockIdx.y * spitch + (size_t)row0 * src_row_bytes;
    uint8_t *d = dst + (size_t)blockIdx.y * dpitch + (size_t)row0 * dst_row_bytes;
    const int schunks = (src_row_bytes >> 4) * nrows;          // the rows are contiguous in the frame
    for (int c = threadIdx.x; c < schunks; c += blockDim.x) cp_async16(&s_row[c], s + (size_t)c * 16);
    cp_async_wait_all();
    __syncthreads();
    const uint8_t *sb = reinterpret_cast<const uint8_t *>(s_row);
    const int wpr = dst_row_bytes >> 2;
    for (int ow = threadIdx.x; ow < wpr; ow += blockDim.x) {
        const uint2 o2 = __ldg(reinterpret_cast<const uint2 *>(toff) + ow);      // four 16-bit offsets
        const uint4 w4 = __ldg(reinterpret_cast<const uint4 *>(twgt) + ow);      // four weight pairs
        const uint32_t off[4] = {o2.x & 0xFFFFu, o2.x >> 16, o2.y & 0xFFFFu, o2.y >> 16};
        const uint32_t wt[4] = {w4.x, w4.y, w4.z, w4.w};
        for (int r = 0; r < nrows; r++) {
            const uint8_t *rb = sb + (size_t)r * src_row_bytes;
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t a = rb[off[j]];
                const uint32_t b = rb[off[j] + ((wt[j] >> 16) ? Bpp : 0)];       // copy taps: never look at the neighbour
                asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(v[j]) : "r"(wt[j]), "r"(a | (b << 8)), "r"(32768u));
            }
            __stcs(reinterpret_cast<uint32_t *>(d + (size_t)r * dst_row_bytes) + ow,
                   __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410));
        }
    }
}

// Window form of the same kernel (the default whenever the table allows it, i.e. for every ratio up to ~2:1): the four
// first taps of an output word lie within 8 consecutive source bytes, and the second taps are those bytes + Bpp.  So
// instead of eight byte gathers the thread reads three or four aligned shared-memory words at the window base, funnel-shifts
// them to the base's byte alignment, picks the taps with one PRMT each using host-built selectors, interleaves them with two
// more PRMTs and blends two outputs per word with dp2a.lo / dp2a.hi.
// meta[ow] = { window base (source byte offset in the row, 16 bits) | selector of the first taps (16 bits),
//              selector of the second taps (NARROW only) }; weights as above.
// NARROW: every second tap that carries weight also lies inside the window's 8 bytes (always when enlarging a Bpp 1 plane,
// and up to ~2:1 when shrinking one), so both tap sets come out of the same two shifted words: three shared loads, two
// funnel shifts and two PRMTs per row instead of four, five and two.
// The kernel is bound by instruction issue (ncu: issue 78 %, ALU pipe 71 % enlarging 1920 -> 2560), so the shape of the row
// loop matters: rows of a full block are unrolled without per-row branches, the shared window address is formed once per
// output word (shared-space loads through a 32-bit address; the generic form re-derived the shared base in every row), and
// the store pointer advances by the row pitch.
__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
template <int BPP, bool NARROW>
__device__ __forceinline__ uint32_t resize_word(uint32_t saddr, uint32_t sh, uint32_t selA, uint32_t selB, const uint4 &w4)
{
    const uint32_t w0 = lds32(saddr), w1 = lds32(saddr + 4), w2 = lds32(saddr + 8);
    const uint32_t W0 = __funnelshift_r(w0, w1, sh), W1 = __funnelshift_r(w1, w2, sh);
    uint32_t A, B;
    if (NARROW) {
        A = __byte_perm(W0, W1, selA);
        B = __byte_perm(W0, W1, selB);
    } else {
        const uint32_t w3 = BPP == 1 ? 0u : lds32(saddr + 12);
        const uint32_t W2 = __funnelshift_r(w2, w3, sh);
        A = __byte_perm(W0, W1, selA);
        B = __byte_perm(__funnelshift_r(W0, W1, 8 * BPP), __funnelshift_r(W1, W2, 8 * BPP), selA);
    }
    const uint32_t X0 = __byte_perm(A, B, 0x5140), X1 = __byte_perm(A, B, 0x7362);
    uint32_t v0, v1, v2, v3;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(v0) : "r"(w4.x), "r"(X0), "r"(32768u));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(v1) : "r"(w4.y), "r"(X0), "r"(32768u));
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(v2) : "r"(w4.z), "r"(X1), "r"(32768u));
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(v3) : "r"(w4.w), "r"(X1), "r"(32768u));
    return __byte_perm(__byte_perm(v0, v1, 0x0062), __byte_perm(v2, v3, 0x0062), 0x5410);
}
template <int BPP, bool NARROW>
__global__ void __launch_bounds__(256) k_resize_h_win(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                                                     const uint2 *__restrict__ meta, const uint32_t *__restrict__ twgt,
                                                     int src_row_bytes, int dst_row_bytes, int rows)
{
    extern __shared__ uint4 s_row[];
    const int row0 = blockIdx.x * kResizeRows;
    const int nrows = min(kResizeRows, rows - row0);
    const uint8_t *s = src + (size_t)blockIdx.y * spitch + (size_t)row0 * src_row_bytes;
    uint8_t *d = dst + (size_t)blockIdx.y * dpitch + (size_t)row0 * dst_row_bytes;
    const int schunks = (src_row_bytes >> 4) * nrows;          // the rows are contiguous in the frame
    for (int c = threadIdx.x; c < schunks; c += blockDim.x) cp_async16(&s_row[c], s + (size_t)c * 16);
    cp_async_wait_all();
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_row);
    const int wpr = dst_row_bytes >> 2;
    for (int ow = threadIdx.x; ow < wpr; ow += blockDim.x) {
        const uint2 m = __ldg(meta + ow);
        const uint4 w4 = __ldg(reinterpret_cast<const uint4 *>(twgt) + ow);      // four weight pairs
        const uint32_t base = m.x & 0xFFFFu, selA = m.x >> 16, selB = m.y, sh = (base & 3u) * 8;
        uint32_t saddr = sbase + (base & ~3u);
        uint32_t *out = reinterpret_cast<uint32_t *>(d) + ow;
        if (nrows == kResizeRows) {
#pragma unroll
            for (int r = 0; r < kResizeRows; r++) {
                __stcs(out, resize_word<BPP, NARROW>(saddr, sh, selA, selB, w4));
                saddr += (uint32_t)src_row_bytes;
                out += wpr;
            }
        } else {
            for (int r = 0; r < nrows; r++) {
                __stcs(out, resize_word<BPP, NARROW>(saddr, sh, selA, selB, w4));
                saddr += (uint32_t)src_row_bytes;
                out += wpr;
            }
        }
    }
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// Host side of the tiled kernel: blocks of kRows operations with their distinct source rows.
int build_row_blocks(const acgpu_rowop *ops, int nops, std::vector<RowBlk> &out)
{
    int max_nsrc = 1;
    out.clear();
    for (int i0 = 0; i0 < nops; i0 += RowBlk::kRows) {
        RowBlk b;
        memset(&b, 0, sizeof(b));
        b.nops = nops - i0 < RowBlk::kRows ? nops - i0 : RowBlk::kRows;
        auto slot = [&](int64_t off) -> uint8_t {
            for (int r = 0; r < b.nsrc; r++)
                if (b.src_off[r] == off) return (uint8_t)r;
            b.src_off[b.nsrc] = off;
            return (uint8_t)b.nsrc++;
        };
        for (int k = 0; k < b.nops; k++) {
            const acgpu_rowop &o = ops[i0 + k];
            RowTask &t = b.t[k];
            t.dest_off = o.dest_off;
            t.w1 = o.weight1; t.w2 = o.weight2;
            // aclib/rescale.c:26-29: a weight >= 65536 turns the blend into a copy and the other row is NEVER read
            if (o.op == ACGPU_ROW_COPY || (o.op == ACGPU_ROW_RESCALE && o.weight1 >= 0x10000u)) {
                t.op = ACGPU_ROW_COPY; t.s1 = t.s2 = t.s3 = slot(o.src1_off);
            } else if (o.op == ACGPU_ROW_RESCALE && o.weight2 >= 0x10000u) {
                t.op = ACGPU_ROW_COPY; t.s1 = t.s2 = t.s3 = slot(o.src2_off);
            } else {
                t.op = (uint8_t)o.op;
                t.s1 = slot(o.src1_off);
                t.s2 = slot(o.src2_off);
                t.s3 = o.op == ACGPU_ROW_AVERAGE3 ? slot(o.src3_off) : t.s1;
            }
        }
        if (b.nsrc > max_nsrc) max_nsrc = b.nsrc;
        out.push_back(b);
    }
    return max_nsrc;
}

bool rowops_tiled_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const RowBlk *d_blks, int nblks,
                         int max_nsrc, int row_bytes, int nframes, cudaStream_t st)
{
    if (nblks <= 0 || row_bytes <= 0 || nframes <= 0) return true;
    const int ntiles = (row_bytes + 2047) / 2048;
    const int tile_bytes = (((row_bytes + ntiles - 1) / ntiles) + 15) / 16 * 16;
    const int threads = ((tile_bytes / 16) + 31) / 32 * 32;
    const size_t smem = (size_t)max_nsrc * tile_bytes;
    if (smem > 48 * 1024 && !check(cudaFuncSetAttribute(k_rowops_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr"))
        return false;
    dim3 g((unsigned)(nblks * ntiles), (unsigned)nframes);
    k_rowops_tiled<<<g, threads, smem, st>>>(src, spitch, dst, dpitch, d_blks, row_bytes, tile_bytes, ntiles);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_rowops_tiled");
    return true;
}

bool rowops_bytes_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                         const acgpu_rowop *d_ops, int nops, int row_bytes, int nframes, cudaStream_t st)
{
    if (nops <= 0 || row_bytes <= 0 || nframes <= 0) return true;
    const size_t n = (size_t)nops * row_bytes;
    dim3 g((unsigned)((n + TPB - 1) / TPB), (unsigned)nframes);
    k_rowops_bytes<<<g, TPB, 0, st>>>(src, spitch, dst, dpitch, d_ops, nops, row_bytes);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_rowops_bytes");
    return true;
}

bool rowops_vectorisable(const uint8_t *src, size_t spitch, const uint8_t *dst, size_t dpitch, int row_bytes)
{
    return aligned16(src) && aligned16(dst) && spitch % 16 == 0 && dpitch % 16 == 0 && row_bytes % 16 == 0;
}

bool blend_launch(const uint8_t *s1, const uint8_t *s2, uint8_t *d, size_t bytes,
                  uint32_t w1, uint32_t w2, int op, cudaStream_t st)
{
    if (bytes == 0) return true;
    const int vec = aligned16(s1) && aligned16(s2) && aligned16(d) && bytes % 16 == 0;
    const size_t n = vec ? bytes / 16 : bytes;
    k_blend_flat<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, st>>>(s1, s2, d, bytes, w1, w2, op, vec);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_blend_flat");
    return true;
}

bool resize_h_row_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint16_t *d_off,
                         const uint32_t *d_wgt, int width, int new_w, int rows, int Bpp, int nframes, cudaStream_t st)
{
    const int srb = width * Bpp, drb = new_w * Bpp;
    if (rows <= 0 || nframes <= 0) return true;
    const size_t smem = (size_t)srb * kResizeRows + 16;
    if (smem > 48 * 1024 && !check(cudaFuncSetAttribute(k_resize_h_row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr"))
        return false;
    dim3 g((unsigned)((rows + kResizeRows - 1) / kResizeRows), (unsigned)nframes);
    k_resize_h_row<<<g, 256, smem, st>>>(src, spitch, dst, dpitch, d_off, d_wgt, srb, drb, Bpp, rows);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_resize_h_row");
    return true;
}

bool resize_h_win_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint2 *d_meta,
                         const uint32_t *d_wgt, int width, int new_w, int rows, int Bpp, bool narrow, int nframes, cudaStream_t st)
{
    const int srb = width * Bpp, drb = new_w * Bpp;
    if (rows <= 0 || nframes <= 0) return true;
    const size_t smem = (size_t)srb * kResizeRows + 32;        // the last window may look up to 16 bytes past the rows
    auto kern = Bpp == 1 ? (narrow ? k_resize_h_win<1, true> : k_resize_h_win<1, false>)
                         : (narrow ? k_resize_h_win<3, true> : k_resize_h_win<3, false>);
    if (smem > 48 * 1024 && !check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr"))
        return false;
    dim3 g((unsigned)((rows + kResizeRows - 1) / kResizeRows), (unsigned)nframes);
    // block size: the output words of a row split evenly over the passes (320 words: 2 x 160 threads, not 256 + 64)
    const int wpr = drb / 4, passes = (wpr + 255) / 256;
    const int threads = std::min(256, (((wpr + passes - 1) / passes) + 31) / 32 * 32);
    kern<<<g, threads, smem, st>>>(src, spitch, dst, dpitch, d_meta, d_wgt, srb, drb, rows);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_resize_h_win");
    return true;
}

bool resize_h_vectorisable(const uint8_t *src, size_t spitch, const uint8_t *dst, size_t dpitch, int width, int new_w, int Bpp)
{
    return aligned16(src) && aligned16(dst) && spitch % 16 == 0 && dpitch % 16 == 0 && (width * Bpp) % 16 == 0
        && (new_w * Bpp) % 16 == 0 && width * Bpp <= 48000 && new_w * Bpp <= 65000;
}

bool resize_h_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                     const int32_t *d_source, const uint32_t *d_w1, const uint32_t *d_w2,
                     int width, int new_w, int new_h, int Bpp, int scale_w, int nframes, cudaStream_t st)
{
    const size_t nbytes = (size_t)new_w * new_h * Bpp;
    if (nbytes == 0 || nframes <= 0) return true;
    dim3 g((unsigned)((nbytes + TPB - 1) / TPB), (unsigned)nframes);
    k_resize_h<<<g, TPB, 0, st>>>(src, spitch, dst, dpitch, d_source, d_w1, d_w2,
                                  width / scale_w, new_w / scale_w, Bpp, nbytes);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_resize_h");
    return true;
}

}  // namespace acgpu
