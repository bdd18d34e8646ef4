// tcvops.cu -- the element-wise libtcvideo plane operations that surround the aclib path:
// tcv_clip, tcv_reduce, tcv_flip_v, tcv_flip_h, tcv_gamma_correct, tcv_antialias (libtcvideo/tcvideo.c:184-250,
// 681-980).  One plane of width x height pixels, Bpp 1 or 3, tightly packed; a batch is `nframes` such planes a fixed
// pitch apart.  Everything here is data movement or a per-byte table map, so the bound is HBM: each kernel reads and
// writes every byte once with 16-byte accesses wherever the geometry keeps chunks aligned, and falls back to a
// byte-granular path (same kernel, warp-uniform branch) where it does not.
#include "fast_common.cuh"

namespace acgpu {
namespace {

using namespace fast;

constexpr int kThreads = 256;

inline dim3 grid_for(uint64_t work_items, int nframes, int nwaves = 32)
{
    long want = ((long)sm_count() * 8 * waves(nwaves) + nframes - 1) / nframes;
    const long maxgx = (long)((work_items + kThreads - 1) / kThreads);
    if (want > maxgx) want = maxgx;
    if (want < 1) want = 1;
    return dim3((unsigned)want, (unsigned)nframes);
}

// 16 bytes from any address: one LDG.128 when aligned, else the aligned 32-bit words that hold them, funnel-shifted.
// COHERENT: plain loads (the caller's source may alias its destination); otherwise the read-only path.
template <bool COHERENT>
__device__ __forceinline__ uint32_t ld32(const uint32_t *q) { return COHERENT ? *q : __ldg(q); }
template <bool COHERENT>
__device__ __forceinline__ uint4 ld16_any(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 15) == 0) return COHERENT ? *reinterpret_cast<const uint4 *>(p) : ldg128(p);
    // the two aligned 16-byte blocks that hold the bytes (one L1 wavefront each for a warp of consecutive chunks; five
    // 32-bit loads per chunk cost five), then a shift by whole words and one by bytes
    const uint4 *q = reinterpret_cast<const uint4 *>(a & ~(uintptr_t)15);
    const uint4 lo = COHERENT ? q[0] : __ldg(q), hi = COHERENT ? q[1] : __ldg(q + 1);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t w0, w1, w2, w3, w4;
    switch ((a >> 2) & 3) {
    case 0:  w0 = lo.x; w1 = lo.y; w2 = lo.z; w3 = lo.w; w4 = hi.x; break;
    case 1:  w0 = lo.y; w1 = lo.z; w2 = lo.w; w3 = hi.x; w4 = hi.y; break;
    case 2:  w0 = lo.z; w1 = lo.w; w2 = hi.x; w3 = hi.y; w4 = hi.z; break;
    default: w0 = lo.w; w1 = hi.x; w2 = hi.y; w3 = hi.z; w4 = hi.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// ---------------------------------------------------------------------------------------------------
// Window copy: destination row y shows bytes [sxb, sxb + cn) of source row y*row_mul + row_add at its bytes
// [cl, cl + cn) and `fill` everywhere else (and in every row whose source row falls outside the source).
// Serves tcv_clip (tcvideo.c:184-250: crop and/or grow with black borders) and the rows-only case of tcv_reduce
// (:706-711).  The destination frame is walked as flat 16-byte chunks; a chunk that lies inside one row's copy span
// is one (possibly unaligned) 16-byte load and one aligned store.

__device__ __forceinline__ uint32_t window_byte(const TcvWindow &p, const uint8_t *src, uint32_t off)
{
    const uint32_t y = off / p.dBpl, xb = off - y * p.dBpl;
    const int sr = (int)y * p.row_mul + p.row_add;
    if (sr < 0 || sr >= p.srows || xb < p.cl || xb >= p.cl + p.cn) return p.fill & 0xFFu;
    return src[(size_t)sr * p.sBpl + p.sxb + (xb - p.cl)];
}

__device__ __forceinline__ void window_row_of(const TcvWindow &p, uint32_t off, uint32_t &y, uint32_t &xb) { p.row_div.divmod(off, y, xb); }

// Bytes [sh, sh + 16) of the 32 bytes lo | hi, sh = 0..15, without a branch: drop two words, drop one word, shift by bytes.
__device__ __forceinline__ uint4 shift16(const uint4 &lo, const uint4 &hi, uint32_t sh)
{
    const bool two = (sh & 8u) != 0, one = (sh & 4u) != 0;
    const uint32_t x0 = two ? lo.z : lo.x, x1 = two ? lo.w : lo.y, x2 = two ? hi.x : lo.z, x3 = two ? hi.y : lo.w,
                   x4 = two ? hi.z : hi.x, x5 = two ? hi.w : hi.y;
    const uint32_t y0 = one ? x1 : x0, y1 = one ? x2 : x1, y2 = one ? x3 : x2, y3 = one ? x4 : x3, y4 = one ? x5 : x4;
    const uint32_t bs = (sh & 3u) * 8u;
    return make_uint4(__funnelshift_r(y0, y1, bs), __funnelshift_r(y1, y2, bs), __funnelshift_r(y2, y3, bs), __funnelshift_r(y3, y4, bs));
}

// One destination chunk: classified and its loads ISSUED; nothing here waits for them, so the chunks of a trip are all in
// flight together (shifting inside this step made every unaligned chunk wait for its own data before the next one was
// even requested).  kind: 0 = nothing (past the end, or a mixed chunk: those belong to the second phase), 1 = store lo
// (an aligned source), 2 = byte-granular tail / unaligned destination, 3 = store shift16(lo, hi, sh), 4 = store the fill.
// SHIFTED = false: the launcher has checked that every chunk's source is 16-byte aligned (whole-unit crops, borders,
// rows-only reduce): one load per chunk and no shift state, which keeps that kernel at the copy rate.
struct WinChunk { uint4 lo, hi; uint32_t sh; int kind; };
template <bool SHIFTED>
__device__ __forceinline__ void window_chunk(const TcvWindow &p, const uint8_t *src, uint32_t c, uint32_t N, WinChunk &o)
{
    const uint32_t off = c * 16;
    o.kind = 0;
    if (off >= N) return;
    if (!(p.vec && off + 16 <= N)) { o.kind = 2; return; }
    uint32_t y, xb;
    window_row_of(p, off, y, xb);
    if (xb + 16 > p.dBpl) return;
    const int sr = (int)y * p.row_mul + p.row_add;
    if (sr < 0 || sr >= p.srows || xb + 16 <= p.cl || xb >= p.cl + p.cn) { o.kind = 4; return; }
    if (!(xb >= p.cl && xb + 16 <= p.cl + p.cn)) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(src + (size_t)sr * p.sBpl + p.sxb + (xb - p.cl));
    if (!SHIFTED) {
        o.lo = __ldg(reinterpret_cast<const uint4 *>(a));
        o.kind = 1;
        return;
    }
    const uint4 *q = reinterpret_cast<const uint4 *>(a & ~(uintptr_t)15);
    o.sh = (uint32_t)(a & 15);
    o.lo = __ldg(q);
    if (o.sh) o.hi = __ldg(q + 1);        // an unaligned chunk ends inside the next block; an aligned one may end the buffer
    o.kind = o.sh ? 3 : 1;
}

// A chunk with a row end or a border edge inside it: its 16 bytes one by one with a running (row, column).
__device__ __forceinline__ uint4 window_mixed_chunk(const TcvWindow &p, const uint8_t *src, uint32_t off)
{
    uint32_t w[4] = {0, 0, 0, 0}, yy, xx;
    window_row_of(p, off, yy, xx);
    int sr = (int)yy * p.row_mul + p.row_add;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t b = p.fill & 0xFFu;
        if (sr >= 0 && sr < p.srows && xx >= p.cl && xx < p.cl + p.cn) b = __ldg(src + (size_t)sr * p.sBpl + p.sxb + (xx - p.cl));
        w[i >> 2] |= b << (8 * (i & 3));
        if (++xx == p.dBpl) { xx = 0; yy++; sr += p.row_mul; }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kWinUnroll = 4;     // independent 16-byte loads in flight per thread (a lone load per thread is latency-bound)

// Two phases.  (1) every whole chunk that lies in ONE region (copy span, border, filled row): one load, one store, four
// chunks in flight per thread.  (2) the mixed chunks, found from the other side: a chunk is mixed exactly when a row
// start, the left edge of the copy span or its right edge falls strictly inside it, so the (at most three per row)
// edges are enumerated and each rebuilds the chunk it cuts byte by byte.  Leaving them in phase 1 made every warp run
// the byte path (a 1912-byte row puts one mixed chunk into every 120: 0.55 of the copy rate); two edges in one chunk
// write the same sixteen bytes twice.
template <bool SHIFTED>
__global__ void __launch_bounds__(kThreads) k_window(TcvWindow p)
{
    const uint8_t *src = p.src + (size_t)blockIdx.y * p.spitch;
    uint8_t *dst = p.dst + (size_t)blockIdx.y * p.dpitch;
    const uint32_t N = p.dBpl * (uint32_t)p.drows, nchunks = (N + 15) / 16;
    const uint32_t per_block = kThreads * kWinUnroll;
    for (uint32_t base = blockIdx.x * per_block; base < nchunks; base += gridDim.x * per_block) {
        WinChunk ch[kWinUnroll];
#pragma unroll
        for (int k = 0; k < kWinUnroll; k++) window_chunk<SHIFTED>(p, src, base + k * kThreads + threadIdx.x, N, ch[k]);
#pragma unroll
        for (int k = 0; k < kWinUnroll; k++) {
            const uint32_t off = (base + k * kThreads + threadIdx.x) * 16;
            if (ch[k].kind == 1) {
                stg128(dst + off, ch[k].lo);
            } else if (SHIFTED && ch[k].kind == 3) {
                stg128(dst + off, shift16(ch[k].lo, ch[k].hi, ch[k].sh));
            } else if (ch[k].kind == 4) {
                stg128(dst + off, make_uint4(p.fill, p.fill, p.fill, p.fill));
            } else if (ch[k].kind == 2) {
                for (uint32_t i = 0; i < 16 && off + i < N; i++) dst[off + i] = (uint8_t)window_byte(p, src, off + i);
            }
        }
    }
    if (!p.vec) return;
    const uint32_t nedges = 3u * (uint32_t)p.drows;
    for (uint32_t e = blockIdx.x * kThreads + threadIdx.x; e < nedges; e += gridDim.x * kThreads) {
        const uint32_t y = e / 3u, k = e - 3u * y;
        if ((k == 1 && p.cl == 0) || (k == 2 && p.cl + p.cn >= p.dBpl)) continue;      // the same place as a row start
        const uint32_t b = y * p.dBpl + (k == 0 ? 0u : k == 1 ? p.cl : p.cl + p.cn);
        const uint32_t off = b & ~15u;
        if (off == b || off + 16 > N) continue;                                         // on a chunk boundary / the byte-granular tail
        stg128(dst + off, window_mixed_chunk(p, src, off));
    }
}

// ---------------------------------------------------------------------------------------------------
// tcv_reduce, general case (tcvideo.c:694-704): destination pixel (x, y) = source pixel (x*rw, y*rh).
// A thread gathers four consecutive destination pixels (they may wrap to the next destination row) and, when the
// destination allows it, writes them as whole 32-bit words.
template <int BPP>
__global__ void __launch_bounds__(kThreads) k_reduce(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                      uint32_t w, uint32_t ow, uint32_t oh, uint32_t rw, uint32_t rh, int vec)
{
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t n = ow * oh, nq = (n + 3) / 4, stride = gridDim.x * blockDim.x;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const uint32_t i0 = q * 4;
        uint32_t y = i0 / ow, x = i0 - y * ow;
        uint8_t b[4 * BPP];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < n) {
                const uint8_t *s = src + ((size_t)y * rh * w + (size_t)x * rw) * BPP;
#pragma unroll
                for (int k = 0; k < BPP; k++) b[j * BPP + k] = __ldg(s + k);
            } else {
#pragma unroll
                for (int k = 0; k < BPP; k++) b[j * BPP + k] = 0;
            }
            if (++x == ow) { x = 0; y++; }
        }
        uint8_t *d = dst + (size_t)i0 * BPP;
        if (vec && i0 + 4 <= n) {
#pragma unroll
            for (int k = 0; k < BPP; k++)
                stg32(d + 4 * k, (uint32_t)b[4 * k] | ((uint32_t)b[4 * k + 1] << 8) | ((uint32_t)b[4 * k + 2] << 16) | ((uint32_t)b[4 * k + 3] << 24));
        } else {
            for (uint32_t k = 0; k < 4 * BPP && i0 * BPP + k < n * BPP; k++) d[k] = b[k];
        }
    }
}

// reduce_w == 2 (the usual "half size" preview), width % 32 == 0, aligned planes: a thread turns 32 source pixels
// (two or six 128-bit loads) into 16 destination pixels (one or three 128-bit stores) by keeping the even ones.
template <int BPP>
__global__ void __launch_bounds__(kThreads) k_reduce_w2(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                         uint32_t w, uint32_t oh, uint32_t rh)
{
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t upr = w / 32, n = upr * oh, stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t y = i / upr, u = i - y * upr;
        const uint8_t *s = src + ((size_t)y * rh * w + (size_t)u * 32) * BPP;
        uint8_t *d = dst + ((size_t)y * (w / 2) + (size_t)u * 16) * BPP;
        if (BPP == 1) {
            const uint4 a = ldg128(s), b = ldg128(s + 16);
            stg128(d, make_uint4(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420),
                                 __byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420)));
        } else {
            uint32_t px[32], o[12];
            load_rgb16<L_RGB24>(s, 0, true, px);
            load_rgb16<L_RGB24>(s + 48, 0, true, px + 16);
#pragma unroll
            for (int g = 0; g < 4; g++) {       // four kept pixels (source 8g, 8g+2, 8g+4, 8g+6) -> three words
                const uint32_t p0 = px[8 * g], p1 = px[8 * g + 2], p2 = px[8 * g + 4], p3 = px[8 * g + 6];
                o[3 * g + 0] = __byte_perm(p0, p1, 0x4210);
                o[3 * g + 1] = __byte_perm(p1, p2, 0x5421);
                o[3 * g + 2] = __byte_perm(p2, p3, 0x6542);
            }
#pragma unroll
            for (int k = 0; k < 3; k++) stg128(d + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
        }
    }
}

// reduce_w == RW (3 .. 6), output rows a multiple of 16 pixels, aligned planes: a thread turns 16*RW source pixels
// (RW*BPP 128-bit loads) into 16 destination pixels (BPP 128-bit stores); which source byte lands where is known at
// compile time, so the selection is a handful of byte permutes instead of one byte load per destination byte.
template <int BPP, int RW>
__global__ void __launch_bounds__(kThreads) k_reduce_wn(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                         uint32_t w, uint32_t ow, uint32_t oh, uint32_t rh)
{
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t upr = ow / 16, n = upr * oh, stride = gridDim.x * blockDim.x;
    constexpr int NS = 4 * RW * BPP, ND = 4 * BPP;          // source / destination words per thread
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t y = i / upr, u = i - y * upr;
        const uint8_t *s = src + ((size_t)y * rh * w + (size_t)u * 16 * RW) * BPP;
        uint8_t *d = dst + ((size_t)y * ow + (size_t)u * 16) * BPP;
        uint32_t S[NS], O[ND];
#pragma unroll
        for (int k = 0; k < NS / 4; k++) {
            const uint4 v = ldg128(s + 16 * k);
            S[4 * k] = v.x; S[4 * k + 1] = v.y; S[4 * k + 2] = v.z; S[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < ND; k++) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int ob = 4 * k + j, px = ob / BPP, ch = ob % BPP, sb = px * RW * BPP + ch;    // destination byte -> source byte
                word |= ((S[sb >> 2] >> (8 * (sb & 3))) & 0xFFu) << (8 * j);
            }
            O[k] = word;
        }
#pragma unroll
        for (int k = 0; k < ND / 4; k++) stg128(d + 16 * k, make_uint4(O[4 * k], O[4 * k + 1], O[4 * k + 2], O[4 * k + 3]));
    }
}

// ---------------------------------------------------------------------------------------------------
// tcv_flip_v (tcvideo.c:739-766): row y <-> row height-1-y.  Every thread owns one 16-byte column slice of a row PAIR:
// it reads both slices, then writes both, so src == dest (which the reference allows, :757-762) needs no temporary.
// UNROLL slices per thread are read before any is written.  The loads are coherent (src may be dest) and streaming
// (ld.global.cs): with plain loads the in-place flip ran at 0.69 of the copy rate, with streaming ones at 1.0 for every
// UNROLL measured (1, 2, 4) -- lines that are about to be overwritten should not stay in L1.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads) k_flip_v(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                      uint32_t Bpl, uint32_t h, int vec)
{
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t ncr = (Bpl + 15) / 16, n = ((h + 1) / 2) * ncr;
    if (vec) {
        const uint32_t per_block = kThreads * UNROLL;
        for (uint32_t base = blockIdx.x * per_block; base < n; base += gridDim.x * per_block) {
            uint4 a[UNROLL], b[UNROLL];
            size_t oa[UNROLL], ob[UNROLL];
#pragma unroll
            for (int k = 0; k < UNROLL; k++) {
                const uint32_t i = base + k * kThreads + threadIdx.x;
                if (i < n) {
                    const uint32_t y = i / ncr, xb = (i - y * ncr) * 16, m = h - 1 - y;
                    oa[k] = (size_t)y * Bpl + xb;
                    ob[k] = (size_t)m * Bpl + xb;
                    a[k] = __ldcs(reinterpret_cast<const uint4 *>(src + oa[k]));      // coherent (src may be dest), streaming
                    b[k] = __ldcs(reinterpret_cast<const uint4 *>(src + ob[k]));
                }
            }
#pragma unroll
            for (int k = 0; k < UNROLL; k++)
                if (base + k * kThreads + threadIdx.x < n) {
                    stg128(dst + oa[k], b[k]);
                    stg128(dst + ob[k], a[k]);
                }
        }
        return;
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t y = i / ncr, xb = (i - y * ncr) * 16, m = h - 1 - y;
        const size_t oa = (size_t)y * Bpl + xb, ob = (size_t)m * Bpl + xb;
        const uint32_t cnt = min(16u, Bpl - xb);
        uint8_t a[16], b[16];
        for (uint32_t k = 0; k < cnt; k++) { a[k] = src[oa + k]; b[k] = src[ob + k]; }
        for (uint32_t k = 0; k < cnt; k++) { dst[oa + k] = b[k]; dst[ob + k] = a[k]; }
    }
}

// ---------------------------------------------------------------------------------------------------
// tcv_flip_h (tcvideo.c:787-818): pixel x <-> pixel width-1-x inside every row; src == dest allowed (:806-814).
// Vector form (width % 16 == 0): a thread owns one 16-pixel group and its mirror group, reverses the pixel order of
// both in registers and writes them to each other's place.
__device__ __forceinline__ uint4 reverse_bytes16(uint4 v)
{
    return make_uint4(__byte_perm(v.w, 0, 0x0123), __byte_perm(v.z, 0, 0x0123), __byte_perm(v.y, 0, 0x0123), __byte_perm(v.x, 0, 0x0123));
}
// 16 RGB pixels in 12 words -> the same pixels in reverse order
__device__ __forceinline__ void reverse_pixels48(const uint32_t *w, uint32_t *o)
{
    uint32_t px[16];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        px[4 * g + 0] = w[3 * g];
        px[4 * g + 1] = __byte_perm(w[3 * g], w[3 * g + 1], 0x0543);
        px[4 * g + 2] = __byte_perm(w[3 * g + 1], w[3 * g + 2], 0x0432);
        px[4 * g + 3] = w[3 * g + 2] >> 8;
    }
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t p0 = px[15 - 4 * g], p1 = px[14 - 4 * g], p2 = px[13 - 4 * g], p3 = px[12 - 4 * g];
        o[3 * g + 0] = __byte_perm(p0, p1, 0x4210);
        o[3 * g + 1] = __byte_perm(p1, p2, 0x5421);
        o[3 * g + 2] = __byte_perm(p2, p3, 0x6542);
    }
}

template <int BPP>
__global__ void __launch_bounds__(kThreads) k_flip_h(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                      uint32_t w, uint32_t h, int vec)
{
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t stride = gridDim.x * blockDim.x;
    if (vec) {
        const uint32_t gpr = w / 16, hg = (gpr + 1) / 2, n = hg * h;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t y = i / hg, g = i - y * hg, m = gpr - 1 - g;
            const size_t row = (size_t)y * w * BPP, oa = row + (size_t)g * 16 * BPP, ob = row + (size_t)m * 16 * BPP;
            if (BPP == 1) {
                const uint4 a = *reinterpret_cast<const uint4 *>(src + oa), b = *reinterpret_cast<const uint4 *>(src + ob);
                stg128(dst + oa, reverse_bytes16(b));
                stg128(dst + ob, reverse_bytes16(a));
            } else {
                uint32_t a[12], b[12], ra[12], rb[12];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const uint4 va = *reinterpret_cast<const uint4 *>(src + oa + 16 * k), vb = *reinterpret_cast<const uint4 *>(src + ob + 16 * k);
                    a[4 * k] = va.x; a[4 * k + 1] = va.y; a[4 * k + 2] = va.z; a[4 * k + 3] = va.w;
                    b[4 * k] = vb.x; b[4 * k + 1] = vb.y; b[4 * k + 2] = vb.z; b[4 * k + 3] = vb.w;
                }
                reverse_pixels48(a, ra);
                reverse_pixels48(b, rb);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    stg128(dst + oa + 16 * k, make_uint4(rb[4 * k], rb[4 * k + 1], rb[4 * k + 2], rb[4 * k + 3]));
                    stg128(dst + ob + 16 * k, make_uint4(ra[4 * k], ra[4 * k + 1], ra[4 * k + 2], ra[4 * k + 3]));
                }
            }
        }
    } else {
        const uint32_t hw = (w + 1) / 2, n = hw * h;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t y = i / hw, x = i - y * hw, m = w - 1 - x;
            const size_t row = (size_t)y * w * BPP, oa = row + (size_t)x * BPP, ob = row + (size_t)m * BPP;
            uint8_t a[BPP], b[BPP];
#pragma unroll
            for (int k = 0; k < BPP; k++) { a[k] = src[oa + k]; b[k] = src[ob + k]; }
#pragma unroll
            for (int k = 0; k < BPP; k++) { dst[oa + k] = b[k]; dst[ob + k] = a[k]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// tcv_gamma_correct (tcvideo.c:840-858): dest[i] = table[src[i]].  The 256-byte table (built on the host with the
// reference's double arithmetic) sits in shared memory as 64 words: byte lookups conflict at most two-way.  (32 private
// copies of 32-bit entries, conflict-free by construction, were measured and lost: 0.91 -> 0.81-0.84 of the copy rate --
// the fill of 32 KB per block and the third instruction per lookup cost more than the replays, profiles/r2_experiments.md.)
#ifndef ACGPU_LUT_UNROLL
#define ACGPU_LUT_UNROLL 2
#endif
__global__ void __launch_bounds__(kThreads) k_lut(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                   const uint8_t *table, uint32_t nbytes, int vec)
{
    __shared__ uint8_t s_t[256];
    if (threadIdx.x < 64) reinterpret_cast<uint32_t *>(s_t)[threadIdx.x] = __ldg(reinterpret_cast<const uint32_t *>(table) + threadIdx.x);
    __syncthreads();
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t nchunks = (nbytes + 15) / 16, stride = gridDim.x * blockDim.x;
    auto map16 = [&](uint4 v) {
        const uint32_t in[4] = {v.x, v.y, v.z, v.w};
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t a = s_t[in[k] & 0xFFu], b = s_t[(in[k] >> 8) & 0xFFu], cc = s_t[(in[k] >> 16) & 0xFFu], d = s_t[in[k] >> 24];
            out[k] = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(cc, d, 0x0040), 0x5410);
        }
        return make_uint4(out[0], out[1], out[2], out[3]);
    };
    // several chunks per trip: all loads are issued before any is used (one 16-byte load per thread in flight does not
    // cover HBM latency)
    constexpr int U = ACGPU_LUT_UNROLL;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += U * stride) {
        uint4 v[U];
        bool whole[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const uint32_t ck = c + k * stride, off = ck * 16;
            whole[k] = ck < nchunks && vec && off + 16 <= nbytes;
            v[k] = make_uint4(0, 0, 0, 0);
            if (whole[k]) v[k] = *reinterpret_cast<const uint4 *>(src + off);
        }
#pragma unroll
        for (int k = 0; k < U; k++) {
            const uint32_t ck = c + k * stride, off = ck * 16;
            if (whole[k]) stg128(dst + off, map16(v[k]));
            else if (ck < nchunks) for (uint32_t i = off; i < off + 16 && i < nbytes; i++) dst[i] = s_t[src[i]];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// tcv_antialias (tcvideo.c:886-980).  Border pixels are copied.  An interior pixel is smoothed when its left (or
// right) neighbour has the colour of exactly one vertical neighbour and differs from the other one and from the
// opposite horizontal neighbour ("same colour": every channel differs by < 25, :37,917-927); the smoothed channel is
// (d[UL] + y[U] + d[UR] + x[L] + c[C] + x[R] + d[DL] + y[D] + d[DR] + 32768) >> 16 in uint32 with the four 16.16 weight
// tables of :1209-1224 (built on the host in double, like the reference).  tables = c | x | y | d, 256 entries each.
template <int BPP>
__device__ __forceinline__ bool aa_same(const uint8_t *p, const uint8_t *q)
{
    int worst = 0;
#pragma unroll
    for (int k = 0; k < BPP; k++) worst = max(worst, abs((int)__ldg(q + k) - (int)__ldg(p + k)));
    return worst < 25;
}

template <int BPP>
__global__ void __launch_bounds__(kThreads) k_antialias(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                         const uint32_t *tables, uint32_t w, uint32_t h)
{
    __shared__ uint32_t s_t[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_t[i] = __ldg(tables + i);
    __syncthreads();
    const uint8_t *src = src0 + (size_t)blockIdx.y * spitch;
    uint8_t *dst = dst0 + (size_t)blockIdx.y * dpitch;
    const uint32_t n = w * h, stride = gridDim.x * blockDim.x;
    const ptrdiff_t Bpl = (ptrdiff_t)w * BPP;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t y = i / w, x = i - y * w;
        const uint8_t *c = src + (size_t)i * BPP;
        uint8_t *o = dst + (size_t)i * BPP;
        bool smooth = false;
        if (y > 0 && y < h - 1 && x > 0 && x < w - 1) {
            const uint8_t *l = c - BPP, *r = c + BPP, *u = c - Bpl, *d = c + Bpl;
            const bool lu = aa_same<BPP>(l, u), ld = aa_same<BPP>(l, d), ru = aa_same<BPP>(r, u), rd = aa_same<BPP>(r, d);
            if ((lu != ld) || (ru != rd)) smooth = !aa_same<BPP>(l, r);
        }
        if (smooth) {
#pragma unroll
            for (int k = 0; k < BPP; k++) {
                const uint8_t *q = c + k;
                const uint32_t sum = s_t[768 + __ldg(q - Bpl - BPP)] + s_t[512 + __ldg(q - Bpl)] + s_t[768 + __ldg(q - Bpl + BPP)]
                                   + s_t[256 + __ldg(q - BPP)] + s_t[__ldg(q)] + s_t[256 + __ldg(q + BPP)]
                                   + s_t[768 + __ldg(q + Bpl - BPP)] + s_t[512 + __ldg(q + Bpl)] + s_t[768 + __ldg(q + Bpl + BPP)]
                                   + 32768u;
                o[k] = (uint8_t)(sum >> 16);
            }
        } else {
#pragma unroll
            for (int k = 0; k < BPP; k++) o[k] = __ldg(c + k);
        }
    }
}

// Word-parallel form (width % 4 == 0, 4-byte aligned planes): a thread owns PX = 4*NG horizontally adjacent pixels =
// W = BPP*NG 32-bit words of an image row.  "Channel differs by >= 25" is evaluated on four bytes at once (|a-b| with
// VABSDIFF4, then the carry trick ((d & 0x7f) + 103 | d) & 0x80 per byte); the left / right neighbours are the same
// words funnel-shifted by BPP bytes; for Bpp 3 a pixel's three channel bits are OR-ed onto its first byte with two more
// funnel shifts; and the smoothing rule is two LOP3s per word on those masks.  No branch depends on the picture until
// the (usually sparse) pixels that really are smoothed are gathered into a per-warp queue and take the 9-tap table path
// on full warps.
//   NG = 4: 16 pixels per thread, 128-bit accesses (width % 16 == 0, 16-byte aligned planes); NG = 1: 32-bit accesses.
__device__ __forceinline__ uint32_t diff_bits(uint32_t a, uint32_t b)      // bit 7 of each byte: |a - b| >= 25 (other bits: don't care)
{
    const uint32_t d = __vabsdiffu4(a, b);
    return ((d & 0x7F7F7F7Fu) + 0x67676767u) | d;
}

// Per-byte "differs" bits of the thread's W words -> per-PIXEL bits, left on the first byte of every pixel.
template <int BPP, int W>
__device__ __forceinline__ void pixel_diff(const uint32_t *a, const uint32_t *b, uint32_t *out)
{
    uint32_t m[W];
#pragma unroll
    for (int k = 0; k < W; k++) m[k] = diff_bits(a[k], b[k]);
#pragma unroll
    for (int k = 0; k < W; k++) {
        if (BPP == 1) {
            out[k] = m[k];
        } else {
            const uint32_t nx = k + 1 < W ? m[k + 1] : 0u;      // a pixel never continues past the thread's last word
            out[k] = m[k] | __funnelshift_r(m[k], nx, 8) | __funnelshift_r(m[k], nx, 16);
        }
    }
}

// The four weight tables live in shared memory as 32 private copies, interleaved so that lane l only ever touches bank l:
// entry e of table t for lane l is word ((t*256 + e) << 5) + l.  Nine lookups per smoothed byte with uncorrelated indices
// conflicted 2.5 ways on a single 4 KB copy (61 % of all shared wavefronts, profiles/r1c_ncu_secondary_kernels.md); this
// layout is conflict-free by construction.  128 KB per block, so one persistent block per SM walks the whole batch.
constexpr int kAaTableWords = 1024 * 32;

template <int BPP, int NG, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_antialias_vec(const uint8_t *src0, size_t spitch, uint8_t *dst0, size_t dpitch,
                                                              const uint32_t *tables, uint32_t w, uint32_t h, uint32_t nframes,
                                                              FastDiv div_chunks, FastDiv div_upr)
{
    constexpr int W = BPP * NG, PX = 4 * NG;                // words / pixels per thread
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t *const s_t = s_dyn;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *const queue = s_dyn + kAaTableWords + warp * (32 * PX);     // per warp: byte offsets of the pixels that take the 9-tap path
    for (int i = threadIdx.x; i < kAaTableWords; i += THREADS) s_t[i] = __ldg(tables + (i >> 5));
    __syncthreads();
    const uint32_t *const my_t = s_t + lane;
    const uint32_t upr = w / PX, wpr = upr * W, n = upr * h;
    const uint32_t chunks_per_frame = (n + 31) / 32, chunks = chunks_per_frame * nframes;      // a chunk = 32 units of one frame
    const ptrdiff_t Bpl = (ptrdiff_t)w * BPP;
    constexpr int SHL = BPP == 1 ? 24 : 8, SHR = BPP == 1 ? 8 : 24;       // funnel shifts that move the row by BPP bytes
    auto load_words = [](const uint32_t *p, uint32_t *out) {
        if (NG == 4) {
#pragma unroll
            for (int k = 0; k < W / 4; k++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + k);
                out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) out[k] = __ldg(p + k);
        }
    };
    for (uint32_t c = blockIdx.x * (THREADS / 32) + warp; c < chunks; c += gridDim.x * (THREADS / 32)) {     // warp-uniform
        uint32_t frame, cf;
        div_chunks.divmod(c, frame, cf);
        const uint32_t i = cf * 32 + lane;
        const uint8_t *src = src0 + (size_t)frame * spitch;
        uint8_t *dst = dst0 + (size_t)frame * dpitch;
        uint32_t S[W], any = 0, y = 0, u = 0;
#pragma unroll
        for (int k = 0; k < W; k++) S[k] = 0;
        if (i < n) {
            div_upr.divmod(i, y, u);
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(src) + (size_t)y * wpr + (size_t)u * W;
            uint32_t C[W];
            load_words(cw, C);
            if (y > 0 && y < h - 1) {
                uint32_t U[W], D[W], L[W], R[W], A[W], B[W];
                load_words(cw - wpr, U);
                load_words(cw + wpr, D);
                const uint32_t prev = u > 0 ? __ldg(cw - 1) : 0u, next = u < upr - 1 ? __ldg(cw + W) : 0u;
#pragma unroll
                for (int k = 0; k < W; k++) {
                    L[k] = __funnelshift_r(k ? C[k - 1] : prev, C[k], SHL);
                    R[k] = __funnelshift_r(C[k], k + 1 < W ? C[k + 1] : next, SHR);
                }
                // smooth = LR differ && ((LU differ) != (LD differ) || (RU differ) != (RD differ))      tcvideo.c:945-949
                pixel_diff<BPP, W>(L, U, A);
                pixel_diff<BPP, W>(L, D, B);
#pragma unroll
                for (int k = 0; k < W; k++) A[k] ^= B[k];
                pixel_diff<BPP, W>(R, U, B);
                pixel_diff<BPP, W>(R, D, S);
#pragma unroll
                for (int k = 0; k < W; k++) A[k] |= B[k] ^ S[k];
                pixel_diff<BPP, W>(L, R, S);
#pragma unroll
                for (int k = 0; k < W; k++) {
                    // keep the bit on the first byte of each pixel only
                    constexpr uint32_t kFirst[3] = {0x80000080u, 0x00800000u, 0x00008000u};
                    S[k] &= A[k] & (BPP == 1 ? 0x80808080u : kFirst[k % 3]);
                }
                if (u == 0) S[0] &= ~0x80u;                                                  // x == 0 is copied (:942)
                if (u == upr - 1) S[W - 1] &= BPP == 1 ? ~0x80000000u : ~0x00008000u;        // so is x == w-1 (:968)
#pragma unroll
                for (int k = 0; k < W; k++) any |= S[k];
            }
            uint8_t *ow = dst + ((size_t)y * wpr + (size_t)u * W) * 4;
            if (NG == 4) {
#pragma unroll
                for (int k = 0; k < W / 4; k++) stg128(ow + 16 * k, make_uint4(C[4 * k], C[4 * k + 1], C[4 * k + 2], C[4 * k + 3]));
            } else {
#pragma unroll
                for (int k = 0; k < W; k++) stg32(ow + 4 * k, C[k]);
            }
        }
        // The pixels that ARE smoothed overwrite their bytes after the copies above (each byte is written by the lane that
        // smooths it; __syncwarp orders the two stores).  A warp with only a few of them lets the owning lanes do the 9-tap
        // work themselves; a warp with many gathers them into a dense queue first so the work runs on full warps.
        auto smooth_pixel = [&](uint32_t off) {
            const uint8_t *cc = src + off, *uu = cc - Bpl, *dd = cc + Bpl;
            uint8_t *o = dst + off;
#pragma unroll
            for (int k = 0; k < BPP; k++) {
                const uint32_t sum = my_t[(768u + uu[k - BPP]) << 5] + my_t[(512u + uu[k]) << 5] + my_t[(768u + uu[k + BPP]) << 5]
                                   + my_t[(256u + cc[k - BPP]) << 5] + my_t[(uint32_t)cc[k] << 5]   + my_t[(256u + cc[k + BPP]) << 5]
                                   + my_t[(768u + dd[k - BPP]) << 5] + my_t[(512u + dd[k]) << 5] + my_t[(768u + dd[k + BPP]) << 5]
                                   + 32768u;
                o[k] = (uint8_t)(sum >> 16);
            }
        };
        const uint32_t flagged = __ballot_sync(0xFFFFFFFFu, any != 0);
        if (flagged) {
            const uint32_t b0 = (y * wpr + u * W) * 4;            // byte offset of the thread's first word in the frame
            if (__popc(flagged) <= 3) {
                __syncwarp();
                if (any) {
#pragma unroll
                    for (int k = 0; k < W; k++)
                        for (uint32_t f = S[k]; f; f &= f - 1) smooth_pixel(b0 + (uint32_t)(4 * k + ((__ffs(f) - 1) >> 3)));
                }
                __syncwarp();
                continue;
            }
            uint32_t mine = 0;
#pragma unroll
            for (int k = 0; k < W; k++) mine += __popc(S[k]);
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (mine) {
                uint32_t slot = incl - mine;
#pragma unroll
                for (int k = 0; k < W; k++)
                    for (uint32_t f = S[k]; f; f &= f - 1) queue[slot++] = b0 + (uint32_t)(4 * k + ((__ffs(f) - 1) >> 3));
            }
            __syncwarp();
            for (uint32_t r = lane; r < total; r += 32) smooth_pixel(queue[r]);
            __syncwarp();
        }
    }
}

inline bool al16p(const void *p, size_t pitch, int nframes) { return al16(p) && (nframes <= 1 || pitch % 16 == 0); }

}  // namespace

bool tcv_window_launch(TcvWindow p, int nframes, cudaStream_t st)
{
    p.vec = al16p(p.dst, p.dpitch, nframes) ? 1 : 0;
    const uint64_t N = (uint64_t)p.dBpl * p.drows;
    if (N == 0) return true;
    p.row_div = make_fastdiv(p.dBpl);
    // every whole chunk's source 16-byte aligned?  (chunk at byte xb of destination row y reads source row
    // y * row_mul + row_add from byte sxb + xb - cl on)
    const bool aligned = p.vec && p.sBpl % 16 == 0 && p.dBpl % 16 == 0 && (nframes <= 1 || p.spitch % 16 == 0)
                      && ((reinterpret_cast<uintptr_t>(p.src) + p.sxb - p.cl) & 15) == 0;
    const dim3 g = grid_for((N + 15) / 16 / kWinUnroll + 1, nframes);
    if (aligned) k_window<false><<<g, kThreads, 0, st>>>(p);
    else k_window<true><<<g, kThreads, 0, st>>>(p);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_window");
    return true;
}

bool tcv_reduce_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int ow, int oh, int rw, int rh,
                       int Bpp, int nframes, cudaStream_t st)
{
    const uint64_t n = (uint64_t)ow * oh;
    if (n == 0) return true;
    if (rw == 2 && w % 32 == 0 && al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes)) {
        const dim3 g2 = grid_for((uint64_t)(w / 32) * oh, nframes);
        if (Bpp == 1) k_reduce_w2<1><<<g2, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, oh, rh);
        else k_reduce_w2<3><<<g2, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, oh, rh);
        note_launch();
        ACGPU_CHECK_LAUNCH("k_reduce_w2");
        return true;
    }
    if (rw >= 3 && rw <= 6 && ow % 16 == 0 && ((size_t)w * Bpp) % 16 == 0 && al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes)) {
        // a kept row is fetched whole (every 32-byte sector of it holds kept pixels up to ratio 32 at Bpp 1), so whole-row
        // 128-bit loads + compile-time byte selection are also the DRAM-optimal form -- up to ratio 6: a lane's loads lie
        // 16 * ratio bytes from its neighbour's, so each load instruction touches 4 * ratio cache lines, and at 8 that costs
        // more than the byte gather does (8x8 luma: 0.30 of the copy rate in DRAM bytes against ~0.6; 5x5: 0.92 against 0.62)
        const dim3 gn = grid_for((uint64_t)(ow / 16) * oh, nframes);
        auto go = [&](auto kern) { kern<<<gn, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, ow, oh, rh); };
#define ACGPU_REDUCE_CASE(R) case R: if (Bpp == 1) go(k_reduce_wn<1, R>); else go(k_reduce_wn<3, R>); break;
        switch (rw) {
            ACGPU_REDUCE_CASE(3) ACGPU_REDUCE_CASE(4) ACGPU_REDUCE_CASE(5) ACGPU_REDUCE_CASE(6)
        }
#undef ACGPU_REDUCE_CASE
        note_launch();
        ACGPU_CHECK_LAUNCH("k_reduce_wn");
        return true;
    }
    const dim3 g = grid_for((n + 3) / 4, nframes);
    const int vec = ((uintptr_t)dst & 3) == 0 && (nframes <= 1 || dpitch % 4 == 0);
    if (Bpp == 1) k_reduce<1><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, ow, oh, rw, rh, vec);
    else k_reduce<3><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, ow, oh, rw, rh, vec);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_reduce");
    return true;
}

bool tcv_flip_v_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int h, int Bpp, int nframes, cudaStream_t st)
{
    const uint32_t Bpl = (uint32_t)w * Bpp;
    const int vec = Bpl % 16 == 0 && al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes);
    const uint64_t n = (uint64_t)((h + 1) / 2) * ((Bpl + 15) / 16);
    k_flip_v<2><<<grid_for(n / 2 + 1, nframes), kThreads, 0, st>>>(src, spitch, dst, dpitch, Bpl, h, vec);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_flip_v");
    return true;
}

bool tcv_flip_h_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int h, int Bpp, int nframes, cudaStream_t st)
{
    const int vec = w % 16 == 0 && al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes);
    const uint64_t n = vec ? (uint64_t)((w / 16 + 1) / 2) * h : (uint64_t)((w + 1) / 2) * h;
    const dim3 g = grid_for(n, nframes);
    if (Bpp == 1) k_flip_h<1><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, h, vec);
    else k_flip_h<3><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, w, h, vec);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_flip_h");
    return true;
}

bool tcv_lut_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint8_t *d_table, size_t nbytes,
                    int nframes, cudaStream_t st)
{
    const int vec = al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes);
    k_lut<<<grid_for((nbytes + 15) / 16, nframes, 8), kThreads, 0, st>>>(src, spitch, dst, dpitch, d_table, (uint32_t)nbytes, vec);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_lut");
    return true;
}

bool tcv_antialias_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint32_t *d_tables, int w, int h,
                          int Bpp, int nframes, cudaStream_t st)
{
    const bool vec = w % 4 == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 3) == 0
                  && (nframes <= 1 || (spitch % 4 == 0 && dpitch % 4 == 0));
    if (vec) {
        const bool wide = w % 16 == 0 && al16p(src, spitch, nframes) && al16p(dst, dpitch, nframes);
        // one persistent block per SM (its 128 KB of replicated tables fill most of the SM's shared memory); warps take
        // chunks of 32 units of one frame in turn, so the blocks stay balanced whatever the number of frames is
        const uint32_t units = (uint32_t)(w / (wide ? 16 : 4)) * (uint32_t)h, chunks_per_frame = (units + 31) / 32;
        auto launch = [&](auto kern, int threads, int px) {
            const size_t smem = ((size_t)kAaTableWords + (size_t)(threads / 32) * 32 * px) * sizeof(uint32_t);
            // (per device and cheap: set on every launch rather than tracked per thread and device)
            if (!check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "k_antialias_vec smem")) return false;
            const uint64_t chunks = (uint64_t)chunks_per_frame * nframes;
            const uint64_t want = (chunks + threads / 32 - 1) / (threads / 32);
            const unsigned gx = (unsigned)(want < (uint64_t)sm_count() ? want : (uint64_t)sm_count());
            kern<<<gx, threads, smem, st>>>(src, spitch, dst, dpitch, d_tables, (uint32_t)w, (uint32_t)h, (uint32_t)nframes,
                                            make_fastdiv(chunks_per_frame), make_fastdiv((uint32_t)(w / (wide ? 16 : 4))));
            return true;
        };
        bool ok;
        if (wide && Bpp == 1) ok = launch(k_antialias_vec<1, 4, 1024>, 1024, 16);
        else if (wide) ok = launch(k_antialias_vec<3, 4, 512>, 512, 16);
        else if (Bpp == 1) ok = launch(k_antialias_vec<1, 1, 1024>, 1024, 4);
        else ok = launch(k_antialias_vec<3, 1, 512>, 512, 4);
        if (!ok) return false;
        note_launch();
        ACGPU_CHECK_LAUNCH("k_antialias_vec");
        return true;
    }
    const dim3 g = grid_for((uint64_t)w * h, nframes, 8);
    if (Bpp == 1) k_antialias<1><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, d_tables, w, h);
    else k_antialias<3><<<g, kThreads, 0, st>>>(src, spitch, dst, dpitch, d_tables, w, h);
    note_launch();
    ACGPU_CHECK_LAUNCH("k_antialias");
    return true;
}

}  // namespace acgpu
