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
#include "pixmath.cuh"

namespace acgpu {
namespace {

constexpr int TPB = 256;

__device__ __forceinline__ uint32_t avg4x8(uint32_t a, uint32_t b)
{
    // per-byte (a + b + 1) >> 1 without carries between lanes
    return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu);
}

__device__ __forceinline__ uint32_t rescale4x8(uint32_t a, uint32_t b, uint32_t w1, uint32_t w2)
{
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t x = (a >> (8 * k)) & 0xFF, y = (b >> (8 * k)) & 0xFF;
        r |= (((x * w1 + y * w2 + 32768u) >> 16) & 0xFFu) << (8 * k);
    }
    return r;
}

__device__ __forceinline__ uint4 ld16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st16(uint8_t *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

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

// One thread per 16-byte chunk; chunks_per_row = row_bytes/16.
__global__ void k_rowops_v16(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                             const acgpu_rowop *__restrict__ ops, int nops, int chunks_per_row)
{
    const size_t idx = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (idx >= (size_t)nops * chunks_per_row) return;
    const int row = (int)(idx / chunks_per_row);
    const size_t c = (idx - (size_t)row * chunks_per_row) * 16;
    const acgpu_rowop op = ops[row];
    const uint8_t *s = src + (size_t)blockIdx.y * spitch;
    st16(dst + (size_t)blockIdx.y * dpitch + op.dest_off + c, blend16(op, s, c));
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

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

bool rowops_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                   const acgpu_rowop *d_ops, int nops, int row_bytes, int nframes, cudaStream_t st,
                   bool offsets_aligned16)
{
    if (nops <= 0 || row_bytes <= 0 || nframes <= 0) return true;
    const bool vec = offsets_aligned16 && aligned16(src) && aligned16(dst) && spitch % 16 == 0 && dpitch % 16 == 0
                  && row_bytes % 16 == 0;
    if (vec) {
        const int cpr = row_bytes / 16;
        const size_t n = (size_t)nops * cpr;
        dim3 g((unsigned)((n + TPB - 1) / TPB), (unsigned)nframes);
        k_rowops_v16<<<g, TPB, 0, st>>>(src, spitch, dst, dpitch, d_ops, nops, cpr);
    } else {
        const size_t n = (size_t)nops * row_bytes;
        dim3 g((unsigned)((n + TPB - 1) / TPB), (unsigned)nframes);
        k_rowops_bytes<<<g, TPB, 0, st>>>(src, spitch, dst, dpitch, d_ops, nops, row_bytes);
    }
    note_launch();
    ACGPU_CHECK_LAUNCH("k_rowops");
    return true;
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
