// acgpu_internal.h -- declarations shared by libacgpu's translation units (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "acgpu.h"

namespace acgpu {

// ---------------------------------------------------------------------------------------------
// Format descriptors.  One struct describes any of the 15 layouts; kernels take it by value.
// Offsets follow the reference tables: packed YUV (aclib/img_yuv_rgb.c:104-106), RGB byte orders
// (aclib/img_yuv_rgb.c:129-134), planar subsampling (aclib/imgconvert.h:54-59).

enum Kind : int { K_NONE = 0, K_PLANAR, K_Y8, K_PACKED, K_RGB, K_GRAY };

struct FmtDesc {
    int kind;
    int sx, sy;              // K_PLANAR: log2 chroma subsampling (x, y)
    int yo, uo, vo;          // K_PACKED: byte offsets inside the 4-byte (2-pixel) group
    int bpp, ro, go, bo, ao; // K_RGB: bytes per pixel and channel offsets, ao < 0 when no alpha
};

FmtDesc describe(int fmt);
size_t  frame_bytes(int fmt, int w, int h);
size_t  chroma_plane_bytes(int fmt, int w, int h);
int     nplanes(int fmt);

// Batch geometry handed to every conversion launcher: plane pointers of frame 0 and the byte
// distance between consecutive frames (one pitch for all planes of an image).
struct Image {
    uint8_t *p[3];
    size_t   pitch;
};

struct ConvertArgs {
    Image src, dst;
    int   srcfmt, dstfmt;    // YV12 already folded into YUV420P with swapped planes
    int   w, h, nframes;
    cudaStream_t stream;
};

// Tier 1: any size, any alignment, literal loop bounds (kernels_generic.cu).
bool convert_generic(const ConvertArgs &a);
// Tier 2: 16-byte vectorised kernels; returns false (without launching) when the pair/size/alignment
// is outside its domain so the caller can fall back to tier 1 (kernels_fast.cu).
bool convert_fast(const ConvertArgs &a);
// Tier 3: the tier-2 YUV->RGB24 kernels with bulk (TMA) stores of the output tile (kernels_fast.cu); selectable only.
bool convert_tma(const ConvertArgs &a);
// The tier-3 variant that is chosen automatically where it beats tier 2 (YUV420P -> RGB, tensor-map staged loads); returns
// false without launching when the call is outside its domain.
bool convert_tma_auto(const ConvertArgs &a);

// Fused YUV420P -> RGB (any layout, never materialised) -> YUV420P / 422P / 444P for frame chains (kernels_fast.cu);
// a.src / a.dst are the outer batches.  false (no launch, no error) when outside its domain.
bool convert_fused_yuv420_rgb_yuv(const ConvertArgs &a);

// Fused RGB24 -> gray -> RGB24 in place (kernels_fast_rgb.cu); false when outside the vectorised domain.
bool decolor_rgb24_fast(uint8_t *frames, size_t pitch, int w, int h, int nframes, cudaStream_t st);

// Row blends (rowops.cu).
// One block of consecutive row operations with the distinct source rows it reads (host-built, see rowops.cu).
struct RowTask {
    int64_t  dest_off;
    uint32_t w1, w2;
    uint8_t  op, s1, s2, s3;      // ACGPU_ROW_* and slots into RowBlk::src_off
    uint32_t pad;
};
struct RowBlk {
    static constexpr int kRows = 8, kMaxSrc = 3 * kRows;
    int32_t nsrc, nops;
    int64_t src_off[kMaxSrc];
    RowTask t[kRows];
};
int  build_row_blocks(const acgpu_rowop *ops, int nops, std::vector<RowBlk> &out);   // returns max distinct rows per block
bool rowops_tiled_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const RowBlk *d_blks, int nblks,
                         int max_nsrc, int row_bytes, int nframes, cudaStream_t st);
bool rowops_bytes_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                         const acgpu_rowop *d_ops, int nops, int row_bytes, int nframes, cudaStream_t st);
bool rowops_vectorisable(const uint8_t *src, size_t spitch, const uint8_t *dst, size_t dpitch, int row_bytes);
bool blend_launch(const uint8_t *s1, const uint8_t *s2, uint8_t *d, size_t bytes,
                  uint32_t w1, uint32_t w2, int op, cudaStream_t st);
bool resize_h_row_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint16_t *d_off,
                         const uint32_t *d_wgt, int width, int new_w, int rows, int Bpp, int nframes, cudaStream_t st);
bool resize_h_win_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint2 *d_meta,
                         const uint32_t *d_wgt, int width, int new_w, int rows, int Bpp, bool narrow, int nframes, cudaStream_t st);
bool resize_h_vectorisable(const uint8_t *src, size_t spitch, const uint8_t *dst, size_t dpitch, int width, int new_w, int Bpp);
bool resize_h_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch,
                     const int32_t *d_source, const uint32_t *d_w1, const uint32_t *d_w2,
                     int width, int new_w, int new_h, int Bpp, int scale_w, int nframes, cudaStream_t st);

// libtcvideo plane operations (tcvops.cu).
// Window copy: destination row y shows bytes [sxb, sxb+cn) of source row y*row_mul + row_add at its bytes [cl, cl+cn)
// and the fill byte elsewhere (also in rows whose source row is outside [0, srows)).
// x / d for any 32-bit x without a division instruction: m = floor(2^(32+s) / d) (capped at 2^32 - 1), s = floor(log2 d); the
// multiply-high estimate is at most one too small, one compare fixes it.  Built on the host, used by the kernels.
struct FastDiv {
    uint32_t d, m, s;
#ifdef __CUDACC__
    __device__ __forceinline__ void divmod(uint32_t x, uint32_t &q, uint32_t &r) const
    {
        q = __umulhi(x, m) >> s;
        r = x - q * d;
        if (r >= d) { q++; r -= d; }
    }
#endif
};
inline FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f{d ? d : 1u, 0, 0};
    while ((2u << f.s) <= f.d && f.s < 31) f.s++;
    const uint64_t m = ((uint64_t)1 << (32 + f.s)) / f.d;
    f.m = m > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)m;
    return f;
}

struct TcvWindow {
    const uint8_t *src;
    size_t   spitch;
    uint8_t *dst;
    size_t   dpitch;
    uint32_t dBpl, sBpl;          // bytes per destination / source row
    int      drows, srows;
    int      row_mul, row_add;
    uint32_t cl, cn, sxb;
    uint32_t fill;                // fill byte replicated into all four bytes
    int      vec;                 // set by the launcher: destination chunks are 16-byte aligned
    FastDiv  row_div;             // set by the launcher: off / dBpl
};
bool tcv_window_launch(TcvWindow p, int nframes, cudaStream_t st);
bool tcv_reduce_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int ow, int oh, int rw, int rh,
                       int Bpp, int nframes, cudaStream_t st);
bool tcv_flip_v_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int h, int Bpp, int nframes, cudaStream_t st);
bool tcv_flip_h_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, int w, int h, int Bpp, int nframes, cudaStream_t st);
bool tcv_lut_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint8_t *d_table, size_t nbytes,
                    int nframes, cudaStream_t st);
bool tcv_antialias_launch(const uint8_t *src, size_t spitch, uint8_t *dst, size_t dpitch, const uint32_t *d_tables, int w, int h,
                          int Bpp, int nframes, cudaStream_t st);

// Per-thread bookkeeping (host_api.cu).
void     note_launch(int n = 1);
void     set_error(const char *fmt, ...);
bool     check(cudaError_t e, const char *what);
int      sm_count();

}  // namespace acgpu

#define ACGPU_CHECK_LAUNCH(what)                                   \
    do {                                                           \
        if (!::acgpu::check(cudaGetLastError(), what)) return false; \
    } while (0)
