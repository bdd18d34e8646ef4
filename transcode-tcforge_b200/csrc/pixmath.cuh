// pixmath.cuh -- bit-exact integer pixel arithmetic of aclib's plain-C path, for device (and host) code.
//
// Two statements of the same functions:
//   *_ref  : literal transcription of the C arithmetic (truncating divisions, arithmetic shifts,
//            table formulas) -- used by the generic tier, where obviousness beats speed;
//   *_fast : algebraically reduced forms used by the vectorised tiers.  Each reduction is proven
//            exhaustively against the literal form by tests/test_pixmath.py (host build of this file)
//            and against the oracle on the full 2^24 cubes by the -m gpu tests.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD static inline
#endif

namespace pixmath {

// ---- constants (aclib/img_yuv_rgb.c:25-29,34) ------------------------------------------------
constexpr int kCY = 76309, kCRV = 104597, kCGU = -25675, kCGV = -53279, kCBU = 132201;
constexpr int kYScale = 16;

PM_HD int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

// ---- YUV -> RGB, literal (aclib/img_yuv_rgb.c:41-57) -----------------------------------------
// Chroma offset tables: ((coef*(c-128))*16 + cY/2) / cY with C's truncating division.
PM_HD int chroma_term_ref(int coef, int c) { return (coef * (c - 128) * kYScale + kCY / 2) / kCY; }
// Ylut[i], i in [-4096, 8191]: clamp(((cY*(i-256))/16 + 32768) >> 16)
PM_HD int ylut_ref(int i) { return clamp255(((kCY * (i - 16 * kYScale) / kYScale) + 32768) >> 16); }

struct RGB { int r, g, b; };

PM_HD RGB yuv2rgb_ref(int Y, int U, int V)
{
    const int y16 = Y * kYScale;
    RGB o;
    o.r = ylut_ref(y16 + chroma_term_ref(kCRV, V));
    o.g = ylut_ref(y16 + chroma_term_ref(kCGU, U) + chroma_term_ref(kCGV, V));
    o.b = ylut_ref(y16 + chroma_term_ref(kCBU, U));
    return o;
}

// ---- YUV -> RGB, reduced ----------------------------------------------------------------------
// With j = i - 256:  Ylut[i] == (clamp(j, 0, 3498) * 1220944 + 8388608) >> 24   for every table index.
//   * j >= 0: floor(floor(cY*j/16) + 32768) / 65536) == (cY*j + 2^19) >> 20 (nested floors);
//   * j <  0: both forms clamp to 0 (the truncating division only matters for j in [-6,-1], where
//     the sum is still below 65536);
//   * the result saturates at 255 from j = 3498 on, and 3498*1220944 + 2^23 < 2^32, so after clamping
//     j the product can be formed in uint32 and the answer is simply the TOP BYTE of the word.
constexpr int      kJMax = 3498;
constexpr uint32_t kJMul = 1220944u;   // cY * 16
constexpr uint32_t kJAdd = 8388608u;   // 2^23
PM_HD uint32_t ylut_word_fast(int j)   // answer in bits 31..24
{
    const int jc = j < 0 ? 0 : j > kJMax ? kJMax : j;
    return (uint32_t)jc * kJMul + kJAdd;
}
PM_HD int ylut_fast(int i) { return (int)(ylut_word_fast(i - 256) >> 24); }

// ---- RGB -> YUV (aclib/img_yuv_rgb.c:142-153) ---------------------------------------------------
PM_HD int rgb2y(int r, int g, int b) { return ((16829 * r + 33039 * g + 6416 * b + 32768) >> 16) + 16; }
PM_HD int rgb2u(int r, int g, int b) { return ((-9714 * r - 19070 * g + 28784 * b + 32768) >> 16) + 128; }
PM_HD int rgb2v(int r, int g, int b) { return ((28784 * r - 24103 * g - 4681 * b + 32768) >> 16) + 128; }

// ---- gray maps (aclib/img_yuv_rgb.c:230-245) and RGB -> gray (aclib/img_rgb_packed.c:179-303) ----
PM_HD int y2gray(int i) { return i <= 16 ? 0 : i >= 235 ? 255 : (i - 16) * 255 / 219; }
PM_HD int gray2y(int i) { return 16 + i * 219 / 255; }
PM_HD int rgb2gray(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }

// Reduced forms: both maps are a clamp plus ONE multiply whose TOP BYTE is the answer.
//   floor(x*255/219) == (x*19535115) >> 24 for x in [0,219];  floor(i*219/255) == (i*14408668) >> 24 for i in [0,255]
constexpr uint32_t kY2GrayMul = 19535115u, kGray2YMul = 14408668u;
PM_HD uint32_t y2gray_word_fast(int i)     // answer in bits 31..24
{
    int x = i - 16;
    x = x < 0 ? 0 : x > 219 ? 219 : x;
    return (uint32_t)x * kY2GrayMul;
}
PM_HD uint32_t gray2y_word_fast(int i) { return (uint32_t)i * kGray2YMul + (16u << 24); }

// ---- blends (aclib/average.c:37-38, aclib/rescale.c:44-45) ---------------------------------------
PM_HD int avg2(int a, int b) { return (a + b + 1) / 2; }
PM_HD int avg2_trunc(int a, int b) { return (a + b) / 2; }           // aclib/img_yuv_mixed.c:135,137
PM_HD int avg4(int a, int b, int c, int d) { return (a + b + c + d + 2) / 4; }
PM_HD int rescale1(int a, int b, uint32_t w1, uint32_t w2)
{
    return (int)(((uint32_t)a * w1 + (uint32_t)b * w2 + 32768u) >> 16) & 0xFF;
}

}  // namespace pixmath
