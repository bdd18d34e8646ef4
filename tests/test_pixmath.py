"""Proves the reduced arithmetic of csrc/pixmath.cuh equal to the literal C arithmetic, exhaustively, on the CPU
(the header is host-compilable).  The -m gpu tests then check the kernels that use it against the oracle."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include <cstdio>
#include "pixmath.cuh"
using namespace pixmath;
int main() {
    long bad = 0;
    for (int i = -4096; i < 8192; i++) bad += ylut_fast(i) != ylut_ref(i);
    for (int i = 0; i < 256; i++) {
        bad += (int)(y2gray_word_fast(i) >> 24) != y2gray(i);
        bad += (int)(gray2y_word_fast(i) >> 24) != gray2y(i);
    }
    // the fast YUV->RGB channel form over the full cube, against the literal form
    for (int Y = 0; Y < 256; Y++)
        for (int U = 0; U < 256; U++)
            for (int V = 0; V < 256; V += 1) {
                RGB o = yuv2rgb_ref(Y, U, V);
                int r = (int)(ylut_word_fast(16 * Y + chroma_term_ref(kCRV, V) - 256) >> 24);
                int g = (int)(ylut_word_fast(16 * Y + chroma_term_ref(kCGU, U) + chroma_term_ref(kCGV, V) - 256) >> 24);
                int b = (int)(ylut_word_fast(16 * Y + chroma_term_ref(kCBU, U) - 256) >> 24);
                bad += (r != o.r) + (g != o.g) + (b != o.b);
            }
    // accumulator forms of RGB->YUV: byte 2 of (sum + 32768 + (offset<<16))
    for (int r = 0; r < 256; r += 3) for (int g = 0; g < 256; g += 5) for (int b = 0; b < 256; b += 7) {
        unsigned ay = 16829u*r + 33039u*g + 6416u*b + 32768u + (16u << 16);
        int au = -9714*r - 19070*g + 28784*b + 32768 + (128 << 16);
        int av = 28784*r - 24103*g - 4681*b + 32768 + (128 << 16);
        bad += (int)((ay >> 16) & 0xFF) != rgb2y(r, g, b);
        bad += ((au >> 16) & 0xFF) != rgb2u(r, g, b);
        bad += ((av >> 16) & 0xFF) != rgb2v(r, g, b);
        bad += (au >> 24) != 0 || (av >> 24) != 0 || (ay >> 24) != 0;
    }
    printf("%ld\n", bad);
    return bad != 0;
}
'''


def test_reduced_forms_equal_literal_forms():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(PROG)
        exe = os.path.join(d, "t")
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "transcode-tcforge_b200", "csrc"), "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout


DIV_PROG = r'''
#include "acgpu_internal.h"
#include <cstdio>
#include <cstdint>
#include <random>
// The kernels' division: q = umulhi(x, m) >> s, one compare fixes an estimate that is one too small (FastDiv::divmod).
static bool check(uint32_t d, uint32_t x)
{
    const acgpu::FastDiv f = acgpu::make_fastdiv(d);
    uint32_t q = (uint32_t)(((uint64_t)x * f.m) >> 32) >> f.s, r = x - q * d;
    if (r >= d) { q++; r -= d; }
    return q == x / d && r == x % d;
}
int main()
{
    std::mt19937_64 rng(20261019);
    long bad = 0;
    const uint32_t edge[] = {0u, 1u, 2u, 15u, 16u, 17u, 0x7FFFFFFFu, 0x80000000u, 0x80000001u, 0xFFFFFFF0u, 0xFFFFFFFFu};
    for (uint32_t d = 1; d <= 70000; d++) {
        for (uint32_t x : edge) bad += !check(d, x);
        for (uint32_t k : {d - 1, d, d + 1, 2 * d - 1, 2 * d, 0xFFFFFFFFu / d * d, 0xFFFFFFFFu / d * d - 1}) bad += !check(d, k);
        for (int i = 0; i < 8; i++) bad += !check(d, (uint32_t)rng());
    }
    for (int i = 0; i < 2000000; i++) {
        uint32_t d = (uint32_t)rng();
        if (!d) d = 1;
        bad += !check(d, (uint32_t)rng());
        bad += !check(d, d - 1) + !check(d, d) + !check(d, 0xFFFFFFFFu);
    }
    for (int s = 0; s < 32; s++) for (int o = -1; o <= 1; o++) {       // powers of two and their neighbours
        const uint32_t d = (1u << s) + (uint32_t)o;
        if (!d) continue;
        for (int i = 0; i < 1000; i++) bad += !check(d, (uint32_t)rng());
        for (uint32_t x : edge) bad += !check(d, x);
    }
    printf("%ld\n", bad);
    return bad != 0;
}
'''


def test_multiply_high_division_is_exact():
    """FastDiv (acgpu_internal.h): the row / frame indices of the window copy and antialias kernels come from a
    multiply-high with a host-made reciprocal instead of a division; exact for every 32-bit dividend."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write(DIV_PROG)
        exe = os.path.join(d, "t")
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "transcode-tcforge_b200", "csrc"),
                        "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout
