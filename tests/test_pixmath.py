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
