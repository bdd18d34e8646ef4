/*
 * ac_oracle.h -- CPU restatement of aclib's plain-C pixel path (TEST INFRASTRUCTURE, not product).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  libacgpu never links, loads or calls it: the product path is CUDA-only and fails
 * loudly without a device.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every one of the 256 reachable format pairs
 * byte-for-byte against the unmodified reference compiled into oracle/_ref/libac_ref_c.so
 * (ac_init(AC_NONE)) whenever that file is present, plus the reference's own known-answer vectors
 * (testsuite/test-average.c:177-645, testsuite/newtest.pl:1462-1537 colour bars) and the committed
 * digests in tests/golden/ that were generated from oracle/_ref (tests/golden/make_golden.py).
 *
 * Every symbol is prefixed oracle_ so it can sit in one process beside libacgpu and oracle/_ref.
 */
#ifndef AC_ORACLE_H
#define AC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric format ids as aclib/imgconvert.h:16-40. */
int oracle_imgconvert(uint8_t **src, int srcfmt, uint8_t **dest, int destfmt, int width, int height);

/* aclib/average.c:33-39 */
void oracle_average(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes);

/* aclib/rescale.c:23-46 */
void oracle_rescale(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes,
                    uint32_t weight1, uint32_t weight2);

/* libtcvideo/tcvideo.c:1138-1165 -- the vertical/horizontal resize weight table (host-side doubles). */
void oracle_resize_table(int oldsize, int newsize, int32_t *source, uint32_t *weight1, uint32_t *weight2);

/* libtcvideo/tcvideo.c:340-389 -- deinterlace shapes built from ac_average/ac_memcpy.
 * mode 0 = interpolate, 1 = linear blend (destroys src exactly as the reference does),
 * 2 = drop the top field, 3 = drop the bottom field (height/2 rows written). */
int oracle_deinterlace(uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int mode);

/* libtcvideo/tcvideo.c:427-531 -- tcv_resize (vertical via ac_rescale, horizontal scalar). */
int oracle_resize(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                  int resize_w, int resize_h, int scale_w, int scale_h);

/* libtcvideo/tcvideo.c:184-250, 681-717, 739-818, 840-858, 886-980 -- the remaining element-wise plane operations
 * (SURVEY.md 8f row 3).  Return 1 on success, 0 on the parameter errors the reference rejects. */
int oracle_clip(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                int left, int right, int top, int bottom, uint8_t black);
int oracle_reduce(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h);
int oracle_flip_v(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp);
int oracle_flip_h(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp);
void oracle_gamma_table(double gamma, uint8_t *table /* [256] */);
int oracle_gamma_correct(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma);
void oracle_aa_tables(double weight, double bias, uint32_t *tables /* [1024]: c, x, y, d */);
int oracle_antialias(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias);

#ifdef __cplusplus
}
#endif
#endif
