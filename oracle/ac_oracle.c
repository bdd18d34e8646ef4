/*
 * ac_oracle.c -- CPU restatement of aclib's plain-C pixel path.  TEST INFRASTRUCTURE ONLY (see
 * ac_oracle.h): the checker for libacgpu's CUDA kernels, never the thing measured or shipped.
 *
 * This is a from-scratch, descriptor-driven statement of the arithmetic that the reference spreads
 * over macro-generated functions.  Each block cites the reference file:line it follows (paths are
 * relative to /root/reference).  It is pinned byte-for-byte to the compiled reference
 * (oracle/_ref/libac_ref_c.so) by tests/test_oracle.py; nothing here is trusted without that.
 *
 * Conventions that matter for bit-exactness:
 *   - `/` is C integer division (truncates toward zero); `>>` on int is arithmetic.
 *   - images are tightly packed, no stride (aclib/imgconvert.h:54-65).
 *   - loop bounds are reproduced literally so that "bytes the C path leaves untouched" match too.
 */
#include "ac_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Format descriptors (aclib/imgconvert.h:16-40 for the ids)                                    */

enum {
    F_YUV420P = 0x1001, F_YV12, F_YUV411P, F_YUV422P, F_YUV444P, F_YUY2, F_UYVY, F_YVYU, F_Y8,
    F_RGB24 = 0x2001, F_BGR24, F_RGBA32, F_ABGR32, F_ARGB32, F_BGRA32, F_GRAY8
};

enum { K_NONE, K_PLANAR, K_Y8, K_PACKED, K_RGB, K_GRAY };

typedef struct {
    int kind;
    int sx, sy;         /* K_PLANAR: log2 chroma subsampling */
    int yo, uo, vo;     /* K_PACKED: byte offsets inside the 4-byte / 2-pixel group */
    int bpp, ro, go, bo, ao; /* K_RGB: bytes per pixel and channel offsets (ao<0: no alpha) */
} fmtdesc;

static fmtdesc describe(int fmt)
{
    fmtdesc d;
    memset(&d, 0, sizeof(d));
    d.ao = -1;
    switch (fmt) {
    case F_YUV420P: d.kind = K_PLANAR; d.sx = 1; d.sy = 1; break;
    case F_YUV411P: d.kind = K_PLANAR; d.sx = 2; d.sy = 0; break;
    case F_YUV422P: d.kind = K_PLANAR; d.sx = 1; d.sy = 0; break;
    case F_YUV444P: d.kind = K_PLANAR; d.sx = 0; d.sy = 0; break;
    case F_Y8:      d.kind = K_Y8; break;
    /* img_yuv_rgb.c:104-106 */
    case F_YUY2: d.kind = K_PACKED; d.yo = 0; d.uo = 1; d.vo = 3; break;
    case F_UYVY: d.kind = K_PACKED; d.yo = 1; d.uo = 0; d.vo = 2; break;
    case F_YVYU: d.kind = K_PACKED; d.yo = 0; d.uo = 3; d.vo = 1; break;
    /* img_yuv_rgb.c:129-134 */
    case F_RGB24:  d.kind = K_RGB; d.bpp = 3; d.ro = 0; d.go = 1; d.bo = 2; break;
    case F_BGR24:  d.kind = K_RGB; d.bpp = 3; d.ro = 2; d.go = 1; d.bo = 0; break;
    case F_RGBA32: d.kind = K_RGB; d.bpp = 4; d.ro = 0; d.go = 1; d.bo = 2; d.ao = 3; break;
    case F_ABGR32: d.kind = K_RGB; d.bpp = 4; d.ro = 3; d.go = 2; d.bo = 1; d.ao = 0; break;
    case F_ARGB32: d.kind = K_RGB; d.bpp = 4; d.ro = 1; d.go = 2; d.bo = 3; d.ao = 0; break;
    case F_BGRA32: d.kind = K_RGB; d.bpp = 4; d.ro = 2; d.go = 1; d.bo = 0; d.ao = 3; break;
    case F_GRAY8:  d.kind = K_GRAY; break;
    default: d.kind = K_NONE; break;
    }
    return d;
}

static long chroma_plane_bytes(const fmtdesc *d, int w, int h)
{
    return (long)(w >> d->sx) * (h >> d->sy);   /* imgconvert.h:54-59 */
}

/* ------------------------------------------------------------------------------------------ */
/* Lookup tables (img_yuv_rgb.c:25-57 and :227-245)                                             */

#define YSCALE 16
static int    lut_ready;
static int    ylut_store[768 * YSCALE];
static int   *ylut = ylut_store + 256 * YSCALE;
static int    rv_tab[256], gu_tab[256], gv_tab[256], bu_tab[256];
static uint8_t y2gray[256], gray2y[256];

static void build_tables(void)
{
    const int cY = 76309, crV = 104597, cgU = -25675, cgV = -53279, cbU = 132201;
    int i;
    if (lut_ready)
        return;
    for (i = -256 * YSCALE; i < 512 * YSCALE; i++) {
        int v = ((cY * (i - 16 * YSCALE) / YSCALE) + 32768) >> 16;
        ylut[i] = v < 0 ? 0 : v > 255 ? 255 : v;
    }
    for (i = 0; i < 256; i++) {
        int c = i - 128;
        rv_tab[i] = (crV * c * YSCALE + cY / 2) / cY;
        gu_tab[i] = (cgU * c * YSCALE + cY / 2) / cY;
        gv_tab[i] = (cgV * c * YSCALE + cY / 2) / cY;
        bu_tab[i] = (cbU * c * YSCALE + cY / 2) / cY;
        y2gray[i] = i <= 16 ? 0 : i >= 235 ? 255 : (i - 16) * 255 / 219;
        gray2y[i] = 16 + i * 219 / 255;
    }
    lut_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* YUV -> RGB (img_yuv_rgb.c:58-136).  Alpha is NOT written.                                    */

static void fetch_yuv(uint8_t **s, const fmtdesc *sd, int w, int x, int y, int *Y, int *U, int *V)
{
    if (sd->kind == K_PLANAR) {
        long ci = (long)(y >> sd->sy) * (w >> sd->sx) + (x >> sd->sx);  /* :100-103 */
        *Y = s[0][(long)y * w + x];
        *U = s[1][ci];
        *V = s[2][ci];
    } else {
        long cell = ((long)y * w + (x & ~1)) * 2;                       /* :66-68 */
        *Y = s[0][((long)y * w + x) * 2 + sd->yo];
        *U = s[0][cell + sd->uo];
        *V = s[0][cell + sd->vo];
    }
}

static void yuv_to_rgb(uint8_t **s, const fmtdesc *sd, uint8_t **d, const fmtdesc *dd, int w, int h)
{
    int x, y;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            int Y, U, V;
            uint8_t *px = d[0] + ((long)y * w + x) * dd->bpp;
            fetch_yuv(s, sd, w, x, y, &Y, &U, &V);
            Y *= YSCALE;
            px[dd->ro] = ylut[Y + rv_tab[V]];
            px[dd->go] = ylut[Y + gu_tab[U] + gv_tab[V]];
            px[dd->bo] = ylut[Y + bu_tab[U]];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* RGB -> YUV (img_yuv_rgb.c:142-221): every pixel gives Y, chroma is point-sampled.            */

static void rgb_to_yuv(uint8_t **s, const fmtdesc *sd, uint8_t **d, int destfmt, const fmtdesc *dd,
                       int w, int h)
{
    int x, y;
    for (y = 0; y < h; y++) {
        for (x = 0; x < w; x++) {
            const uint8_t *px = s[0] + ((long)y * w + x) * sd->bpp;
            int r = px[sd->ro], g = px[sd->go], b = px[sd->bo];
            int Y = ((16829 * r + 33039 * g + 6416 * b + 32768) >> 16) + 16;
            int U = ((-9714 * r - 19070 * g + 28784 * b + 32768) >> 16) + 128;
            int V = ((28784 * r - 24103 * g - 4681 * b + 32768) >> 16) + 128;
            long i = (long)y * w + x;
            int want_u, want_v;
            if (dd->kind == K_Y8) {
                d[0][i] = Y;
                continue;
            }
            if (dd->kind == K_PACKED) {
                /* :148-153, :170-172 -- the sampled chroma lands in the pixel's own 2-byte cell */
                int even = !(x & 1);
                d[0][i * 2 + dd->yo] = Y;
                if (destfmt == F_YVYU)
                    d[0][i * 2 + 1] = even ? V : U;
                else
                    d[0][i * 2 + (1 - dd->yo)] = even ? U : V;
                continue;
            }
            d[0][i] = Y;
            switch (destfmt) {                                          /* :161-168 */
            case F_YUV420P: want_u = !((x | y) & 1); want_v = (x & y) & 1;    break;
            case F_YUV411P: want_u = !(x & 3);       want_v = !((x ^ 2) & 3); break;
            case F_YUV422P: want_u = !(x & 1);       want_v = x & 1;          break;
            default:        want_u = 1;              want_v = 1;              break;
            }
            {
                long ci = (long)(y >> dd->sy) * (w >> dd->sx) + (x >> dd->sx);
                if (want_u) d[1][ci] = U;
                if (want_v) d[2][ci] = V;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Gray / luma-only maps (img_yuv_rgb.c:254-379, img_rgb_packed.c:179-340)                      */

static const uint8_t *luma_of(uint8_t **s, const fmtdesc *sd, int *step)
{
    if (sd->kind == K_PACKED) { *step = 2; return s[0] + sd->yo; }
    *step = 1;
    return s[0];
}

static void fill_chroma(uint8_t **d, const fmtdesc *dd, int w, int h)
{
    long n = chroma_plane_bytes(dd, w, h);
    memset(d[1], 128, n);
    memset(d[2], 128, n);
}

/* ------------------------------------------------------------------------------------------ */
/* Planar <-> planar chroma resampling (img_yuv_planar.c:66-266).  `s`/`d` are ONE chroma plane. */

static void resample_plane(const uint8_t *s, int sfmt, uint8_t *d, int dfmt, int w, int h)
{
    int x, y;
    const int w2 = w / 2, w4 = w / 4;
#define PAIR(a, b) ((a) * 16 + (b))
    enum { P420, P411, P422, P444 };
    int a = sfmt == F_YUV420P ? P420 : sfmt == F_YUV411P ? P411 : sfmt == F_YUV422P ? P422 : P444;
    int b = dfmt == F_YUV420P ? P420 : dfmt == F_YUV411P ? P411 : dfmt == F_YUV422P ? P422 : P444;
    switch (PAIR(a, b)) {
    case PAIR(P420, P411):                                               /* :66-81 */
        for (y = 0; y < (h & ~1); y += 2) {
            for (x = 0; x < (w2 & ~1); x += 2)
                d[y * w4 + x / 2] = (s[(y / 2) * w2 + x] + s[(y / 2) * w2 + x + 1] + 1) / 2;
            memmove(d + (y + 1) * w4, d + y * w4, w4);
        }
        break;
    case PAIR(P420, P422):                                               /* :83-94 */
        for (y = 0; y < (h & ~1); y += 2) {
            memmove(d + y * w2, s + (y / 2) * w2, w2);
            memmove(d + (y + 1) * w2, s + (y / 2) * w2, w2);
        }
        break;
    case PAIR(P420, P444):                                               /* :96-111 */
        for (y = 0; y < h; y += 2) {
            for (x = 0; x < w; x += 2)
                d[y * w + x] = d[y * w + x + 1] = s[(y / 2) * w2 + x / 2];
            memmove(d + (y + 1) * w, d + y * w, w);
        }
        break;
    case PAIR(P411, P420):                                               /* :115-131 */
        for (y = 0; y < (h & ~1); y += 2)
            for (x = 0; x < (w2 & ~1); x += 2)
                d[(y / 2) * w2 + x] = d[(y / 2) * w2 + x + 1] =
                    (s[y * w4 + x / 2] + s[(y + 1) * w4 + x / 2] + 1) / 2;
        break;
    case PAIR(P411, P422):                                               /* :133-146 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w2 & ~1); x += 2)
                d[y * w2 + x] = d[y * w2 + x + 1] = s[y * w4 + x / 2];
        break;
    case PAIR(P411, P444):                                               /* :148-164 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~3); x += 4)
                d[y * w + x] = d[y * w + x + 1] = d[y * w + x + 2] = d[y * w + x + 3] = s[y * w4 + x / 4];
        break;
    case PAIR(P422, P420):                                               /* :168-181 */
        for (y = 0; y < (h & ~1); y += 2)
            for (x = 0; x < w2; x++)
                d[(y / 2) * w2 + x] = (s[y * w2 + x] + s[(y + 1) * w2 + x] + 1) / 2;
        break;
    case PAIR(P422, P411):                                               /* :183-196 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w2 & ~1); x += 2)
                d[y * w4 + x / 2] = (s[y * w2 + x] + s[y * w2 + x + 1] + 1) / 2;
        break;
    case PAIR(P422, P444):                                               /* :198-211 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~1); x += 2)
                d[y * w + x] = d[y * w + x + 1] = s[y * w2 + x / 2];
        break;
    case PAIR(P444, P420):                                               /* :215-232 */
        for (y = 0; y < (h & ~1); y += 2)
            for (x = 0; x < (w & ~1); x += 2)
                d[(y / 2) * w2 + x / 2] = (s[y * w + x] + s[y * w + x + 1]
                                         + s[(y + 1) * w + x] + s[(y + 1) * w + x + 1] + 2) / 4;
        break;
    case PAIR(P444, P411):                                               /* :234-251 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~3); x += 4)
                d[y * w4 + x / 4] = (s[y * w + x] + s[y * w + x + 1]
                                   + s[y * w + x + 2] + s[y * w + x + 3] + 2) / 4;
        break;
    case PAIR(P444, P422):                                               /* :253-266 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~1); x += 2)
                d[y * w2 + x / 2] = (s[y * w + x] + s[y * w + x + 1] + 1) / 2;
        break;
    default:                                                             /* same format: :26-55 */
        memmove(d, s, (size_t)(a == P420 ? w2 * (h / 2) : a == P411 ? w4 * h : a == P422 ? w2 * h : w * h));
        break;
    }
#undef PAIR
}

/* ------------------------------------------------------------------------------------------ */
/* Packed <-> packed permutes (img_yuv_packed.c:23-78); all are safe for s == d.                 */

static void packed_to_packed(const uint8_t *s, int sfmt, uint8_t *d, int dfmt, int w, int h)
{
    long i, cells = (long)w * h;         /* 2-byte cells */
    long groups = (long)w * h / 2;       /* 4-byte groups */
    if (sfmt == dfmt) {                                                  /* :23-27 */
        memmove(d, s, cells * 2);
    } else if ((sfmt == F_YUY2 && dfmt == F_UYVY) || (sfmt == F_UYVY && dfmt == F_YUY2)) {
        for (i = 0; i < cells; i++) {                                    /* :30-38 */
            uint8_t lo = s[i * 2], hi = s[i * 2 + 1];
            d[i * 2] = hi; d[i * 2 + 1] = lo;
        }
    } else if ((sfmt == F_YUY2 && dfmt == F_YVYU) || (sfmt == F_YVYU && dfmt == F_YUY2)) {
        for (i = 0; i < groups; i++) {                                   /* :41-52 */
            uint8_t b0 = s[i * 4], b1 = s[i * 4 + 1], b2 = s[i * 4 + 2], b3 = s[i * 4 + 3];
            d[i * 4] = b0; d[i * 4 + 1] = b3; d[i * 4 + 2] = b2; d[i * 4 + 3] = b1;
        }
    } else if (sfmt == F_UYVY) {          /* -> YVYU */                  /* :56-66 */
        for (i = 0; i < groups; i++) {
            uint8_t b0 = s[i * 4], b1 = s[i * 4 + 1], b2 = s[i * 4 + 2], b3 = s[i * 4 + 3];
            d[i * 4] = b1; d[i * 4 + 1] = b2; d[i * 4 + 2] = b3; d[i * 4 + 3] = b0;
        }
    } else {                              /* YVYU -> UYVY */             /* :68-78 */
        for (i = 0; i < groups; i++) {
            uint8_t b0 = s[i * 4], b1 = s[i * 4 + 1], b2 = s[i * 4 + 2], b3 = s[i * 4 + 3];
            d[i * 4] = b3; d[i * 4 + 1] = b0; d[i * 4 + 2] = b1; d[i * 4 + 3] = b2;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Planar <-> YUY2 (img_yuv_mixed.c:88-209)                                                      */

static void planar_to_yuy2(uint8_t **s, int sfmt, uint8_t *d, int w, int h)
{
    int x, y;
    long i;
    switch (sfmt) {
    case F_YUV420P:                                                      /* :88-101 */
        for (y = 0; y < (h & ~1); y++)
            for (x = 0; x < (w & ~1); x += 2) {
                uint8_t *o = d + ((long)y * w + x) * 2;
                o[0] = s[0][y * w + x];
                o[1] = s[1][(y / 2) * (w / 2) + x / 2];
                o[2] = s[0][y * w + x + 1];
                o[3] = s[2][(y / 2) * (w / 2) + x / 2];
            }
        break;
    case F_YUV411P:                                                      /* :103-116 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~1); x += 2) {
                uint8_t *o = d + ((long)y * w + x) * 2;
                o[0] = s[0][y * w + x];
                o[1] = s[1][y * (w / 4) + x / 4];
                o[2] = s[0][y * w + x + 1];
                o[3] = s[2][y * (w / 4) + x / 4];
            }
        break;
    case F_YUV422P:                                                      /* :118-128 */
        for (i = 0; i < (long)(w / 2) * h; i++) {
            d[i * 4] = s[0][i * 2];
            d[i * 4 + 1] = s[1][i];
            d[i * 4 + 2] = s[0][i * 2 + 1];
            d[i * 4 + 3] = s[2][i];
        }
        break;
    default: /* 444P: chroma pair averaged with TRUNCATION */            /* :130-140 */
        for (i = 0; i < (long)(w / 2) * h; i++) {
            d[i * 4] = s[0][i * 2];
            d[i * 4 + 1] = (s[1][i * 2] + s[1][i * 2 + 1]) / 2;
            d[i * 4 + 2] = s[0][i * 2 + 1];
            d[i * 4 + 3] = (s[2][i * 2] + s[2][i * 2 + 1]) / 2;
        }
        break;
    }
}

static void yuy2_to_planar(const uint8_t *s, uint8_t **d, int dfmt, int w, int h)
{
    int x, y;
    long i;
    switch (dfmt) {
    case F_YUV420P:                                                      /* :144-164 */
        for (y = 0; y < (h & ~1); y++)
            for (x = 0; x < (w & ~1); x += 2) {
                const uint8_t *in = s + ((long)y * w + x) * 2;
                long ci = (long)(y / 2) * (w / 2) + x / 2;
                d[0][y * w + x] = in[0];
                d[0][y * w + x + 1] = in[2];
                if (y % 2 == 0) {
                    d[1][ci] = in[1];
                    d[2][ci] = in[3];
                } else {
                    d[1][ci] = (d[1][ci] + in[1] + 1) / 2;
                    d[2][ci] = (d[2][ci] + in[3] + 1) / 2;
                }
            }
        break;
    case F_YUV411P:                                                      /* :166-182 */
        for (y = 0; y < h; y++)
            for (x = 0; x < (w & ~3); x += 4) {
                const uint8_t *in = s + ((long)y * w + x) * 2;
                d[0][y * w + x] = in[0];
                d[0][y * w + x + 1] = in[2];
                d[0][y * w + x + 2] = in[4];
                d[0][y * w + x + 3] = in[6];
                d[1][y * (w / 4) + x / 4] = (in[1] + in[5] + 1) / 2;
                d[2][y * (w / 4) + x / 4] = (in[3] + in[7] + 1) / 2;
            }
        break;
    case F_YUV422P:                                                      /* :184-194 */
        for (i = 0; i < (long)(w / 2) * h; i++) {
            d[0][i * 2] = s[i * 4];
            d[1][i] = s[i * 4 + 1];
            d[0][i * 2 + 1] = s[i * 4 + 2];
            d[2][i] = s[i * 4 + 3];
        }
        break;
    default: /* 444P */                                                  /* :196-209 */
        for (i = 0; i < (long)(w & ~1) * h; i += 2) {
            d[0][i] = s[i * 2];
            d[1][i] = d[1][i + 1] = s[i * 2 + 1];
            d[0][i + 1] = s[i * 2 + 2];
            d[2][i] = d[2][i + 1] = s[i * 2 + 3];
        }
        break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* The dispatcher.  Mirrors which C function the reference registers for each pair               */
/* (imgconvert.c:34-64 and the five registration blocks).                  */

static int convert(uint8_t **src, int sfmt, uint8_t **dst, int dfmt, int w, int h)
{
    const fmtdesc sd = describe(sfmt), dd = describe(dfmt);
    long i, n = (long)w * h;
    int step;

    if (sd.kind == K_NONE || dd.kind == K_NONE)
        return 0;
    build_tables();

    /* ---- destination is an RGB layout ---- */
    if (dd.kind == K_RGB) {
        if (sd.kind == K_PLANAR || sd.kind == K_PACKED) {
            yuv_to_rgb(src, &sd, dst, &dd, w, h);
        } else if (sd.kind == K_Y8) {                 /* img_yuv_rgb.c:354-379: alpha untouched */
            for (i = 0; i < n; i++) {
                uint8_t v = y2gray[src[0][i]], *px = dst[0] + i * dd.bpp;
                px[dd.ro] = px[dd.go] = px[dd.bo] = v;
            }
        } else if (sd.kind == K_GRAY) {               /* img_rgb_packed.c:307-340: alpha = 0 */
            for (i = 0; i < n; i++) {
                uint8_t v = src[0][i], *px = dst[0] + i * dd.bpp;
                px[dd.ro] = px[dd.go] = px[dd.bo] = v;
                if (dd.ao >= 0) px[dd.ao] = 0;
            }
        } else {                                      /* RGB -> RGB, img_rgb_packed.c:24-177,207-277 */
            if (sfmt == dfmt) {
                memmove(dst[0], src[0], n * sd.bpp);
            } else {
                for (i = 0; i < n; i++) {
                    const uint8_t *in = src[0] + i * sd.bpp;
                    uint8_t *px = dst[0] + i * dd.bpp;
                    uint8_t r = in[sd.ro], g = in[sd.go], b = in[sd.bo];
                    uint8_t a = sd.ao >= 0 ? in[sd.ao] : 0;
                    px[dd.ro] = r; px[dd.go] = g; px[dd.bo] = b;
                    if (dd.ao >= 0) px[dd.ao] = a;
                }
            }
        }
        return 1;
    }

    /* ---- destination is GRAY8 ---- */
    if (dd.kind == K_GRAY) {
        if (sd.kind == K_GRAY) {
            memmove(dst[0], src[0], n);
        } else if (sd.kind == K_RGB) {                /* img_rgb_packed.c:179-303 */
            for (i = 0; i < n; i++) {
                const uint8_t *in = src[0] + i * sd.bpp;
                dst[0][i] = (19595 * in[sd.ro] + 38470 * in[sd.go] + 7471 * in[sd.bo] + 32768) >> 16;
            }
        } else {                                      /* img_yuv_rgb.c:254-279 (Y8->GRAY8 is a range map) */
            const uint8_t *y = luma_of(src, &sd, &step);
            for (i = 0; i < n; i++)
                dst[0][i] = y2gray[y[i * step]];
        }
        return 1;
    }

    /* ---- destination is a YUV layout, source is RGB / GRAY8 ---- */
    if (sd.kind == K_RGB) {
        rgb_to_yuv(src, &sd, dst, dfmt, &dd, w, h);
        return 1;
    }
    if (sd.kind == K_GRAY) {                          /* img_yuv_rgb.c:283-348 */
        if (dd.kind == K_PACKED) {                    /* YVYU uses the YUY2 routine */
            for (i = 0; i < n; i++) {
                dst[0][i * 2 + dd.yo] = gray2y[src[0][i]];
                dst[0][i * 2 + (1 - dd.yo)] = 128;
            }
        } else {
            for (i = 0; i < n; i++)
                dst[0][i] = gray2y[src[0][i]];
            if (dd.kind == K_PLANAR)
                fill_chroma(dst, &dd, w, h);
        }
        return 1;
    }

    /* ---- YUV -> YUV ---- */
    if (dd.kind == K_Y8) {                            /* img_yuv_planar.c:272-276, img_yuv_mixed.c:232-246 */
        const uint8_t *y = luma_of(src, &sd, &step);
        if (step == 1) {
            memmove(dst[0], y, n);
        } else {
            for (i = 0; i < n; i++)
                dst[0][i] = y[i * 2];
        }
        return 1;
    }
    if (sd.kind == K_Y8) {
        if (dd.kind == K_PLANAR) {                    /* img_yuv_planar.c:278-308 */
            memmove(dst[0], src[0], n);
            fill_chroma(dst, &dd, w, h);
        } else {                                      /* img_yuv_mixed.c:212-230; YVYU == YUY2 */
            for (i = 0; i < n; i++) {
                dst[0][i * 2 + dd.yo] = src[0][i];
                dst[0][i * 2 + (1 - dd.yo)] = 128;
            }
        }
        return 1;
    }
    if (sd.kind == K_PLANAR && dd.kind == K_PLANAR) {
        memmove(dst[0], src[0], n);                   /* Y plane is always a straight copy */
        resample_plane(src[1], sfmt, dst[1], dfmt, w, h);
        resample_plane(src[2], sfmt, dst[2], dfmt, w, h);
        return 1;
    }
    if (sd.kind == K_PACKED && dd.kind == K_PACKED) {
        packed_to_packed(src[0], sfmt, dst[0], dfmt, w, h);
        return 1;
    }
    if (sd.kind == K_PLANAR) {                        /* -> packed; img_yuv_mixed.c:26-60 */
        planar_to_yuy2(src, sfmt, dst[0], w, h);
        if (dfmt != F_YUY2)                           /* second hop runs in place on dest */
            packed_to_packed(dst[0], F_YUY2, dst[0], dfmt, w, h);
        return 1;
    }
    /* packed -> planar; img_yuv_mixed.c:26-36,62-84.  The reference rewrites SRC to YUY2 in place. */
    if (sfmt != F_YUY2)
        packed_to_packed(src[0], sfmt, src[0], F_YUY2, w, h);
    yuy2_to_planar(src[0], dst, dfmt, w, h);
    return 1;
}

int oracle_imgconvert(uint8_t **src, int srcfmt, uint8_t **dest, int destfmt, int width, int height)
{
    uint8_t *s3[3], *d3[3];
    if (srcfmt == F_YV12) {                           /* imgconvert.c:40-56 */
        s3[0] = src[0]; s3[1] = src[2]; s3[2] = src[1];
        src = s3; srcfmt = F_YUV420P;
    }
    if (destfmt == F_YV12) {
        d3[0] = dest[0]; d3[1] = dest[2]; d3[2] = dest[1];
        dest = d3; destfmt = F_YUV420P;
    }
    return convert(src, srcfmt, dest, destfmt, width, height);
}

/* ------------------------------------------------------------------------------------------ */

void oracle_average(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes)
{
    int i;
    for (i = 0; i < bytes; i++)
        dest[i] = (src1[i] + src2[i] + 1) / 2;        /* average.c:37-38 */
}

void oracle_rescale(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes,
                    uint32_t weight1, uint32_t weight2)
{
    int i;
    if (weight1 >= 0x10000) {                         /* rescale.c:26-29: src2 is never read */
        memmove(dest, src1, bytes);
    } else if (weight2 >= 0x10000) {
        memmove(dest, src2, bytes);
    } else {
        for (i = 0; i < bytes; i++)                   /* rescale.c:44-45: uint32 math, low byte kept */
            dest[i] = (uint8_t)((src1[i] * weight1 + src2[i] * weight2 + 32768) >> 16);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* libtcvideo shapes that sit directly on ac_average / ac_rescale.         */

void oracle_resize_table(int oldsize, int newsize, int32_t *source, uint32_t *weight1, uint32_t *weight2)
{
    /* tcvideo.c:1138-1165; PI there is M_PI */
    const double ratio = (double)oldsize / (double)newsize;
    int i;
    for (i = 0; i < newsize / 8; i++) {
        double pos = (double)i * (double)oldsize / (double)newsize;
        int s = (int)pos;
        source[i] = s;
        if (pos + ratio < s + 1) {
            weight1[i] = 65536;
            weight2[i] = 0;
        } else {
            double t = ((s + 1) - pos) / ratio * M_PI / 2;
            weight1[i] = (uint32_t)(sin(t) * sin(t) * 65536 + 0.5);
            weight2[i] = 65536 - weight1[i];
        }
    }
}

static void interpolate_odd_rows(const uint8_t *src, uint8_t *dest, int Bpl, int height)
{
    int y;
    for (y = 0; y < height; y++) {                    /* tcvideo.c:353-364 */
        if (y % 2 == 0)
            memmove(dest + (long)y * Bpl, src + (long)y * Bpl, Bpl);
        else if (y == height - 1)
            memmove(dest + (long)y * Bpl, src + (long)(y - 1) * Bpl, Bpl);
        else
            oracle_average(src + (long)(y - 1) * Bpl, src + (long)(y + 1) * Bpl, dest + (long)y * Bpl, Bpl);
    }
}

int oracle_deinterlace(uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int mode)
{
    const int Bpl = width * Bpp;
    int y;
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3))
        return 0;
    if (mode == 2 || mode == 3) {                     /* tcvideo.c:333-345: keep every second row (mode 2 drops the top field) */
        const uint8_t *first = src + (mode == 2 ? Bpl : 0);
        for (y = 0; y < height / 2; y++)
            memmove(dest + (long)y * Bpl, first + (long)(2 * y) * Bpl, Bpl);
        return 1;
    }
    if (mode != 0 && mode != 1)
        return 0;
    interpolate_odd_rows(src, dest, Bpl, height);
    if (mode == 0)
        return 1;
    /* linear blend, tcvideo.c:368-389: even rows are re-interpolated IN src, then whole-frame mean */
    memmove(src, src + Bpl, Bpl);
    for (y = 2; y < height - 1; y += 2)
        oracle_average(src + (long)(y - 1) * Bpl, src + (long)(y + 1) * Bpl, src + (long)y * Bpl, Bpl);
    if (y < height)
        memmove(src + (long)y * Bpl, src + (long)(y - 1) * Bpl, Bpl);
    oracle_average(src, dest, dest, height * Bpl);
    return 1;
}

int oracle_resize(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                  int resize_w, int resize_h, int scale_w, int scale_h)
{
    int new_w, new_h, i, x, y, k;
    int32_t *tsrc;
    uint32_t *tw1, *tw2;
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3))
        return 0;
    if ((scale_w != 1 && scale_w != 2 && scale_w != 4 && scale_w != 8)
     || (scale_h != 1 && scale_h != 2 && scale_h != 4 && scale_h != 8))
        return 0;
    if (width % scale_w != 0 || height % scale_h != 0)
        return 0;
    new_w = width + resize_w * scale_w;
    new_h = height + resize_h * scale_h;
    if (new_w <= 0 || new_h <= 0)
        return 0;
    {
        int m = new_w > new_h ? new_w : new_h;
        tsrc = malloc(sizeof(*tsrc) * (m + 8));
        tw1 = malloc(sizeof(*tw1) * (m + 8));
        tw2 = malloc(sizeof(*tw2) * (m + 8));
    }
    if (resize_h) {                                   /* tcvideo.c:459-476 */
        const int Bpl = width * Bpp;
        oracle_resize_table(height * 8 / scale_h, new_h * 8 / scale_h, tsrc, tw1, tw2);
        for (i = 0; i < scale_h; i++) {
            const uint8_t *sp = src + (long)(i * (height / scale_h)) * Bpl;
            uint8_t *dp = dest + (long)(i * (new_h / scale_h)) * Bpl;
            for (y = 0; y < new_h / scale_h; y++)
                oracle_rescale(sp + (long)tsrc[y] * Bpl, sp + (long)(tsrc[y] + 1) * Bpl,
                               dp + (long)y * Bpl, Bpl, tw1[y], tw2[y]);
        }
    }
    if (resize_w) {                                   /* tcvideo.c:481-531 */
        oracle_resize_table(width * 8 / scale_w, new_w * 8 / scale_w, tsrc, tw1, tw2);
        for (i = 0; i < new_h * scale_w; i++) {
            const uint8_t *sp = src + (long)(i * (width / scale_w)) * Bpp;
            uint8_t *dp = dest + (long)(i * (new_w / scale_w)) * Bpp;
            for (x = 0; x < new_w / scale_w; x++) {
                const uint8_t *a = sp + (long)tsrc[x] * Bpp, *b = sp + (long)(tsrc[x] + 1) * Bpp;
                for (k = 0; k < Bpp; k++) {
                    if (tw1[x] < 0x10000)
                        dp[x * Bpp + k] = (uint8_t)((a[k] * tw1[x] + b[k] * tw2[x] + 32768) >> 16);
                    else
                        dp[x * Bpp + k] = a[k];
                }
            }
        }
    }
    free(tsrc); free(tw1); free(tw2);
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Remaining element-wise libtcvideo operations.  One plane of width x     */
/* height pixels, Bpp (1 or 3) bytes each, tightly packed.                                       */

static int plane_args_ok(const uint8_t *src, const uint8_t *dest, int width, int height, int Bpp)
{
    return src && dest && width > 0 && height > 0 && (Bpp == 1 || Bpp == 3);
}

/* tcvideo.c:184-250.  Negative clip values grow the frame with `black`; values larger than the frame
 * are folded into the opposite edge first (:204-219). */
int oracle_clip(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp,
                int left, int right, int top, int bottom, uint8_t black)
{
    int out_w, out_h, x0, y0, x, y, k;
    if (!plane_args_ok(src, dest, width, height, Bpp))
        return 0;
    if (left + right >= width || top + bottom >= height)
        return 0;
    if (left > width)    { right += left - width;    left = width; }
    if (right > width)   { left += right - width;    right = width; }
    if (top > height)    { bottom += top - height;   top = height; }
    if (bottom > height) { top += bottom - height;   bottom = height; }
    out_w = width - left - right;
    out_h = height - top - bottom;
    /* Output pixel (x, y) shows source pixel (x + left, y + top) when that lies inside the source, else black.
     * (The reference walks rows with memset/memcpy; the picture is the same.) */
    x0 = left;
    y0 = top;
    for (y = 0; y < out_h; y++) {
        const int sy = y + y0;
        for (x = 0; x < out_w; x++) {
            const int sx = x + x0;
            uint8_t *d = dest + ((long)y * out_w + x) * Bpp;
            if (sx >= 0 && sx < width && sy >= 0 && sy < height) {
                for (k = 0; k < Bpp; k++)
                    d[k] = src[((long)sy * width + sx) * Bpp + k];
            } else {
                for (k = 0; k < Bpp; k++)
                    d[k] = black;
            }
        }
    }
    return 1;
}

/* tcvideo.c:681-717: keep every reduce_w-th pixel of every reduce_h-th row. */
int oracle_reduce(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h)
{
    int x, y, k;
    const int out_w = reduce_w > 0 ? width / reduce_w : 0, out_h = reduce_h > 0 ? height / reduce_h : 0;
    if (!plane_args_ok(src, dest, width, height, Bpp) || reduce_w <= 0 || reduce_h <= 0)
        return 0;
    if (reduce_w == 1 && reduce_h == 1) {             /* :713-715 whole-frame copy (keeps the full width*height) */
        memmove(dest, src, (size_t)width * height * Bpp);
        return 1;
    }
    if (reduce_w == 1) {                              /* :706-711 rows keep their full width */
        for (y = 0; y < out_h; y++)
            memmove(dest + (long)y * width * Bpp, src + (long)y * reduce_h * width * Bpp, (size_t)width * Bpp);
        return 1;
    }
    for (y = 0; y < out_h; y++)                        /* :694-704 */
        for (x = 0; x < out_w; x++)
            for (k = 0; k < Bpp; k++)
                dest[((long)y * out_w + x) * Bpp + k] = src[((long)y * reduce_h * width + (long)x * reduce_w) * Bpp + k];
    return 1;
}

/* tcvideo.c:739-766 (rows mirrored top-bottom) and :787-818 (pixels mirrored left-right); src == dest is legal. */
int oracle_flip_v(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp)
{
    const long Bpl = (long)width * Bpp;
    long i;
    int y;
    if (!plane_args_ok(src, dest, width, height, Bpp))
        return 0;
    for (y = 0; y < (height + 1) / 2; y++) {
        const uint8_t *a = src + y * Bpl, *b = src + (long)(height - 1 - y) * Bpl;
        uint8_t *da = dest + y * Bpl, *db = dest + (long)(height - 1 - y) * Bpl;
        for (i = 0; i < Bpl; i++) {
            const uint8_t t = a[i];
            da[i] = b[i];
            db[i] = t;
        }
    }
    return 1;
}

int oracle_flip_h(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp)
{
    int x, y, k;
    if (!plane_args_ok(src, dest, width, height, Bpp))
        return 0;
    for (y = 0; y < height; y++) {
        const uint8_t *s = src + (long)y * width * Bpp;
        uint8_t *d = dest + (long)y * width * Bpp;
        for (x = 0; x < (width + 1) / 2; x++) {
            const int m = width - 1 - x;
            for (k = 0; k < Bpp; k++) {
                const uint8_t t = s[x * Bpp + k];
                d[x * Bpp + k] = s[m * Bpp + k];
                d[m * Bpp + k] = t;
            }
        }
    }
    return 1;
}

/* tcvideo.c:1180-1189: table[i] = (uint8_t)(pow(i/255, gamma) * 255), truncated. */
void oracle_gamma_table(double gamma, uint8_t *table)
{
    int i;
    for (i = 0; i < 256; i++)
        table[i] = (uint8_t)(pow(i / 255.0, gamma) * 255);
}

int oracle_gamma_correct(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma)
{
    uint8_t table[256];
    long i, n = (long)width * height * Bpp;
    if (!plane_args_ok(src, dest, width, height, Bpp) || gamma <= 0)
        return 0;
    oracle_gamma_table(gamma, table);
    for (i = 0; i < n; i++)                            /* tcvideo.c:853-855 */
        dest[i] = table[src[i]];
    return 1;
}

/* tcvideo.c:1209-1224: 16.16 weights of the centre (c), left/right (x), up/down (y) and diagonal (d) taps;
 * tables[0..255] = c, [256..511] = x, [512..767] = y, [768..1023] = d.  double -> uint32 truncates. */
void oracle_aa_tables(double weight, double bias, uint32_t *tables)
{
    int i;
    for (i = 0; i < 256; i++) {
        tables[i] = i * weight * 65536;
        tables[256 + i] = i * bias * (1 - weight) / 4 * 65536;
        tables[512 + i] = i * (1 - bias) * (1 - weight) / 4 * 65536;
        tables[768 + i] = (tables[256 + i] + tables[512 + i] + 1) / 2;
    }
}

/* "same colour": the largest per-channel difference is below 25 (tcvideo.c:37,917-927) */
static int aa_same(const uint8_t *p, const uint8_t *q, int Bpp)
{
    int k, worst = 0;
    for (k = 0; k < Bpp; k++) {
        const int d = abs((int)q[k] - (int)p[k]);
        if (d > worst)
            worst = d;
    }
    return worst < 25;
}

/* tcvideo.c:886-980.  Border pixels are copied.  An interior pixel is smoothed when its left (or right) neighbour
 * matches exactly one of the vertical neighbours and differs from the other vertical one and from the opposite
 * horizontal one (:945-949); the smoothed value is the 3x3 table-weighted sum + 32768, >> 16, in uint32. */
int oracle_antialias(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias)
{
    uint32_t t[1024];
    const long Bpl = (long)width * Bpp;
    int x, y, k;
    if (!plane_args_ok(src, dest, width, height, Bpp) || weight < 0 || weight > 1 || bias < 0 || bias > 1)
        return 0;
    oracle_aa_tables(weight, bias, t);
    for (y = 0; y < height; y++) {
        const uint8_t *row = src + y * Bpl;
        uint8_t *out = dest + y * Bpl;
        if (y == 0 || y == height - 1) {
            memmove(out, row, (size_t)Bpl);
            /* the reference copies the first and then the last row even when they are the same row (height 1), and a
             * height-2 frame has no interior rows */
            continue;
        }
        for (x = 0; x < width; x++) {
            const uint8_t *c = row + (long)x * Bpp;
            int smooth = 0;
            if (x > 0 && x < width - 1) {
                const uint8_t *l = c - Bpp, *r = c + Bpp, *u = c - Bpl, *d = c + Bpl;
                const int lr = aa_same(l, r, Bpp);
                smooth = (aa_same(l, u, Bpp) && !aa_same(l, d, Bpp) && !lr)
                      || (aa_same(l, d, Bpp) && !aa_same(l, u, Bpp) && !lr)
                      || (aa_same(r, u, Bpp) && !aa_same(r, d, Bpp) && !lr)
                      || (aa_same(r, d, Bpp) && !aa_same(r, u, Bpp) && !lr);
            }
            for (k = 0; k < Bpp; k++) {
                if (smooth) {
                    const uint8_t *q = c + k;
                    const uint32_t sum = t[768 + q[-Bpl - Bpp]] + t[512 + q[-Bpl]] + t[768 + q[-Bpl + Bpp]]
                                       + t[256 + q[-Bpp]]       + t[q[0]]          + t[256 + q[Bpp]]
                                       + t[768 + q[Bpl - Bpp]]  + t[512 + q[Bpl]]  + t[768 + q[Bpl + Bpp]]
                                       + 32768;
                    out[(long)x * Bpp + k] = (uint8_t)(sum >> 16);
                } else {
                    out[(long)x * Bpp + k] = c[k];
                }
            }
        }
    }
    return 1;
}
