/*
 * cpubench.c -- frame-parallel CPU timing harness for the checkers (TEST/BENCH INFRASTRUCTURE).
 *
 * dlopen()s one of
 *     oracle/_ref/libac_ref_c.so        (unmodified reference aclib, plain-C path:   --accel 0)
 *     oracle/_ref/libac_ref_sse2.so     (unmodified reference aclib, SSE2 asm path:  --accel -1)
 *     oracle/_ref/libtcv_ref.so         (unmodified reference libtcvideo over the plain-C aclib)
 *     oracle/_ref/libtcv_ref_sse2.so    (unmodified reference libtcvideo over the SSE2 aclib: what stock transcode runs)
 *     oracle/liboracle.so               (this repo's restatement:                    --oracle)
 * and times it with T pthreads, one frame per thread at a time, the way transcode's frame threads use the libraries
 * (src/frame_threads.c:174-228,316).  Prints one JSON object on stdout.  Used by bench.py's cpu_baseline leg and
 * `--impl reference`.
 *
 * usage: cpubench LIB [--oracle] [--accel N] [--threads T] [--seconds S] [--frames F] -w W -h H
 *          --op convert --src FMT --dst FMT      one ac_imgconvert per frame
 *          --op average|rescale --bpp 1|3        per-row ac_average / ac_rescale over a plane (the libraries without libtcvideo)
 *          --op chain --src FMT --chain SPEC     a do_process_frame-shaped stage list per frame (needs a libtcv_ref library):
 *              SPEC = stage[,stage...]; stage = convert:FMT | clip:L:R:T:B | deint:MODE | resize:RW:RH | reduce:RW:RH |
 *                     flipv | fliph | gamma:G | antialias:W:B        (MODE is transcode's -I number: 1, 4 or 5)
 *              Stages run as src/video_trans.c:192-426 runs them: PROCESS_FRAME stages once per plane with the plane's divided
 *              size and arguments, -I 1/5, gamma and antialias on the first plane only; conversions through ac_imgconvert.
 * The clock starts when every worker has allocated and filled its buffers (a barrier), not before.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int (*init_fn)(int);
typedef int (*convert_fn)(uint8_t **, int, uint8_t **, int, int, int);
typedef void (*average_fn)(const uint8_t *, const uint8_t *, uint8_t *, int);
typedef void (*rescale_fn)(const uint8_t *, const uint8_t *, uint8_t *, int, uint32_t, uint32_t);
typedef void *(*tcv_init_fn)(void);
typedef int (*tcv_clip_fn)(void *, uint8_t *, uint8_t *, int, int, int, int, int, int, int, uint8_t);
typedef int (*tcv_deint_fn)(void *, uint8_t *, uint8_t *, int, int, int, int);
typedef int (*tcv_resize_fn)(void *, uint8_t *, uint8_t *, int, int, int, int, int, int, int);
typedef int (*tcv_reduce_fn)(void *, uint8_t *, uint8_t *, int, int, int, int, int);
typedef int (*tcv_flip_fn)(void *, uint8_t *, uint8_t *, int, int, int);
typedef int (*tcv_gamma_fn)(void *, uint8_t *, uint8_t *, int, int, int, double);
typedef int (*tcv_aa_fn)(void *, uint8_t *, uint8_t *, int, int, int, double, double);

static convert_fn f_convert;
static average_fn f_average;
static rescale_fn f_rescale;
static tcv_init_fn t_init;
static tcv_clip_fn t_clip;
static tcv_deint_fn t_deint;
static tcv_resize_fn t_resize;
static tcv_reduce_fn t_reduce;
static tcv_flip_fn t_flipv, t_fliph;
static tcv_gamma_fn t_gamma;
static tcv_aa_fn t_aa;

static int srcfmt = 0x1001, dstfmt = 0x2001, W = 1920, H = 1080, bpp = 1;
static const char *op = "convert";
static double seconds = 2.0;
static int nbuf = 4;      /* distinct frames each thread cycles through (streams from memory like a real frame queue) */
static volatile int stop_flag;
static pthread_barrier_t ready;        /* workers + main: buffers are initialised, start the clock */

enum { S_CONVERT, S_CLIP, S_DEINT, S_RESIZE, S_REDUCE, S_FLIPV, S_FLIPH, S_GAMMA, S_AA };
typedef struct { int kind, p[4]; double d[2]; } stage_t;
static stage_t stages[32];
static int nstages;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static long uvsize(int fmt, int w, int h)
{
    switch (fmt) {
    case 0x1001: case 0x1002: return (long)(w / 2) * (h / 2);
    case 0x1003: return (long)(w / 4) * h;
    case 0x1004: return (long)(w / 2) * h;
    case 0x1005: return (long)w * h;
    default: return 0;
    }
}

static long frame_bytes(int fmt, int w, int h)
{
    long p = (long)w * h;
    switch (fmt) {
    case 0x1001: case 0x1002: case 0x1003: case 0x1004: case 0x1005: return p + 2 * uvsize(fmt, w, h);
    case 0x1006: case 0x1007: case 0x1008: return p * 2;
    case 0x1009: case 0x2007: return p;
    case 0x2001: case 0x2002: return p * 3;
    default: return p * 4;
    }
}

static void planes(uint8_t **pl, uint8_t *buf, int fmt, int w, int h)
{
    pl[0] = buf;
    pl[1] = buf + (long)w * h;
    pl[2] = pl[1] + uvsize(fmt, w, h);
}

/* ---- do_process_frame-shaped chains -------------------------------------------------------------------------------- */
static int parse_chain(const char *spec)
{
    char *copy = strdup(spec), *save = NULL, *tok;
    for (tok = strtok_r(copy, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
        stage_t *s = &stages[nstages];
        char name[32] = {0};
        double v[4] = {0, 0, 0, 0};
        int n = 0, i;
        char *colon = strchr(tok, ':');
        if (nstages == 32) return 0;
        snprintf(name, sizeof(name), "%.*s", colon ? (int)(colon - tok) : (int)strlen(tok), tok);
        while (colon && n < 4) { v[n++] = strtod(colon + 1, NULL); colon = strchr(colon + 1, ':'); }
        if (!strcmp(name, "convert")) { s->kind = S_CONVERT; s->p[0] = (int)strtol(strchr(tok, ':') + 1, NULL, 0); }
        else if (!strcmp(name, "clip")) s->kind = S_CLIP;
        else if (!strcmp(name, "deint")) s->kind = S_DEINT;
        else if (!strcmp(name, "resize")) s->kind = S_RESIZE;
        else if (!strcmp(name, "reduce")) s->kind = S_REDUCE;
        else if (!strcmp(name, "flipv")) s->kind = S_FLIPV;
        else if (!strcmp(name, "fliph")) s->kind = S_FLIPH;
        else if (!strcmp(name, "gamma")) { s->kind = S_GAMMA; s->d[0] = v[0]; }
        else if (!strcmp(name, "antialias")) { s->kind = S_AA; s->d[0] = v[0]; s->d[1] = v[1]; }
        else return 0;
        if (s->kind != S_CONVERT) for (i = 0; i < 4; i++) s->p[i] = (int)v[i];
        nstages++;
    }
    free(copy);
    return 1;
}

typedef struct { int n, Bpp, wd[3], hd[3]; uint8_t black[3]; } pset_t;

static int plane_set(int fmt, pset_t *ps)      /* set_vtd, src/video_trans.c:71-118 */
{
    int i;
    ps->n = 1; ps->Bpp = 1;
    for (i = 0; i < 3; i++) { ps->wd[i] = ps->hd[i] = 1; ps->black[i] = 0; }
    if (fmt == 0x1001 || fmt == 0x1004) {
        ps->n = 3;
        ps->wd[1] = ps->wd[2] = 2;
        ps->hd[1] = ps->hd[2] = fmt == 0x1001 ? 2 : 1;
        ps->black[1] = ps->black[2] = 128;
    } else if (fmt == 0x2001) ps->Bpp = 3;
    else if (fmt != 0x1009 && fmt != 0x2007) return 0;
    return 1;
}

static long plane_off(const pset_t *ps, int w, int h, int i)
{
    long off = 0;
    int k;
    for (k = 0; k < i; k++) off += (long)(w / ps->wd[k]) * (h / ps->hd[k]) * ps->Bpp;
    return off;
}

/* Runs the stage list on the frame in `a` (fmt, w, h) with `b` as the second buffer; returns the buffer holding the result. */
static uint8_t *run_chain(void *handle, uint8_t *a, uint8_t *b, int fmt, int w, int h)
{
    int k, i;
    for (k = 0; k < nstages; k++) {
        const stage_t *s = &stages[k];
        pset_t ps;
        uint8_t *t;
        int nw = w, nh = h;
        if (s->kind == S_CONVERT) {
            uint8_t *sp[3], *dp[3];
            if (s->p[0] == fmt) continue;
            planes(sp, a, fmt, w, h);
            planes(dp, b, s->p[0], w, h);
            if (!f_convert(sp, fmt, dp, s->p[0], w, h)) { fprintf(stderr, "cpubench: conversion failed\n"); exit(1); }
            fmt = s->p[0];
            t = a; a = b; b = t;
            continue;
        }
        if (!plane_set(fmt, &ps)) { fprintf(stderr, "cpubench: stage %d cannot run on format 0x%x\n", k, fmt); exit(1); }
        switch (s->kind) {
        case S_CLIP:   nw = w - s->p[0] - s->p[1]; nh = h - s->p[2] - s->p[3]; break;
        case S_DEINT:  if (s->p[0] == 4) nh = h / 2; break;
        case S_REDUCE: nw = w / s->p[0]; nh = h / s->p[1]; break;
        default: break;
        }
        if (s->kind == S_RESIZE) {            /* video_trans.c:281-297: rows first, then columns */
            if (s->p[1]) {
                nh = h + s->p[1] * 8;
                for (i = 0; i < ps.n; i++)
                    t_resize(handle, a + plane_off(&ps, w, h, i), b + plane_off(&ps, w, nh, i), w / ps.wd[i], h / ps.hd[i], ps.Bpp,
                             0, s->p[1], 8 / ps.wd[i], 8 / ps.hd[i]);
                h = nh;
                t = a; a = b; b = t;
            }
            if (s->p[0]) {
                nw = w + s->p[0] * 8;
                for (i = 0; i < ps.n; i++)
                    t_resize(handle, a + plane_off(&ps, w, h, i), b + plane_off(&ps, nw, h, i), w / ps.wd[i], h / ps.hd[i], ps.Bpp,
                             s->p[0], 0, 8 / ps.wd[i], 8 / ps.hd[i]);
                w = nw;
                t = a; a = b; b = t;
            }
            continue;
        }
        if (s->kind == S_GAMMA) {             /* video_trans.c:390-396: first plane, in place */
            t_gamma(handle, a, a, w, h, ps.Bpp, s->d[0]);
            continue;
        }
        for (i = 0; i < ps.n; i++) {
            uint8_t *sp = a + plane_off(&ps, w, h, i), *dp = b + plane_off(&ps, nw, nh, i);
            const int pw = w / ps.wd[i], ph = h / ps.hd[i];
            const long pbytes = (long)pw * ph * ps.Bpp;
            switch (s->kind) {
            case S_CLIP:
                t_clip(handle, sp, dp, pw, ph, ps.Bpp, s->p[0] / ps.wd[i], s->p[1] / ps.wd[i], s->p[2] / ps.hd[i], s->p[3] / ps.hd[i], ps.black[i]);
                break;
            case S_DEINT:                     /* -I 1 / 5: first plane, others copied; -I 4: every plane (video_trans.c:227-277) */
                if (s->p[0] == 4) t_deint(handle, sp, dp, pw, ph, ps.Bpp, 1);
                else if (i == 0) t_deint(handle, sp, dp, pw, ph, ps.Bpp, s->p[0] == 1 ? 2 : 3);
                else memcpy(dp, sp, pbytes);
                break;
            case S_REDUCE: t_reduce(handle, sp, dp, pw, ph, ps.Bpp, s->p[0], s->p[1]); break;
            case S_FLIPV:  t_flipv(handle, sp, dp, pw, ph, ps.Bpp); break;
            case S_FLIPH:  t_fliph(handle, sp, dp, pw, ph, ps.Bpp); break;
            case S_AA:
                if (i == 0) t_aa(handle, sp, dp, pw, ph, ps.Bpp, s->d[0], s->d[1]);
                else memcpy(dp, sp, pbytes);
                break;
            default: break;
            }
        }
        w = nw; h = nh;
        t = a; a = b; b = t;
    }
    return a;
}

static long chain_max_bytes(void)
{
    long best = frame_bytes(srcfmt, W, H);
    int fmt = srcfmt, w = W, h = H, k;
    for (k = 0; k < nstages; k++) {
        const stage_t *s = &stages[k];
        long b;
        switch (s->kind) {
        case S_CONVERT: fmt = s->p[0]; break;
        case S_CLIP:    w -= s->p[0] + s->p[1]; h -= s->p[2] + s->p[3]; break;
        case S_DEINT:   if (s->p[0] == 4) h /= 2; break;
        case S_RESIZE:
            h += s->p[1] * 8;
            b = frame_bytes(fmt, w, h);
            if (b > best) best = b;
            w += s->p[0] * 8;
            break;
        case S_REDUCE:  w /= s->p[0]; h /= s->p[1]; break;
        default: break;
        }
        b = frame_bytes(fmt, w, h);
        if (b > best) best = b;
    }
    return best;
}

typedef struct { long frames; int id; } worker_t;

static void *worker(void *arg)
{
    worker_t *wk = arg;
    uint64_t seed = 0x1234 + wk->id;
    long i;
    if (!strcmp(op, "convert") || !strcmp(op, "chain")) {
        const int chain = !strcmp(op, "chain");
        long sb = frame_bytes(srcfmt, W, H), db = chain ? chain_max_bytes() : frame_bytes(dstfmt, W, H);
        uint8_t *s = malloc(sb + 64), **s2 = malloc(sizeof(*s2) * nbuf), **d = malloc(sizeof(*d) * nbuf);
        void *handle = chain ? t_init() : NULL;
        uint8_t *sp[3], *dp[3];
        int k, cur = 0;
        for (i = 0; i < sb; i++) s[i] = (uint8_t)sm64(&seed);
        for (k = 0; k < nbuf; k++) {
            s2[k] = malloc((chain && db > sb ? db : sb) + 64);
            d[k] = malloc(db + 64);
            memcpy(s2[k], s, sb);
            s2[k][k % sb] ^= (uint8_t)(k + 1);       /* distinct frames */
            memset(d[k], 0x55, db);
        }
        /* UYVY/YVYU sources are rewritten in place by the reference's planar wrapper (aclib/img_yuv_mixed.c:30-32): refresh
         * those every frame */
        int refresh = !chain && (srcfmt == 0x1007 || srcfmt == 0x1008);
        pthread_barrier_wait(&ready);
        while (!stop_flag) {
            if (chain) {
                /* stages ping-pong between the frame's two buffers as do_process_frame does (vframe_list_t.video_buf_Y[0/1],
                 * src/video_trans.c:130-150).  No copy is charged for keeping the source: later stages land in it, so the
                 * next visit converts whatever bytes the last one left -- the same work per frame. */
                run_chain(handle, s2[cur], d[cur], srcfmt, W, H);
            } else {
                if (refresh) memcpy(s2[cur], s, sb);
                planes(sp, s2[cur], srcfmt, W, H);
                planes(dp, d[cur], dstfmt, W, H);
                f_convert(sp, srcfmt, dp, dstfmt, W, H);
            }
            wk->frames++;
            if (++cur == nbuf) cur = 0;
        }
        for (k = 0; k < nbuf; k++) { free(s2[k]); free(d[k]); }
        free(s); free(s2); free(d);
    } else {
        long Bpl = (long)W * bpp, n = Bpl * H;
        uint8_t *s = malloc(n + Bpl + 64), *d = malloc(n + 64);
        int y;
        for (i = 0; i < n + Bpl; i++) s[i] = (uint8_t)sm64(&seed);
        memset(d, 0x55, n);
        pthread_barrier_wait(&ready);
        while (!stop_flag) {
            if (!strcmp(op, "average")) {
                /* deinterlace-interpolate shape, libtcvideo/tcvideo.c:353-364 */
                for (y = 0; y < H; y++) {
                    if (y % 2 == 0 || y == H - 1) memcpy(d + y * Bpl, s + (y & ~1) * Bpl, Bpl);
                    else f_average(s + (y - 1) * Bpl, s + (y + 1) * Bpl, d + y * Bpl, Bpl);
                }
            } else {
                /* one blended row per output row: the 3:2 shrink weights of tcv_resize (tcvideo.c:464-475) */
                for (y = 0; y < H; y++)
                    f_rescale(s + y * Bpl, s + (y + 1) * Bpl, d + y * Bpl, Bpl,
                              (y & 1) ? 16384 : 49152, (y & 1) ? 49152 : 16384);
            }
            wk->frames++;
        }
        free(s); free(d);
    }
    return NULL;
}

int main(int argc, char **argv)
{
    int threads = (int)sysconf(_SC_NPROCESSORS_ONLN), accel = 0, oracle = 0, i;
    const char *lib, *chain_spec = NULL;
    void *h;
    if (argc < 2) { fprintf(stderr, "usage: cpubench LIB [options]\n"); return 2; }
    lib = argv[1];
    for (i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--oracle")) oracle = 1;
        else if (!strcmp(argv[i], "--accel") && i + 1 < argc) accel = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--op") && i + 1 < argc) op = argv[++i];
        else if (!strcmp(argv[i], "--src") && i + 1 < argc) srcfmt = (int)strtol(argv[++i], NULL, 0);
        else if (!strcmp(argv[i], "--dst") && i + 1 < argc) dstfmt = (int)strtol(argv[++i], NULL, 0);
        else if (!strcmp(argv[i], "--chain") && i + 1 < argc) chain_spec = argv[++i];
        else if (!strcmp(argv[i], "-w") && i + 1 < argc) W = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-h") && i + 1 < argc) H = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--bpp") && i + 1 < argc) bpp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = atof(argv[++i]);
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) nbuf = atoi(argv[++i]) > 0 ? atoi(argv[i]) : 1;
        else { fprintf(stderr, "cpubench: bad argument %s\n", argv[i]); return 2; }
    }
    if (threads < 1) threads = 1;
    h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "cpubench: %s\n", dlerror()); return 1; }
    if (oracle) {
        f_convert = (convert_fn)dlsym(h, "oracle_imgconvert");
        f_average = (average_fn)dlsym(h, "oracle_average");
        f_rescale = (rescale_fn)dlsym(h, "oracle_rescale");
    } else {
        init_fn init = (init_fn)dlsym(h, "ac_init");
        f_convert = (convert_fn)dlsym(h, "ac_imgconvert");
        f_average = (average_fn)dlsym(h, "ac_average");
        f_rescale = (rescale_fn)dlsym(h, "ac_rescale");
        if (!init || !init(accel)) { fprintf(stderr, "cpubench: ac_init failed\n"); return 1; }
    }
    if (!f_convert || !f_average || !f_rescale) { fprintf(stderr, "cpubench: missing symbols\n"); return 1; }
    if (!strcmp(op, "chain")) {
        if (!chain_spec || !parse_chain(chain_spec)) { fprintf(stderr, "cpubench: bad --chain\n"); return 2; }
        t_init = (tcv_init_fn)dlsym(h, "tcv_init");
        t_clip = (tcv_clip_fn)dlsym(h, "tcv_clip");
        t_deint = (tcv_deint_fn)dlsym(h, "tcv_deinterlace");
        t_resize = (tcv_resize_fn)dlsym(h, "tcv_resize");
        t_reduce = (tcv_reduce_fn)dlsym(h, "tcv_reduce");
        t_flipv = (tcv_flip_fn)dlsym(h, "tcv_flip_v");
        t_fliph = (tcv_flip_fn)dlsym(h, "tcv_flip_h");
        t_gamma = (tcv_gamma_fn)dlsym(h, "tcv_gamma_correct");
        t_aa = (tcv_aa_fn)dlsym(h, "tcv_antialias");
        if (!t_init || !t_clip || !t_deint || !t_resize || !t_reduce || !t_flipv || !t_fliph || !t_gamma || !t_aa) {
            fprintf(stderr, "cpubench: %s has no libtcvideo (use a libtcv_ref library for --op chain)\n", lib);
            return 1;
        }
    }

    {
        pthread_t *tid = calloc(threads, sizeof(*tid));
        worker_t *wk = calloc(threads, sizeof(*wk));
        double t0, t1;
        long total = 0;
        /* warm-up: builds the lazily-created LUTs single-threaded (img_yuv_rgb.c:42-56 is racy) */
        {
            uint8_t *s = calloc(1, 64 * 16 * 4 + 64), *d = calloc(1, 64 * 16 * 4 + 64), *sp[3], *dp[3];
            planes(sp, s, 0x1001, 64, 16); dp[0] = d;
            f_convert(sp, 0x1001, dp, 0x2001, 64, 16);
            sp[0] = s; planes(dp, d, 0x1001, 64, 16);
            f_convert(sp, 0x2007, dp, 0x1009, 64, 16);
            free(s); free(d);
        }
        pthread_barrier_init(&ready, NULL, (unsigned)threads + 1);
        for (i = 0; i < threads; i++) { wk[i].id = i; pthread_create(&tid[i], NULL, worker, &wk[i]); }
        pthread_barrier_wait(&ready);       /* every worker holds initialised (page-touched) buffers */
        t0 = now();
        while (now() - t0 < seconds) usleep(2000);
        stop_flag = 1;
        for (i = 0; i < threads; i++) pthread_join(tid[i], NULL);
        t1 = now();
        for (i = 0; i < threads; i++) total += wk[i].frames;
        printf("{\"lib\": \"%s\", \"op\": \"%s\", \"chain\": \"%s\", \"srcfmt\": %d, \"dstfmt\": %d, \"width\": %d, \"height\": %d, "
               "\"bpp\": %d, \"accel\": %d, \"threads\": %d, \"frames\": %ld, \"seconds\": %.4f, \"frames_per_s\": %.2f}\n",
               lib, op, chain_spec ? chain_spec : "", srcfmt, dstfmt, W, H, bpp, accel, threads, total, t1 - t0, total / (t1 - t0));
    }
    return 0;
}
