/*
 * cpubench.c -- frame-parallel CPU timing harness for the checkers (TEST/BENCH INFRASTRUCTURE).
 *
 * dlopen()s one of
 *     oracle/_ref/libac_ref_c.so      (unmodified reference, plain-C path:   --accel 0)
 *     oracle/_ref/libac_ref_sse2.so   (unmodified reference, SSE2 asm path:  --accel -1)
 *     oracle/liboracle.so             (this repo's restatement:              --oracle)
 * and times ac_imgconvert / ac_rescale / ac_average with T pthreads, one frame per thread at a time,
 * the way transcode's frame threads use the library (src/frame_threads.c:174-228,316).
 * Prints one JSON object on stdout.  Used by bench.py's cpu_baseline leg and `--impl reference`.
 *
 * usage: cpubench LIB [--oracle] [--accel N] --op convert --src FMT --dst FMT -w W -h H
 *                 [--frames F (distinct frames per thread, default 4)] [--threads T] [--seconds S]
 *        cpubench LIB ... --op average|rescale  (row-blend a W*H*bpp plane; --bpp 1|3)
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

static convert_fn f_convert;
static average_fn f_average;
static rescale_fn f_rescale;

static int srcfmt = 0x1001, dstfmt = 0x2001, W = 1920, H = 1080, bpp = 1;
static const char *op = "convert";
static double seconds = 2.0;
static int nbuf = 4;      /* distinct frames each thread cycles through (streams from memory like a real frame queue) */
static volatile int stop_flag;

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

typedef struct { long frames; int id; } worker_t;

static void *worker(void *arg)
{
    worker_t *wk = arg;
    uint64_t seed = 0x1234 + wk->id;
    long i;
    if (!strcmp(op, "convert")) {
        long sb = frame_bytes(srcfmt, W, H), db = frame_bytes(dstfmt, W, H);
        uint8_t *s = malloc(sb + 64), **s2 = malloc(sizeof(*s2) * nbuf), **d = malloc(sizeof(*d) * nbuf);
        uint8_t *sp[3], *dp[3];
        int k, cur = 0;
        for (i = 0; i < sb; i++) s[i] = (uint8_t)sm64(&seed);
        for (k = 0; k < nbuf; k++) {
            s2[k] = malloc(sb + 64);
            d[k] = malloc(db + 64);
            memcpy(s2[k], s, sb);
            s2[k][k % sb] ^= (uint8_t)(k + 1);       /* distinct frames */
            memset(d[k], 0x55, db);
        }
        /* UYVY/YVYU sources are rewritten in place by the reference: refresh src each frame only then */
        int refresh = (srcfmt == 0x1007 || srcfmt == 0x1008);
        while (!stop_flag) {
            if (refresh) memcpy(s2[cur], s, sb);
            planes(sp, s2[cur], srcfmt, W, H);
            planes(dp, d[cur], dstfmt, W, H);
            f_convert(sp, srcfmt, dp, dstfmt, W, H);
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
    const char *lib;
    void *h;
    if (argc < 2) { fprintf(stderr, "usage: cpubench LIB [options]\n"); return 2; }
    lib = argv[1];
    for (i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "--oracle")) oracle = 1;
        else if (!strcmp(argv[i], "--accel") && i + 1 < argc) accel = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--op") && i + 1 < argc) op = argv[++i];
        else if (!strcmp(argv[i], "--src") && i + 1 < argc) srcfmt = (int)strtol(argv[++i], NULL, 0);
        else if (!strcmp(argv[i], "--dst") && i + 1 < argc) dstfmt = (int)strtol(argv[++i], NULL, 0);
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
        t0 = now();
        for (i = 0; i < threads; i++) { wk[i].id = i; pthread_create(&tid[i], NULL, worker, &wk[i]); }
        while (now() - t0 < seconds) usleep(2000);
        stop_flag = 1;
        for (i = 0; i < threads; i++) pthread_join(tid[i], NULL);
        t1 = now();
        for (i = 0; i < threads; i++) total += wk[i].frames;
        printf("{\"lib\": \"%s\", \"op\": \"%s\", \"srcfmt\": %d, \"dstfmt\": %d, \"width\": %d, \"height\": %d, "
               "\"bpp\": %d, \"accel\": %d, \"threads\": %d, \"frames\": %ld, \"seconds\": %.4f, \"frames_per_s\": %.2f}\n",
               lib, op, srcfmt, dstfmt, W, H, bpp, accel, threads, total, t1 - t0, total / (t1 - t0));
    }
    return 0;
}
