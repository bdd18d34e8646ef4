/*
 * tcv_ref_stubs.c -- the three libtc symbols the reference's libtcvideo needs at link time
 * (TEST INFRASTRUCTURE: linked only into oracle/_ref/libtcv_ref.so, never into the product).
 *
 * libtcvideo/tcvideo.c allocates its handle with tc_zalloc()/tc_malloc() (libtcutil/memutils.h:54-97) and reports
 * bad parameters through tc_log() (libtcutil/logging.h:251).  libtc itself drags in the whole of transcode, so the
 * reference build of libtcvideo links these stand-ins instead: plain malloc/calloc and a stderr printer that can be
 * silenced (tests exercise the error returns on purpose).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

int tcv_ref_stubs_quiet = 1;

void *_tc_malloc(const char *file, int line, size_t size)
{
    (void)file; (void)line;
    return malloc(size);
}

void *_tc_zalloc(const char *file, int line, size_t size)
{
    (void)file; (void)line;
    return calloc(1, size);
}

int tc_log(int type, const char *tag, const char *fmt, ...)
{
    va_list ap;
    (void)type;
    if (tcv_ref_stubs_quiet) return 0;
    fprintf(stderr, "[%s] ", tag ? tag : "?");
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    return 0;
}
