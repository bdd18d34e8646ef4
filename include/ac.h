/*
 * include/ac.h -- libacgpu's copy of the aclib core interface.
 *
 * Drop-in for the reference header aclib/ac.h:24-91: identical names, argument meaning, return
 * conventions (1 = success, 0 = failure) and flag values, so transcode's callers compile and link
 * unchanged.  What is new is one acceleration bit, AC_CUDA, and the fact that it is the ONLY
 * implementation this library contains: there is no C or SIMD fallback behind these symbols.
 *
 *   reference symbol (file:line)                    libacgpu behaviour
 *   ac_init            aclib/accore.c:29-40         selects the CUDA path; fails (0) when the masked
 *                                                   accel lacks AC_CUDA or no sm_100 device is usable
 *   ac_cpuinfo         aclib/accore.c:46-53         AC_CUDA when a usable device exists, else 0
 *   ac_endian          aclib/accore.c:59-68         host endianness
 *   ac_flagstotext     aclib/accore.c:76-99         adds the token "cuda"
 *   ac_parseflags      aclib/accore.c:105-167       accepts "cuda" (and "C" = none) on every host
 *   ac_memcpy          aclib/memcpy.c:16-25         host: memmove; device pointers: D2D copy
 *   ac_average         aclib/average.c:22-39        CUDA kernel; host pointers are staged
 *   ac_rescale         aclib/rescale.c:23-46        CUDA kernel; host pointers are staged
 */
#ifndef ACGPU_AC_H
#define ACGPU_AC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Acceleration bits understood by ac_init().  Values 0x0001..0x4000 are the reference's x86 bits
 * (aclib/ac.h:26-40); libacgpu parses and prints them but implements none of them. */
enum {
    AC_IA32ASM  = 0x0001,
    AC_AMD64ASM = 0x0002,
    AC_CMOVE    = 0x0004,
    AC_MMX      = 0x0008,
    AC_MMXEXT   = 0x0010,
    AC_3DNOW    = 0x0020,
    AC_3DNOWEXT = 0x0040,
    AC_SSE      = 0x0080,
    AC_SSE2     = 0x0100,
    AC_SSE3     = 0x0200,
    AC_SSSE3    = 0x0400,
    AC_SSE41    = 0x0800,
    AC_SSE42    = 0x1000,
    AC_SSE4A    = 0x2000,
    AC_SSE5     = 0x4000,
    AC_CUDA     = 0x8000   /* NEW: NVIDIA sm_100a kernels (first free bit after AC_SSE5) */
};
#define AC_NONE 0
#define AC_ALL  (~0)

#define AC_LITTLE_ENDIAN 1
#define AC_BIG_ENDIAN    2

int ac_init(int accel);
int ac_cpuinfo(void);
int ac_endian(void);
const char *ac_flagstotext(int accel);
int ac_parseflags(const char *text, int *accel);

void *ac_memcpy(void *dest, const void *src, size_t size);
void ac_average(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes);
void ac_rescale(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes,
                uint32_t weight1, uint32_t weight2);

#ifdef __cplusplus
}
#endif
#endif /* ACGPU_AC_H */
