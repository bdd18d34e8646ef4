// host_api.cu -- the C ABI of libacgpu: the aclib entry points (include/ac.h, include/imgconvert.h)
// and the batched / device-resident additions (include/acgpu.h).
//
// Host-side responsibilities, mirroring the reference's accore.c / imgconvert.c:
//   * ac_init / ac_cpuinfo / ac_parseflags / ac_flagstotext with the new AC_CUDA bit
//     (aclib/accore.c:29-167);
//   * ac_imgconvert's YV12 plane swap and pair lookup (aclib/imgconvert.c:34-64);
//   * classification of caller pointers (host pageable / host pinned / device) and staging of host
//     frames through per-thread device buffers -- callers are N concurrent frame threads
//     (src/frame_threads.c:174-228), so all mutable state is thread-local: one CUDA stream and one
//     staging arena per (thread, device);
//   * kernel-tier selection (TMA-staged -> vectorised -> generic) with NO CPU fallback: if no usable
//     device exists every entry point fails loudly.
#include "acgpu_internal.h"

#include <math.h>

#include <algorithm>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include <atomic>
#include <vector>

namespace acgpu {
namespace {

constexpr int kMaxDev = 16;
constexpr int kPipeSlots = 3;

struct Blob {               // small device-resident tables cached by content (row-op lists, weights)
    uint64_t hash;
    std::vector<uint8_t> host;   // the content itself: a hash match alone is not trusted
    void    *dptr;
};

struct DevCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t sleep_ev = nullptr;      // blocking-sync event: how a legacy call waits when many threads are converting
    uint8_t *arena = nullptr;            // device staging for the legacy host-pointer calls
    size_t   arena_cap = 0;
    cudaEvent_t arena_ev = nullptr;      // recorded after the last asynchronous use of the arena (see arena_acquire)
    uint8_t *bounce = nullptr;           // pinned host mirror of the arena (pageable caller buffers go through it)
    size_t   bounce_cap = 0;
    std::vector<Blob> blobs;
    cudaStream_t pipe_stream[kPipeSlots] = {nullptr, nullptr, nullptr};
    uint8_t *pipe_buf[kPipeSlots] = {nullptr, nullptr, nullptr};
    size_t   pipe_cap[kPipeSlots] = {0, 0, 0};
};

struct ThreadCtx {
    int      device = -1;
    DevCtx   dev[kMaxDev];
    char     err[512] = {0};
    uint64_t launches = 0;
    int      last_tier = 0;
    int      force_tier = 0;

    // A caller thread that exits gives its stream, staging buffers and cached tables back.  Best effort: at process
    // exit the CUDA runtime may already be gone, in which case these calls fail harmlessly.
    ~ThreadCtx()
    {
        for (int d = 0; d < kMaxDev; d++) {
            DevCtx &c = dev[d];
            if (!c.stream && !c.arena && !c.bounce && c.blobs.empty() && !c.pipe_stream[0]) continue;
            if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaDeviceSynchronize();     // batched calls may have run on caller-supplied streams that still read our tables
            for (int s = 0; s < kPipeSlots; s++) {
                if (c.pipe_stream[s]) { cudaStreamSynchronize(c.pipe_stream[s]); cudaStreamDestroy(c.pipe_stream[s]); }
                if (c.pipe_buf[s]) cudaFree(c.pipe_buf[s]);
            }
            for (Blob &b : c.blobs) cudaFree(b.dptr);
            if (c.arena) cudaFree(c.arena);
            if (c.bounce) cudaFreeHost(c.bounce);
            if (c.sleep_ev) cudaEventDestroy(c.sleep_ev);
            if (c.arena_ev) cudaEventDestroy(c.arena_ev);
            if (c.stream) cudaStreamDestroy(c.stream);
            cudaGetLastError();
        }
    }
};

thread_local ThreadCtx tls;
std::atomic<int> g_initialised{0};
std::atomic<int> g_sm_count{0};

int default_device()
{
    const char *e = getenv("ACGPU_DEVICE");
    return e ? atoi(e) : 0;
}

int cur_device()
{
    if (tls.device < 0) tls.device = default_device();
    return tls.device;
}

bool bind_device()
{
    const int d = cur_device();
    if (d < 0 || d >= kMaxDev) {
        set_error("device ordinal %d out of range", d);
        return false;
    }
    return check(cudaSetDevice(d), "cudaSetDevice");
}

DevCtx *ctx()
{
    if (!bind_device()) return nullptr;
    DevCtx *c = &tls.dev[cur_device()];
    if (!c->stream && !check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        return nullptr;
    return c;
}

cudaStream_t pick_stream(DevCtx *c, acgpu_stream_t s) { return s ? reinterpret_cast<cudaStream_t>(s) : c->stream; }

[[noreturn]] void fatal(const char *what)
{
    // void-returning aclib entry points (ac_average, ac_rescale) cannot report failure; a silent no-op
    // would corrupt frames, so stop the process instead.
    fprintf(stderr, "libacgpu: fatal: %s: %s\n", what, tls.err);
    abort();
}

enum PtrKind { PK_HOST = 0, PK_PINNED = 1, PK_DEVICE = 2 };
PtrKind classify(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return PK_HOST;
    }
    switch (at.type) {
    case cudaMemoryTypeDevice:
    case cudaMemoryTypeManaged: return PK_DEVICE;
    case cudaMemoryTypeHost:    return PK_PINNED;
    default:                    return PK_HOST;
    }
}

bool ensure_arena(DevCtx *c, size_t bytes)
{
    if (bytes <= c->arena_cap) return true;
    if (c->arena) {
        cudaStreamSynchronize(c->stream);
        cudaFree(c->arena);
        c->arena = nullptr;
        c->arena_cap = 0;
    }
    const size_t cap = bytes + bytes / 4 + (1 << 20);
    if (!check(cudaMalloc(&c->arena, cap), "cudaMalloc(staging arena)")) return false;
    c->arena_cap = cap;
    return true;
}

// The arena is one buffer per (thread, device) but the batched entry points run on whatever stream the caller names.
// A use of the arena on stream B must not start before an earlier, still unfinished use on stream A is over:
// arena_acquire makes B wait for the event the previous user recorded, arena_release records it.
bool arena_acquire(DevCtx *c, cudaStream_t st)
{
    if (!c->arena_ev) return true;
    return check(cudaStreamWaitEvent(st, c->arena_ev, 0), "arena wait");
}
bool arena_release(DevCtx *c, cudaStream_t st)
{
    if (!c->arena_ev && !check(cudaEventCreateWithFlags(&c->arena_ev, cudaEventDisableTiming), "cudaEventCreate")) return false;
    return check(cudaEventRecord(c->arena_ev, st), "arena record");
}

bool ensure_bounce(DevCtx *c, size_t bytes)
{
    if (bytes <= c->bounce_cap) return true;
    if (c->bounce) {
        cudaStreamSynchronize(c->stream);
        cudaFreeHost(c->bounce);
        c->bounce = nullptr;
        c->bounce_cap = 0;
    }
    const size_t cap = bytes + bytes / 4 + (1 << 20);
    if (!check(cudaHostAlloc(&c->bounce, cap, cudaHostAllocDefault), "cudaHostAlloc(bounce)")) return false;
    c->bounce_cap = cap;
    return true;
}

// Waits for the thread's stream at the end of a legacy host-pointer call.  A few concurrent callers spin (lowest
// latency); when more threads than that are inside staged calls at once -- 16 frame threads on a 16-core host -- they
// sleep on a blocking-sync event instead, so the waiters stop stealing the cores the copy submissions need
// (tools/legacy_bench.c, 16 C threads on pinned 1080p frames: 3.5 k frames/s spinning, 6.9 k sleeping; pageable frames,
// whose threads spend most of the call in their own memcpy, are better off spinning: 4.6 k vs 3.6 k).
bool wait_stream(DevCtx *c, int concurrent_callers, const char *who)
{
    if (concurrent_callers < 4) return check(cudaStreamSynchronize(c->stream), who);
    if (!c->sleep_ev && !check(cudaEventCreateWithFlags(&c->sleep_ev, cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreate"))
        return false;
    return check(cudaEventRecord(c->sleep_ev, c->stream), who) && check(cudaEventSynchronize(c->sleep_ev), who);
}

// Number of threads currently inside a staged legacy call.  A lone caller lets the driver copy straight from
// pageable memory (lowest latency: ~1.06 ms per 1080p frame); concurrent callers -- transcode's N frame threads,
// src/frame_threads.c:174-228 -- each memcpy through their own pinned bounce buffer so the PCIe copies are true
// async DMA and do not serialise behind the driver's single pageable staging path (measured: 2.2 k frames/s flat).
std::atomic<int> g_staged_calls{0};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// grid.y carries the frame index (<= 65535): longer batches are cut into launches of 32768 frames
template <class F>
int per_frame_chunk(int nframes, F launch)
{
    for (int f0 = 0; f0 < nframes; f0 += 32768)
        if (!launch(f0, nframes - f0 < 32768 ? nframes - f0 : 32768)) return 0;
    return 1;
}


void plane_sizes(int fmt, int w, int h, size_t out[3], int *np)
{
    const FmtDesc d = describe(fmt);
    if (d.kind == K_PLANAR) {
        out[0] = (size_t)w * h;
        out[1] = out[2] = chroma_plane_bytes(fmt, w, h);
        *np = 3;
    } else {
        out[0] = frame_bytes(fmt, w, h);
        out[1] = out[2] = 0;
        *np = 1;
    }
}

void size_units(int fmt, int *uw, int *uh)
{
    switch (fmt) {
    case IMG_YUV420P: *uw = 2; *uh = 2; break;
    case IMG_YUV411P: *uw = 4; *uh = 1; break;
    case IMG_YUV422P: case IMG_YUY2: case IMG_UYVY: case IMG_YVYU: *uw = 2; *uh = 1; break;
    default: *uw = 1; *uh = 1; break;
    }
}

// Does the C path write every byte of the destination frame?  If not (alpha left alone by
// yuv->rgb32, img_yuv_rgb.c:62-64; tails left alone on off-grid sizes) a host destination must be
// uploaded first so the untouched bytes survive the round trip.
bool overwrites_whole_dest(int sfmt, int dfmt, int w, int h)
{
    const FmtDesc sd = describe(sfmt), dd = describe(dfmt);
    if (dd.kind == K_RGB && dd.bpp == 4 && (sd.kind == K_PLANAR || sd.kind == K_PACKED || sd.kind == K_Y8)) return false;
    int uw1, uh1, uw2, uh2;
    size_units(sfmt, &uw1, &uh1);
    size_units(dfmt, &uw2, &uh2);
    const int uw = uw1 > uw2 ? uw1 : uw2, uh = uh1 > uh2 ? uh1 : uh2;
    return w % uw == 0 && h % uh == 0;
}

bool run_convert(const ConvertArgs &a)
{
    // Automatic order: vectorised tier, else generic.  Tier 3 (bulk/TMA stores) is selectable but NOT the default:
    // measured 2-3 % slower than tier 2 on the headline pair (profiles/r1_experiments.md).
    const int force = tls.force_tier;
    tls.err[0] = 0;      // a tier that declines a call leaves this empty; a failed launch leaves its CUDA error
    if (force == 3) {
        if (convert_tma(a)) { tls.last_tier = 3; return true; }
        if (!tls.err[0]) set_error("tier 3 (bulk stores) does not cover this pair/size/alignment");
        return false;
    }
    if ((force == 0 || force == 2) && convert_fast(a)) { tls.last_tier = 2; return true; }
    if (tls.err[0]) return false;                       // a launch failed: do not paper over it with another tier
    if (force == 2) { set_error("tier 2 (vectorised) does not cover this pair/size/alignment"); return false; }
    if (!convert_generic(a)) return false;
    tls.last_tier = 1;
    return true;
}

bool fold_yv12(uint8_t *const *src, int *sfmt, uint8_t *const *dst, int *dfmt, Image *si, Image *di)
{
    // aclib/imgconvert.c:40-56
    const FmtDesc sd0 = describe(*sfmt == IMG_YV12 ? IMG_YUV420P : *sfmt);
    const FmtDesc dd0 = describe(*dfmt == IMG_YV12 ? IMG_YUV420P : *dfmt);
    if (sd0.kind == K_NONE || dd0.kind == K_NONE) return false;
    si->p[0] = src[0];
    di->p[0] = dst[0];
    si->p[1] = si->p[2] = di->p[1] = di->p[2] = nullptr;
    if (sd0.kind == K_PLANAR) { si->p[1] = src[1]; si->p[2] = src[2]; }
    if (dd0.kind == K_PLANAR) { di->p[1] = dst[1]; di->p[2] = dst[2]; }
    if (*sfmt == IMG_YV12) { *sfmt = IMG_YUV420P; uint8_t *t = si->p[1]; si->p[1] = si->p[2]; si->p[2] = t; }
    if (*dfmt == IMG_YV12) { *dfmt = IMG_YUV420P; uint8_t *t = di->p[1]; di->p[1] = di->p[2]; di->p[2] = t; }
    return true;
}

// One frame through ac_imgconvert's legacy signature; planes may live on the host or on the device.
bool convert_one(Image si, int sfmt, Image di, int dfmt, int w, int h)
{
    DevCtx *c = ctx();
    if (!c) return false;
    size_t ssz[3], dsz[3];
    int snp, dnp;
    plane_sizes(sfmt, w, h, ssz, &snp);
    plane_sizes(dfmt, w, h, dsz, &dnp);
    const bool src_host = classify(si.p[0]) != PK_DEVICE;
    const bool dst_host = classify(di.p[0]) != PK_DEVICE;

    size_t need = 0, soff[3] = {0, 0, 0}, doff[3] = {0, 0, 0};
    if (src_host)
        for (int p = 0; p < snp; p++) { soff[p] = need; need += align_up(ssz[p] + 16, 256); }
    if (dst_host)
        for (int p = 0; p < dnp; p++) { doff[p] = need; need += align_up(dsz[p] + 16, 256); }
    if (need && (!ensure_arena(c, need) || !arena_acquire(c, c->stream))) return false;

    ConvertArgs a{};
    a.srcfmt = sfmt; a.dstfmt = dfmt; a.w = w; a.h = h; a.nframes = 1; a.stream = c->stream;
    a.src = si; a.dst = di;
    a.src.pitch = a.dst.pitch = 0;
    // pageable caller memory: through the pinned bounce buffer when other threads are converting too
    const bool src_pageable = src_host && classify(si.p[0]) == PK_HOST;
    const bool dst_pageable = dst_host && classify(di.p[0]) == PK_HOST;
    struct Busy {
        bool on;
        int  others;
        explicit Busy(bool o) : on(o), others(o ? g_staged_calls.fetch_add(1) : 0) {}
        ~Busy() { if (on) g_staged_calls.fetch_sub(1); }
    } busy(src_host || dst_host);
    const bool bounce = (src_pageable || dst_pageable) && busy.others > 0 && ensure_bounce(c, need);
    if (src_host)
        for (int p = 0; p < snp; p++) {
            a.src.p[p] = c->arena + soff[p];
            const uint8_t *from = si.p[p];
            if (bounce && src_pageable) {
                memcpy(c->bounce + soff[p], si.p[p], ssz[p]);
                from = c->bounce + soff[p];
            }
            if (!check(cudaMemcpyAsync(a.src.p[p], from, ssz[p], cudaMemcpyHostToDevice, c->stream), "H2D src plane"))
                return false;
        }
    if (dst_host) {
        const bool preload = !overwrites_whole_dest(sfmt, dfmt, w, h);
        for (int p = 0; p < dnp; p++) {
            a.dst.p[p] = c->arena + doff[p];
            if (!preload) continue;
            const uint8_t *from = di.p[p];
            if (bounce && dst_pageable) {
                memcpy(c->bounce + doff[p], di.p[p], dsz[p]);
                from = c->bounce + doff[p];
            }
            if (!check(cudaMemcpyAsync(a.dst.p[p], from, dsz[p], cudaMemcpyHostToDevice, c->stream), "H2D dest plane"))
                return false;
        }
    }
    if (!run_convert(a)) return false;
    if (dst_host)
        for (int p = 0; p < dnp; p++) {
            uint8_t *to = (bounce && dst_pageable) ? c->bounce + doff[p] : di.p[p];
            if (!check(cudaMemcpyAsync(to, a.dst.p[p], dsz[p], cudaMemcpyDeviceToHost, c->stream), "D2H dest plane"))
                return false;
        }
    // callers that also memcpy through the bounce buffer do better spinning at every thread count measured
    if (!wait_stream(c, bounce ? 0 : busy.others, "ac_imgconvert")) return false;
    if (dst_host && bounce && dst_pageable)
        for (int p = 0; p < dnp; p++) memcpy(di.p[p], c->bounce + doff[p], dsz[p]);
    return true;
}

uint64_t fnv1a(const void *p, size_t n)
{
    const uint8_t *b = static_cast<const uint8_t *>(p);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Returns a device copy of a small host table, cached per (thread, device) by content.
constexpr size_t kBlobCacheEntries = 64;
void *device_blob(DevCtx *c, const void *host, size_t bytes, cudaStream_t st)
{
    // The list is kept in least-recently-used order: a hit moves to the back, a full cache drops its OLDER HALF.  One API
    // call asks for a handful of tables at most, so a table handed out earlier in the same call (hit or miss) is never
    // among the ones freed by a later miss of that call.
    const uint64_t h = fnv1a(host, bytes);
    for (size_t i = 0; i < c->blobs.size(); i++) {
        Blob &b = c->blobs[i];
        if (b.hash == h && b.host.size() == bytes && memcmp(b.host.data(), host, bytes) == 0) {
            void *d = b.dptr;
            if (i + 1 != c->blobs.size()) std::rotate(c->blobs.begin() + (ptrdiff_t)i, c->blobs.begin() + (ptrdiff_t)i + 1, c->blobs.end());
            return d;
        }
    }
    if (c->blobs.size() >= kBlobCacheEntries) {
        cudaStreamSynchronize(c->stream);
        cudaDeviceSynchronize();          // kernels on caller-supplied streams may still be reading the old tables
        const size_t drop = c->blobs.size() / 2;
        for (size_t i = 0; i < drop; i++) cudaFree(c->blobs[i].dptr);
        c->blobs.erase(c->blobs.begin(), c->blobs.begin() + (ptrdiff_t)drop);
    }
    void *d = nullptr;
    if (!check(cudaMalloc(&d, bytes ? bytes : 16), "cudaMalloc(table)")) return nullptr;
    // pageable source: the runtime stages it before returning, so `host` may die right after
    if (!check(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, st), "H2D table")
        || !check(cudaStreamSynchronize(st), "H2D table")) {   // later calls may use another stream
        cudaFree(d);
        return nullptr;
    }
    Blob nb;
    nb.hash = h;
    nb.host.assign(static_cast<const uint8_t *>(host), static_cast<const uint8_t *>(host) + bytes);
    nb.dptr = d;
    c->blobs.push_back(std::move(nb));
    return d;
}

// libtcvideo/tcvideo.c:1138-1165 -- sin^2-weighted two-tap table, newsize/8 entries.
void build_resize_table(int oldsize, int newsize, std::vector<int32_t> &src, std::vector<uint32_t> &w1,
                        std::vector<uint32_t> &w2)
{
    const int n = newsize / 8;
    src.resize(n); w1.resize(n); w2.resize(n);
    const double ratio = (double)oldsize / (double)newsize;
    for (int i = 0; i < n; i++) {
        const double pos = (double)i * (double)oldsize / (double)newsize;
        const int s = (int)pos;
        src[i] = s;
        if (pos + ratio < s + 1) {
            w1[i] = 65536; w2[i] = 0;
        } else {
            const double t = ((s + 1) - pos) / ratio * M_PI / 2;
            w1[i] = (uint32_t)(sin(t) * sin(t) * 65536 + 0.5);
            w2[i] = 65536 - w1[i];
        }
    }
}

bool device_usable(int *sms)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return false; }
    int d = default_device();
    if (d < 0 || d >= n) d = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) != cudaSuccess) { cudaGetLastError(); return false; }
    if (sms) *sms = prop.multiProcessorCount;
    // the fatbin holds sm_100a SASS only: arch-specific code runs on compute capability 10.0 and nothing else
    return prop.major == 10 && prop.minor == 0;
}

}  // namespace

// ---- bookkeeping shared with the kernel files -------------------------------------------------------
void note_launch(int n) { tls.launches += (uint64_t)n; }

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls.err, sizeof(tls.err), fmt, ap);
    va_end(ap);
    if (getenv("ACGPU_VERBOSE")) fprintf(stderr, "libacgpu: %s\n", tls.err);
}

bool check(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return false;
}

int sm_count()
{
    int n = g_sm_count.load();
    if (n <= 0) {
        device_usable(&n);
        g_sm_count.store(n);
    }
    return n > 0 ? n : 148;
}

}  // namespace acgpu

using namespace acgpu;

// =====================================================================================================
// aclib core (include/ac.h)
// =====================================================================================================
extern "C" {

int ac_cpuinfo(void) { return device_usable(nullptr) ? AC_CUDA : 0; }

int ac_imgconvert_init(int accel)
{
    if (!(accel & AC_CUDA)) {
        set_error("ac_imgconvert_init: accel 0x%x lacks AC_CUDA and libacgpu has no CPU implementation", accel);
        return 0;
    }
    return 1;
}

int ac_init(int accel)
{
    // aclib/accore.c:29-40: mask with what the machine offers, then run the sub-initialisers.
    accel &= ac_cpuinfo();
    if (!(accel & AC_CUDA)) {
        g_initialised.store(0);
        set_error("ac_init: no usable CUDA device (need compute capability 10.0) or AC_CUDA not requested; "
                  "libacgpu has no CPU fallback");
        fprintf(stderr, "libacgpu: %s\n", tls.err);
        return 0;
    }
    if (!ac_imgconvert_init(accel)) return 0;
    if (!ctx()) {
        fprintf(stderr, "libacgpu: ac_init: %s\n", tls.err);
        return 0;
    }
    g_initialised.store(1);
    return 1;
}

int ac_endian(void)
{
    const uint16_t probe = 1;
    return *reinterpret_cast<const uint8_t *>(&probe) ? AC_LITTLE_ENDIAN : AC_BIG_ENDIAN;
}

static const struct { int bit; const char *name; } kFlagNames[] = {
    {AC_CUDA, "cuda"},     {AC_SSE5, "sse5"},   {AC_SSE4A, "sse4a"},       {AC_SSE42, "sse42"},
    {AC_SSE41, "sse41"},   {AC_SSSE3, "ssse3"}, {AC_SSE3, "sse3"},         {AC_SSE2, "sse2"},
    {AC_SSE, "sse"},       {AC_3DNOWEXT, "3dnowext"}, {AC_3DNOW, "3dnow"}, {AC_MMXEXT, "mmxext"},
    {AC_MMX, "mmx"},       {AC_CMOVE, "cmove"}, {AC_IA32ASM | AC_AMD64ASM, "asm"},
};

const char *ac_flagstotext(int accel)
{
    // Same vocabulary and order as aclib/accore.c:76-99 (most capable first), plus "cuda".
    static thread_local char buf[256];
    if (!accel) return "none";
    size_t len = 0;
    buf[0] = 0;
    for (const auto &f : kFlagNames)
        if (accel & f.bit) len += snprintf(buf + len, sizeof(buf) - len, "%s%s", len ? " " : "", f.name);
    return buf;
}

int ac_parseflags(const char *text, int *accel)
{
    // Comma-separated tokens, case-insensitive, "C" = no acceleration (aclib/accore.c:105-167).
    if (!text || !accel) return 0;
    *accel = 0;
    const char *p = text;
    for (;;) {
        const char *comma = strchr(p, ',');
        size_t len = comma ? (size_t)(comma - p) : strlen(p);
        if (len > 16) len = 16;
        char tok[17];
        memcpy(tok, p, len);
        tok[len] = 0;
        if (strcasecmp(tok, "C") == 0) {
            /* no bits */
        } else {
            int bit = 0;
            for (const auto &f : kFlagNames)
                if (strcasecmp(tok, f.name) == 0) bit = f.bit;
            if (!bit) return 0;
            if (bit == (AC_IA32ASM | AC_AMD64ASM)) bit = sizeof(void *) == 8 ? AC_AMD64ASM : AC_IA32ASM;
            *accel |= bit;
        }
        if (!comma) break;
        p = comma + 1;
    }
    return 1;
}

void *ac_memcpy(void *dest, const void *src, size_t size)
{
    // aclib/memcpy.c:16-25 is memmove (ascending copy guarantee, ac.h:80-82).  It is not pixel math, so
    // host buffers stay on the host; device buffers are copied on the device.
    if (size == 0 || dest == src) return dest;
    const PtrKind kd = classify(dest), ks = classify(src);
    if (kd != PK_DEVICE && ks != PK_DEVICE) return memmove(dest, src, size);
    DevCtx *c = ctx();
    if (!c || !check(cudaMemcpyAsync(dest, src, size, cudaMemcpyDefault, c->stream), "ac_memcpy")
        || !check(cudaStreamSynchronize(c->stream), "ac_memcpy"))
        fatal("ac_memcpy");
    return dest;
}

static void blend_legacy(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes,
                         uint32_t w1, uint32_t w2, int op, const char *who)
{
    if (bytes <= 0) return;
    DevCtx *c = ctx();
    if (!c) fatal(who);
    const bool h1 = classify(src1) != PK_DEVICE, h2 = classify(src2) != PK_DEVICE, hd = classify(dest) != PK_DEVICE;
    const size_t slot = align_up((size_t)bytes, 256);
    if ((h1 || h2 || hd) && (!ensure_arena(c, 3 * slot) || !arena_acquire(c, c->stream))) fatal(who);
    const uint8_t *d1 = src1, *d2 = src2;
    uint8_t *dd = dest;
    bool ok = true;
    if (h1) { ok = ok && check(cudaMemcpyAsync(c->arena, src1, bytes, cudaMemcpyHostToDevice, c->stream), who); d1 = c->arena; }
    if (h2) { ok = ok && check(cudaMemcpyAsync(c->arena + slot, src2, bytes, cudaMemcpyHostToDevice, c->stream), who); d2 = c->arena + slot; }
    if (hd) dd = c->arena + 2 * slot;
    ok = ok && blend_launch(d1, d2, dd, (size_t)bytes, w1, w2, op, c->stream);
    if (hd) ok = ok && check(cudaMemcpyAsync(dest, dd, bytes, cudaMemcpyDeviceToHost, c->stream), who);
    ok = ok && check(cudaStreamSynchronize(c->stream), who);
    if (!ok) fatal(who);
}

void ac_average(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes)
{
    blend_legacy(src1, src2, dest, bytes, 0, 0, ACGPU_ROW_AVERAGE, "ac_average");
}

void ac_rescale(const uint8_t *src1, const uint8_t *src2, uint8_t *dest, int bytes, uint32_t weight1, uint32_t weight2)
{
    // aclib/rescale.c:23-32: the copy branches never touch the other source.
    if (bytes <= 0) return;
    if (weight1 >= 0x10000u) { ac_memcpy(dest, src1, (size_t)bytes); return; }
    if (weight2 >= 0x10000u) { ac_memcpy(dest, src2, (size_t)bytes); return; }
    blend_legacy(src1, src2, dest, bytes, weight1, weight2, ACGPU_ROW_RESCALE, "ac_rescale");
}

// =====================================================================================================
// ac_imgconvert (include/imgconvert.h)
// =====================================================================================================
int ac_imgconvert(uint8_t **src, ImageFormat srcfmt, uint8_t **dest, ImageFormat destfmt, int width, int height)
{
    if (!g_initialised.load()) {
        // the reference's table is empty before ac_init -> every pair "unknown" (imgconvert.c:58-63)
        set_error("ac_imgconvert called before a successful ac_init(AC_CUDA)");
        return 0;
    }
    if (!src || !dest) return 0;
    int sfmt = srcfmt, dfmt = destfmt;
    Image si{}, di{};
    if (!fold_yv12(src, &sfmt, dest, &dfmt, &si, &di)) return 0;
    if (width <= 0 || height <= 0) return 1;    // the C loops simply do not iterate
    return convert_one(si, sfmt, di, dfmt, width, height) ? 1 : 0;
}

// =====================================================================================================
// libacgpu additions (include/acgpu.h)
// =====================================================================================================
const char *acgpu_version(void) { return "libacgpu 0.1 (sm_100a)"; }
const char *acgpu_last_error(void) { return tls.err; }

int acgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int acgpu_set_device(int ordinal)
{
    if (ordinal < 0 || ordinal >= kMaxDev || ordinal >= acgpu_device_count()) {
        set_error("acgpu_set_device: no device %d", ordinal);
        return 0;
    }
    tls.device = ordinal;
    return bind_device() ? 1 : 0;
}

int acgpu_get_device(void) { return cur_device(); }
int acgpu_device_sm_count(void) { return sm_count(); }
int acgpu_last_kernel_tier(void) { return tls.last_tier; }
void acgpu_force_tier(int tier) { tls.force_tier = tier; }

uint64_t acgpu_launch_count(int reset)
{
    const uint64_t n = tls.launches;
    if (reset) tls.launches = 0;
    return n;
}

void *acgpu_malloc(size_t bytes)
{
    void *p = nullptr;
    if (!bind_device() || !check(cudaMalloc(&p, bytes ? bytes : 1), "acgpu_malloc")) return nullptr;
    return p;
}
void acgpu_free(void *dptr) { if (dptr && bind_device()) cudaFree(dptr); }

void *acgpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (!bind_device() || !check(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable), "acgpu_host_alloc")) return nullptr;
    return p;
}
void acgpu_host_free(void *hptr) { if (hptr) cudaFreeHost(hptr); }

static int copy_async(void *d, const void *s, size_t n, cudaMemcpyKind k, acgpu_stream_t st, const char *who)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    return check(cudaMemcpyAsync(d, s, n, k, pick_stream(c, st)), who) ? 1 : 0;
}
int acgpu_memcpy_h2d(void *d, const void *h, size_t n, acgpu_stream_t st) { return copy_async(d, h, n, cudaMemcpyHostToDevice, st, "acgpu_memcpy_h2d"); }
int acgpu_memcpy_d2h(void *h, const void *d, size_t n, acgpu_stream_t st) { return copy_async(h, d, n, cudaMemcpyDeviceToHost, st, "acgpu_memcpy_d2h"); }
int acgpu_memcpy_d2d(void *d, const void *s, size_t n, acgpu_stream_t st) { return copy_async(d, s, n, cudaMemcpyDeviceToDevice, st, "acgpu_memcpy_d2d"); }

int acgpu_memset(void *dptr, int value, size_t bytes, acgpu_stream_t st)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    return check(cudaMemsetAsync(dptr, value, bytes, pick_stream(c, st)), "acgpu_memset") ? 1 : 0;
}

acgpu_stream_t acgpu_stream_create(void)
{
    cudaStream_t s = nullptr;
    if (!bind_device() || !check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "acgpu_stream_create")) return nullptr;
    return reinterpret_cast<acgpu_stream_t>(s);
}
void acgpu_stream_destroy(acgpu_stream_t s) { if (s) cudaStreamDestroy(reinterpret_cast<cudaStream_t>(s)); }

int acgpu_stream_sync(acgpu_stream_t s)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    return check(cudaStreamSynchronize(pick_stream(c, s)), "acgpu_stream_sync") ? 1 : 0;
}

acgpu_event_t acgpu_event_create(void)
{
    cudaEvent_t e = nullptr;
    if (!bind_device() || !check(cudaEventCreate(&e), "acgpu_event_create")) return nullptr;
    return reinterpret_cast<acgpu_event_t>(e);
}
void acgpu_event_destroy(acgpu_event_t e) { if (e) cudaEventDestroy(reinterpret_cast<cudaEvent_t>(e)); }

int acgpu_event_record(acgpu_event_t e, acgpu_stream_t s)
{
    DevCtx *c = ctx();
    if (!c || !e) return 0;
    return check(cudaEventRecord(reinterpret_cast<cudaEvent_t>(e), pick_stream(c, s)), "acgpu_event_record") ? 1 : 0;
}
int acgpu_event_sync(acgpu_event_t e) { return e && check(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(e)), "acgpu_event_sync") ? 1 : 0; }

float acgpu_event_elapsed_ms(acgpu_event_t a, acgpu_event_t b)
{
    float ms = -1.0f;
    if (!a || !b || !check(cudaEventElapsedTime(&ms, reinterpret_cast<cudaEvent_t>(a), reinterpret_cast<cudaEvent_t>(b)), "acgpu_event_elapsed_ms"))
        return -1.0f;
    return ms;
}

// ---- batched conversion ------------------------------------------------------------------------------
int acgpu_imgconvert_batch(uint8_t *const *src, ImageFormat srcfmt, size_t src_frame_pitch,
                           uint8_t *const *dest, ImageFormat destfmt, size_t dest_frame_pitch,
                           int width, int height, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!src || !dest) { set_error("acgpu_imgconvert_batch: null plane array"); return 0; }
    int sfmt = srcfmt, dfmt = destfmt;
    ConvertArgs a{};
    if (!fold_yv12(src, &sfmt, dest, &dfmt, &a.src, &a.dst)) {
        set_error("acgpu_imgconvert_batch: unknown format pair 0x%x -> 0x%x", (int)srcfmt, (int)destfmt);
        return 0;
    }
    if (width <= 0 || height <= 0 || nframes <= 0) return 1;
    a.srcfmt = sfmt; a.dstfmt = dfmt; a.w = width; a.h = height;
    a.src.pitch = src_frame_pitch; a.dst.pitch = dest_frame_pitch;
    a.stream = pick_stream(c, stream);
    // grid.y carries the frame index: split very large batches
    for (int f0 = 0; f0 < nframes; f0 += 32768) {
        ConvertArgs b = a;
        b.nframes = nframes - f0 < 32768 ? nframes - f0 : 32768;
        for (int p = 0; p < 3; p++) {
            if (b.src.p[p]) b.src.p[p] += (size_t)f0 * src_frame_pitch;
            if (b.dst.p[p]) b.dst.p[p] += (size_t)f0 * dest_frame_pitch;
        }
        if (!run_convert(b)) return 0;
    }
    return 1;
}

int acgpu_imgconvert_frames_host(const uint8_t *src_frames, ImageFormat srcfmt, uint8_t *dest_frames,
                                 ImageFormat destfmt, int width, int height, int nframes)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    const int sf = srcfmt == IMG_YV12 ? IMG_YUV420P : (int)srcfmt, df = destfmt == IMG_YV12 ? IMG_YUV420P : (int)destfmt;
    if (describe(sf).kind == K_NONE || describe(df).kind == K_NONE) { set_error("unknown format pair"); return 0; }
    if (width <= 0 || height <= 0 || nframes <= 0) return 1;
    const size_t sfb = frame_bytes(sf, width, height), dfb = frame_bytes(df, width, height);
    const size_t sp = align_up(sfb, 256), dp = align_up(dfb, 256);     // device frame pitches
    const bool preload = !overwrites_whole_dest(sf, df, width, height);
    // ~16 MiB of the larger side per chunk: big enough to amortise launches and reach full PCIe rate (tools/pcie_probe.py:
    // 16 MB copies already run at 56 GB/s), small enough that the fill/drain bubbles of the pipeline stay short
    size_t per = (size_t)(16u << 20) / (sp > dp ? sp : dp);
    if (per < 1) per = 1;
    if (per > (size_t)nframes) per = nframes;
    const size_t slot_bytes = per * (sp + dp);
    for (int s = 0; s < kPipeSlots; s++) {
        if (!c->pipe_stream[s] && !check(cudaStreamCreateWithFlags(&c->pipe_stream[s], cudaStreamNonBlocking), "pipe stream")) return 0;
        if (c->pipe_cap[s] < slot_bytes) {
            if (c->pipe_buf[s]) { cudaStreamSynchronize(c->pipe_stream[s]); cudaFree(c->pipe_buf[s]); c->pipe_buf[s] = nullptr; c->pipe_cap[s] = 0; }
            if (!check(cudaMalloc(&c->pipe_buf[s], slot_bytes), "cudaMalloc(pipeline slot)")) return 0;
            c->pipe_cap[s] = slot_bytes;
        }
    }
    int chunk = 0;
    for (int f0 = 0; f0 < nframes; f0 += (int)per, chunk++) {
        const int s = chunk % kPipeSlots;
        const int n = nframes - f0 < (int)per ? nframes - f0 : (int)per;
        cudaStream_t st = c->pipe_stream[s];
        uint8_t *dsrc = c->pipe_buf[s], *ddst = dsrc + per * sp;
        if (!check(cudaMemcpy2DAsync(dsrc, sp, src_frames + (size_t)f0 * sfb, sfb, sfb, n, cudaMemcpyHostToDevice, st), "H2D frames")) return 0;
        if (preload && !check(cudaMemcpy2DAsync(ddst, dp, dest_frames + (size_t)f0 * dfb, dfb, dfb, n, cudaMemcpyHostToDevice, st), "H2D dest frames")) return 0;
        uint8_t *sp3[3], *dp3[3];
        sp3[0] = dsrc; sp3[1] = dsrc + (size_t)width * height; sp3[2] = sp3[1] + chroma_plane_bytes(sf, width, height);
        dp3[0] = ddst; dp3[1] = ddst + (size_t)width * height; dp3[2] = dp3[1] + chroma_plane_bytes(df, width, height);
        if (!acgpu_imgconvert_batch(sp3, srcfmt, sp, dp3, destfmt, dp, width, height, n, reinterpret_cast<acgpu_stream_t>(st))) return 0;
        if (!check(cudaMemcpy2DAsync(dest_frames + (size_t)f0 * dfb, dfb, ddst, dp, dfb, n, cudaMemcpyDeviceToHost, st), "D2H frames")) return 0;
    }
    for (int s = 0; s < kPipeSlots; s++)
        if (!check(cudaStreamSynchronize(c->pipe_stream[s]), "acgpu_imgconvert_frames_host")) return 0;
    return 1;
}

// ---- row operations -----------------------------------------------------------------------------------
int acgpu_rowops_run(const uint8_t *src, size_t spitch, uint8_t *dest, size_t dpitch, const acgpu_rowop *ops,
                     int nops, int row_bytes, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nops <= 0 || row_bytes <= 0 || nframes <= 0) return 1;
    if (!ops) { set_error("acgpu_rowops_run: null op list"); return 0; }
    cudaStream_t st = pick_stream(c, stream);
    bool al = true;
    for (int i = 0; i < nops && al; i++) {
        const acgpu_rowop &o = ops[i];
        al = ((o.src1_off | o.dest_off) & 15) == 0;
        const bool two = o.op == ACGPU_ROW_AVERAGE || o.op == ACGPU_ROW_AVERAGE3
                      || (o.op == ACGPU_ROW_RESCALE && o.weight1 < 0x10000u);
        if (two) al = al && (o.src2_off & 15) == 0;
        if (o.op == ACGPU_ROW_AVERAGE3) al = al && (o.src3_off & 15) == 0;
    }
    const bool vec = al && rowops_vectorisable(src, spitch, dest, dpitch, row_bytes);
    if (vec) {
        std::vector<RowBlk> blks;
        const int max_nsrc = build_row_blocks(ops, nops, blks);
        const RowBlk *d_blks = static_cast<const RowBlk *>(device_blob(c, blks.data(), sizeof(RowBlk) * blks.size(), st));
        if (!d_blks) return 0;
        for (int f0 = 0; f0 < nframes; f0 += 32768) {
            const int n = nframes - f0 < 32768 ? nframes - f0 : 32768;
            if (!rowops_tiled_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_blks,
                                     (int)blks.size(), max_nsrc, row_bytes, n, st))
                return 0;
        }
        return 1;
    }
    const acgpu_rowop *d_ops = static_cast<const acgpu_rowop *>(device_blob(c, ops, sizeof(acgpu_rowop) * (size_t)nops, st));
    if (!d_ops) return 0;
    for (int f0 = 0; f0 < nframes; f0 += 32768) {
        const int n = nframes - f0 < 32768 ? nframes - f0 : 32768;
        if (!rowops_bytes_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_ops, nops, row_bytes, n, st))
            return 0;
    }
    return 1;
}

int acgpu_average(const uint8_t *s1, const uint8_t *s2, uint8_t *d, size_t bytes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    return blend_launch(s1, s2, d, bytes, 0, 0, ACGPU_ROW_AVERAGE, pick_stream(c, stream)) ? 1 : 0;
}

int acgpu_rescale(const uint8_t *s1, const uint8_t *s2, uint8_t *d, size_t bytes, uint32_t w1, uint32_t w2,
                  acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    cudaStream_t st = pick_stream(c, stream);
    if (bytes == 0) return 1;
    if (w1 >= 0x10000u) return s1 == d ? 1 : (check(cudaMemcpyAsync(d, s1, bytes, cudaMemcpyDeviceToDevice, st), "acgpu_rescale") ? 1 : 0);
    if (w2 >= 0x10000u) return s2 == d ? 1 : (check(cudaMemcpyAsync(d, s2, bytes, cudaMemcpyDeviceToDevice, st), "acgpu_rescale") ? 1 : 0);
    return blend_launch(s1, s2, d, bytes, w1, w2, ACGPU_ROW_RESCALE, st) ? 1 : 0;
}

// ---- libtcvideo shapes ----------------------------------------------------------------------------------
int acgpu_deinterlace_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int mode,
                            size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:290-311 argument checks
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("acgpu_deinterlace_batch: invalid frame parameters"); return 0; }
    if (mode < ACGPU_DEINT_INTERPOLATE || mode > ACGPU_DEINT_DROP_FIELD_BOTTOM) { set_error("acgpu_deinterlace_batch: invalid mode %d", mode); return 0; }
    const int64_t Bpl = (int64_t)width * Bpp;
    if (mode == ACGPU_DEINT_DROP_FIELD_TOP || mode == ACGPU_DEINT_DROP_FIELD_BOTTOM) {
        // tcvideo.c:326-338: keep every other line, starting at line 1 when the top field is dropped
        std::vector<acgpu_rowop> ops((size_t)(height / 2));
        for (int y = 0; y < height / 2; y++) {
            acgpu_rowop o{};
            o.op = ACGPU_ROW_COPY;
            o.src1_off = (int64_t)(2 * y + (mode == ACGPU_DEINT_DROP_FIELD_TOP ? 1 : 0)) * Bpl;
            o.dest_off = y * Bpl;
            ops[(size_t)y] = o;
        }
        return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), height / 2, (int)Bpl, nframes, stream);
    }
    std::vector<acgpu_rowop> ops((size_t)height);
    for (int y = 0; y < height; y++) {
        acgpu_rowop o{};
        o.dest_off = y * Bpl;
        if (mode == ACGPU_DEINT_INTERPOLATE || height < 2) {
            // tcvideo.c:353-364: even rows copied, odd rows = mean of neighbours, odd last row = copy of y-1
            if (y % 2 == 0)            { o.op = ACGPU_ROW_COPY;    o.src1_off = y * Bpl; }
            else if (y == height - 1)  { o.op = ACGPU_ROW_COPY;    o.src1_off = (y - 1) * Bpl; }
            else                       { o.op = ACGPU_ROW_AVERAGE; o.src1_off = (y - 1) * Bpl; o.src2_off = (y + 1) * Bpl; }
        } else {
            // tcvideo.c:368-389 fused: out = avg(A, B) where one of A/B is the source row and the other the
            // mean of its neighbours -- the same expression for odd and even interior rows; the first and
            // last rows average with their single neighbour (the copies made at :377 and :381).
            if (y == 0)               { o.op = ACGPU_ROW_AVERAGE; o.src1_off = Bpl;           o.src2_off = 0; }
            else if (y == height - 1) { o.op = ACGPU_ROW_AVERAGE; o.src1_off = (y - 1) * Bpl; o.src2_off = y * Bpl; }
            else { o.op = ACGPU_ROW_AVERAGE3; o.src1_off = (y - 1) * Bpl; o.src2_off = (y + 1) * Bpl; o.src3_off = y * Bpl; }
        }
        ops[(size_t)y] = o;
    }
    return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), height, (int)Bpl, nframes, stream);
}

int acgpu_resize_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int resize_w, int resize_h,
                       int scale_w, int scale_h, size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:436-457 argument checks
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("acgpu_resize_batch: invalid frame parameters"); return 0; }
    auto ok_scale = [](int s) { return s == 1 || s == 2 || s == 4 || s == 8; };
    if (!ok_scale(scale_w) || !ok_scale(scale_h)) { set_error("acgpu_resize_batch: invalid scale parameters"); return 0; }
    if (width % scale_w != 0 || height % scale_h != 0) { set_error("acgpu_resize_batch: scale does not divide the frame"); return 0; }
    if (resize_w && resize_h) { set_error("acgpu_resize_batch: only one of resize_w / resize_h may be non-zero"); return 0; }
    const int new_w = width + resize_w * scale_w, new_h = height + resize_h * scale_h;
    if (new_w <= 0 || new_h <= 0) { set_error("acgpu_resize_batch: resulting size is not positive"); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    cudaStream_t st = pick_stream(c, stream);
    std::vector<int32_t> ts;
    std::vector<uint32_t> w1, w2;
    if (resize_h) {
        const int64_t Bpl = (int64_t)width * Bpp;
        build_resize_table(height * 8 / scale_h, new_h * 8 / scale_h, ts, w1, w2);
        const int rows = new_h / scale_h;
        std::vector<acgpu_rowop> ops((size_t)rows * scale_h);
        for (int i = 0; i < scale_h; i++)
            for (int y = 0; y < rows; y++) {
                acgpu_rowop o{};
                o.op = ACGPU_ROW_RESCALE;
                o.src1_off = ((int64_t)i * (height / scale_h) + ts[y]) * Bpl;
                o.src2_off = o.src1_off + Bpl;
                o.dest_off = ((int64_t)i * rows + y) * Bpl;
                o.weight1 = w1[y];
                o.weight2 = w2[y];
                ops[(size_t)i * rows + y] = o;
            }
        return acgpu_rowops_run(src, spitch, dest, dpitch, ops.data(), (int)ops.size(), (int)Bpl, nframes, stream);
    }
    if (resize_w) {
        build_resize_table(width * 8 / scale_w, new_w * 8 / scale_w, ts, w1, w2);
        const int n = (int)ts.size();
        const size_t sp_ = spitch ? spitch : (size_t)width * height * Bpp, dp_ = dpitch ? dpitch : (size_t)new_w * new_h * Bpp;
        if (resize_h_vectorisable(src, sp_, dest, dp_, width, new_w, Bpp)) {
            // per-row byte tables: source byte offset (first tap) and packed weights for every output byte of a row
            const int src_block = width / scale_w, dst_block = new_w / scale_w;
            std::vector<uint16_t> off((size_t)new_w * Bpp);
            std::vector<uint32_t> wgt((size_t)new_w * Bpp);
            for (int b = 0; b < scale_w; b++)
                for (int x = 0; x < dst_block; x++)
                    for (int k = 0; k < Bpp; k++) {
                        const size_t o = ((size_t)b * dst_block + x) * Bpp + k;
                        size_t so = ((size_t)b * src_block + ts[x]) * Bpp + k;
                        uint32_t wp;
                        if (w1[x] >= 0x10000u) wp = 0x0000FFFFu;                          // tap 1 untouched
                        else if (w2[x] >= 0x10000u) { wp = 0x0000FFFFu; so += Bpp; }      // w1 == 0: tap 2 untouched
                        else wp = (w1[x] & 0xFFFFu) | (w2[x] << 16);
                        off[o] = (uint16_t)so;
                        wgt[o] = wp;
                    }
            const uint32_t *dwgt = static_cast<const uint32_t *>(device_blob(c, wgt.data(), wgt.size() * sizeof(uint32_t), st));
            if (!dwgt) return 0;
            // window form: the four first taps of every output word within 8 source bytes (any ratio up to ~2:1)
            std::vector<uint32_t> meta(off.size() / 4);
            bool windowed = tls.force_tier != 1;
            for (size_t ow = 0; ow < meta.size() && windowed; ow++) {
                uint32_t lo = off[4 * ow], hi = lo;
                for (int j = 1; j < 4; j++) { lo = std::min<uint32_t>(lo, off[4 * ow + j]); hi = std::max<uint32_t>(hi, off[4 * ow + j]); }
                if (hi - lo > 7) { windowed = false; break; }
                uint32_t sel = 0;
                for (int j = 0; j < 4; j++) sel |= (off[4 * ow + j] - lo) << (4 * j);
                meta[ow] = lo | (sel << 16);
            }
            if (windowed) {
                const uint32_t *dmeta = static_cast<const uint32_t *>(device_blob(c, meta.data(), meta.size() * sizeof(uint32_t), st));
                if (!dmeta) return 0;
                for (int f0 = 0; f0 < nframes; f0 += 32768) {
                    const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
                    if (!resize_h_win_launch(src + (size_t)f0 * sp_, sp_, dest + (size_t)f0 * dp_, dp_, dmeta, dwgt, width, new_w,
                                             new_h, Bpp, nf, st))
                        return 0;
                }
                return 1;
            }
            const uint16_t *doff = static_cast<const uint16_t *>(device_blob(c, off.data(), off.size() * sizeof(uint16_t), st));
            if (!doff) return 0;
            for (int f0 = 0; f0 < nframes; f0 += 32768) {
                const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
                if (!resize_h_row_launch(src + (size_t)f0 * sp_, sp_, dest + (size_t)f0 * dp_, dp_, doff, dwgt, width, new_w,
                                         new_h, Bpp, nf, st))
                    return 0;
            }
            return 1;
        }
        const int32_t *dts = static_cast<const int32_t *>(device_blob(c, ts.data(), sizeof(int32_t) * n, st));
        const uint32_t *dw1 = static_cast<const uint32_t *>(device_blob(c, w1.data(), sizeof(uint32_t) * n, st));
        const uint32_t *dw2 = static_cast<const uint32_t *>(device_blob(c, w2.data(), sizeof(uint32_t) * n, st));
        if (!dts || !dw1 || !dw2) return 0;
        for (int f0 = 0; f0 < nframes; f0 += 32768) {
            const int nf = nframes - f0 < 32768 ? nframes - f0 : 32768;
            if (!resize_h_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, dts, dw1, dw2,
                                 width, new_w, new_h, Bpp, scale_w, nf, st))
                return 0;
        }
        return 1;
    }
    // no resize requested: the reference leaves dest untouched (tcvideo.c:459,481)
    return 1;
}

int acgpu_convert_batch(uint8_t *src, uint8_t *dest, int width, int height, ImageFormat srcfmt, ImageFormat destfmt,
                        size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    // libtcvideo/tcvideo.c:1001-1067
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!src || !dest || width <= 0 || height <= 0 || !srcfmt || !destfmt) { set_error("acgpu_convert_batch: invalid image parameters"); return 0; }
    const int sf = srcfmt == IMG_YV12 ? IMG_YUV420P : (int)srcfmt, df = destfmt == IMG_YV12 ? IMG_YUV420P : (int)destfmt;
    if (describe(sf).kind == K_NONE || describe(df).kind == K_NONE) { set_error("acgpu_convert_batch: unknown format"); return 0; }
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    const size_t sfb = frame_bytes(sf, width, height), dfb = frame_bytes(df, width, height);
    if (srcfmt == destfmt) {
        if (src == dest) return 1;
        return check(cudaMemcpy2DAsync(dest, dpitch ? dpitch : dfb, src, spitch ? spitch : sfb, dfb, nframes,
                                       cudaMemcpyDeviceToDevice, st), "acgpu_convert_batch copy") ? 1 : 0;
    }
    uint8_t *real = dest;
    size_t rpitch = dpitch;
    if (src == dest) {                    // in place: convert into a temporary, then copy back (tcvideo.c:1044-1064)
        rpitch = align_up(dfb, 256);
        if (!ensure_arena(c, rpitch * (size_t)nframes) || !arena_acquire(c, st)) return 0;
        real = c->arena;
    }
    uint8_t *sp[3], *dp[3];
    sp[0] = src;  sp[1] = src + (size_t)width * height;  sp[2] = sp[1] + chroma_plane_bytes(sf, width, height);
    dp[0] = real; dp[1] = real + (size_t)width * height; dp[2] = dp[1] + chroma_plane_bytes(df, width, height);
    if (!acgpu_imgconvert_batch(sp, srcfmt, spitch, dp, destfmt, rpitch, width, height, nframes, stream)) return 0;
    if (src == dest)
        return check(cudaMemcpy2DAsync(dest, dpitch ? dpitch : dfb, real, rpitch, dfb, nframes, cudaMemcpyDeviceToDevice, st),
                     "acgpu_convert_batch copy back") && arena_release(c, st) ? 1 : 0;
    return 1;
}

int acgpu_decolor_rgb24_batch(uint8_t *frames, int width, int height, size_t pitch, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (!frames || width <= 0 || height <= 0) { set_error("acgpu_decolor_rgb24_batch: invalid frame parameters"); return 0; }
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    if (tls.force_tier != 1 && per_frame_chunk(nframes, [&](int f0, int nf) {
            return decolor_rgb24_fast(frames + (size_t)f0 * pitch, pitch, width, height, nf, st); })) { tls.last_tier = 2; return 1; }
    // outside the vectorised domain: the reference's own two steps through a temporary gray plane
    const size_t gpitch = align_up((size_t)width * height, 256);
    if (!ensure_arena(c, gpitch * (size_t)nframes) || !arena_acquire(c, st)) return 0;
    uint8_t *rgb[3] = {frames, nullptr, nullptr}, *gray[3] = {c->arena, nullptr, nullptr};
    return acgpu_imgconvert_batch(rgb, IMG_RGB24, pitch, gray, IMG_GRAY8, gpitch, width, height, nframes, stream)
        && acgpu_imgconvert_batch(gray, IMG_GRAY8, gpitch, rgb, IMG_RGB24, pitch, width, height, nframes, stream)
        && arena_release(c, st);
}

// ---- the remaining element-wise libtcvideo operations ------------------------------------
static bool plane_args_ok(const char *who, const void *src, const void *dest, int width, int height, int Bpp)
{
    // the check every tcv_* function opens with (e.g. libtcvideo/tcvideo.c:192-195)
    if (!src || !dest || width <= 0 || height <= 0 || (Bpp != 1 && Bpp != 3)) { set_error("%s: invalid frame parameters", who); return false; }
    if ((uint64_t)width * height * Bpp >= 0x7FFFFFF0ull) { set_error("%s: plane too large", who); return false; }
    return true;
}

int acgpu_clip_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int clip_left, int clip_right,
                     int clip_top, int clip_bottom, uint8_t black_pixel, size_t spitch, size_t dpitch, int nframes,
                     acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_clip_batch", src, dest, width, height, Bpp)) return 0;
    if ((int64_t)clip_left + clip_right >= width || (int64_t)clip_top + clip_bottom >= height) {   // tcvideo.c:196-202
        set_error("acgpu_clip_batch: clipping parameters (%d,%d,%d,%d) invalid for frame size %dx%d", clip_top, clip_left,
                  clip_bottom, clip_right, width, height);
        return 0;
    }
    // tcvideo.c:204-219: a clip wider than the frame eats into the opposite (negative) border
    if (clip_left > width)    { clip_right += clip_left - width;    clip_left = width; }
    if (clip_right > width)   { clip_left += clip_right - width;    clip_right = width; }
    if (clip_top > height)    { clip_bottom += clip_top - height;   clip_top = height; }
    if (clip_bottom > height) { clip_top += clip_bottom - height;   clip_bottom = height; }
    const int64_t new_w = (int64_t)width - clip_left - clip_right, new_h = (int64_t)height - clip_top - clip_bottom;
    const int64_t copy_w = (int64_t)width - (clip_left < 0 ? 0 : clip_left) - (clip_right < 0 ? 0 : clip_right);
    if (new_w * new_h * Bpp >= 0x7FFFFFF0ll) { set_error("acgpu_clip_batch: result too large"); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    TcvWindow p{};
    p.src = src; p.spitch = spitch; p.dst = dest; p.dpitch = dpitch;
    p.dBpl = (uint32_t)(new_w * Bpp); p.sBpl = (uint32_t)width * Bpp;
    p.drows = (int)new_h; p.srows = height;
    p.row_mul = 1; p.row_add = clip_top;
    p.cl = (uint32_t)((clip_left < 0 ? -(int64_t)clip_left : 0) * Bpp);
    p.cn = (uint32_t)((copy_w > 0 ? copy_w : 0) * Bpp);
    p.sxb = (uint32_t)((clip_left > 0 ? clip_left : 0) * Bpp);
    p.fill = 0x01010101u * black_pixel;
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        TcvWindow q = p;
        q.src += (size_t)f0 * spitch; q.dst += (size_t)f0 * dpitch;
        return tcv_window_launch(q, nf, st);
    });
}

int acgpu_reduce_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, int reduce_w, int reduce_h,
                       size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_reduce_batch", src, dest, width, height, Bpp)) return 0;
    if (reduce_w <= 0 || reduce_h <= 0) { set_error("acgpu_reduce_batch: invalid reduction parameters (%d,%d)", reduce_w, reduce_h); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    if (reduce_w != 1)      // tcvideo.c:694-704
        return per_frame_chunk(nframes, [&](int f0, int nf) {
            return tcv_reduce_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, width / reduce_w,
                                     height / reduce_h, reduce_w, reduce_h, Bpp, nf, st);
        });
    // :706-715 whole rows: every reduce_h-th one, or (reduce_h == 1) the plain copy
    TcvWindow p{};
    p.src = src; p.spitch = spitch; p.dst = dest; p.dpitch = dpitch;
    p.dBpl = p.sBpl = (uint32_t)width * Bpp;
    p.drows = height / reduce_h; p.srows = height;
    p.row_mul = reduce_h; p.row_add = 0;
    p.cl = 0; p.cn = p.dBpl; p.sxb = 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        TcvWindow q = p;
        q.src += (size_t)f0 * spitch; q.dst += (size_t)f0 * dpitch;
        return tcv_window_launch(q, nf, st);
    });
}

int acgpu_flip_v_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, size_t spitch, size_t dpitch,
                       int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_flip_v_batch", src, dest, width, height, Bpp)) return 0;
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_flip_v_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, height, Bpp, nf, st);
    });
}

int acgpu_flip_h_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, size_t spitch, size_t dpitch,
                       int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_flip_h_batch", src, dest, width, height, Bpp)) return 0;
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_flip_h_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, width, height, Bpp, nf, st);
    });
}

int acgpu_gamma_correct_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double gamma,
                              size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_gamma_correct_batch", src, dest, width, height, Bpp)) return 0;
    if (!(gamma > 0)) { set_error("acgpu_gamma_correct_batch: invalid gamma (%.3f)", gamma); return 0; }   // tcvideo.c:848-851
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    uint8_t table[256];
    for (int i = 0; i < 256; i++) table[i] = (uint8_t)(pow((i / 255.0), gamma) * 255);    // tcvideo.c:1180-1189, host doubles
    const uint8_t *d_table = static_cast<const uint8_t *>(device_blob(c, table, sizeof(table), st));
    if (!d_table) return 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_lut_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_table, (size_t)width * height * Bpp, nf, st);
    });
}

int acgpu_antialias_batch(const uint8_t *src, uint8_t *dest, int width, int height, int Bpp, double weight, double bias,
                          size_t spitch, size_t dpitch, int nframes, acgpu_stream_t stream)
{
    if (!plane_args_ok("acgpu_antialias_batch", src, dest, width, height, Bpp)) return 0;
    if (!(weight >= 0 && weight <= 1 && bias >= 0 && bias <= 1)) {                       // tcvideo.c:899-903
        set_error("acgpu_antialias_batch: invalid antialiasing parameters (weight=%.3f, bias=%.3f)", weight, bias);
        return 0;
    }
    if (src == dest) { set_error("acgpu_antialias_batch: src and dest must not overlap (a 3x3 neighbourhood is read)"); return 0; }
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nframes <= 0) return 1;
    cudaStream_t st = pick_stream(c, stream);
    uint32_t t[1024];     // c | x | y | d, tcvideo.c:1209-1224 (double -> uint32 truncation, evaluation order kept)
    for (int i = 0; i < 256; i++) {
        t[i] = i * weight * 65536;
        t[256 + i] = i * bias * (1 - weight) / 4 * 65536;
        t[512 + i] = i * (1 - bias) * (1 - weight) / 4 * 65536;
        t[768 + i] = (t[256 + i] + t[512 + i] + 1) / 2;
    }
    const uint32_t *d_t = static_cast<const uint32_t *>(device_blob(c, t, sizeof(t), st));
    if (!d_t) return 0;
    return per_frame_chunk(nframes, [&](int f0, int nf) {
        return tcv_antialias_launch(src + (size_t)f0 * spitch, spitch, dest + (size_t)f0 * dpitch, dpitch, d_t, width, height, Bpp, nf, st);
    });
}

}  // extern "C"
