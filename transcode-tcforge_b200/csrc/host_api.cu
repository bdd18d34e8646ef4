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
#include "host_ctx.h"

#include <math.h>

#include <algorithm>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace acgpu {

thread_local ThreadCtx tls;

namespace {

std::atomic<int> g_initialised{0};
std::atomic<int> g_sm_count{0};
// Set by the first acgpu_* call that can put device memory into a caller's hands (acgpu_malloc, acgpu_set_device,
// acgpu_stream_create, any batched entry point).  Until then every pointer handed to ac_memcpy is host memory.
std::atomic<bool> g_cuda_aware{false};
// Set at process exit: thread-local and static destructors then leave CUDA alone (the runtime may already be gone).
std::atomic<bool> g_exiting{false};
struct ExitHook { ExitHook() { atexit([] { g_exiting.store(true); }); } } g_exit_hook;

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

}  // namespace

bool process_exiting() { return g_exiting.load(); }

int pipe_slots()
{
    static const int n = [] {
        const char *e = getenv("ACGPU_PIPE_SLOTS");
        const int v = e ? atoi(e) : 4;
        return v < 2 ? 2 : v > kPipeSlots ? kPipeSlots : v;
    }();
    return n;
}

size_t pipe_chunk_bytes()
{
    static const size_t n = [] {
        const char *e = getenv("ACGPU_PIPE_CHUNK_MB");
        const int v = e ? atoi(e) : 32;
        return (size_t)(v < 1 ? 1 : v > 256 ? 256 : v) << 20;
    }();
    return n;
}

DevCtx *ctx()
{
    if (!bind_device()) return nullptr;
    DevCtx *c = &tls.dev[cur_device()];
    if (!c->stream && !check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        return nullptr;
    return c;
}

cudaStream_t pick_stream(DevCtx *c, acgpu_stream_t s)
{
    g_cuda_aware.store(true, std::memory_order_relaxed);     // only acgpu_* entry points name streams
    return s ? reinterpret_cast<cudaStream_t>(s) : c->stream;
}

// Staged (host-pointer) work runs on the thread's private stream.  When the caller named another stream, device data it
// hands over may still be in flight there: the private stream is ordered after everything queued on the caller's
// stream so far.
bool order_after(DevCtx *c, acgpu_stream_t caller)
{
    cudaStream_t cs = reinterpret_cast<cudaStream_t>(caller);
    if (!cs || cs == c->stream) return true;
    if (!c->order_ev && !check(cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming), "cudaEventCreate")) return false;
    return check(cudaEventRecord(c->order_ev, cs), "order record") && check(cudaStreamWaitEvent(c->stream, c->order_ev, 0), "order wait");
}

namespace {

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

}  // namespace

bool is_device_pointer(const void *p) { return classify(p) == PK_DEVICE; }

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

int pointer_kind(const void *p) { return (int)classify(p); }

namespace {

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


// ---- parallel host copies ---------------------------------------------------------------------------------------------
// A lone caller on pageable frames is bound by its own memcpy into / out of pinned memory (~10 GB/s on one core; the link
// moves 55).  A few helper threads share those copies.  Process-wide, created on first use, never joined (see g_workers).
class CopyPool {
    struct Job { uint8_t *d; const uint8_t *s; size_t n; std::atomic<int> *left; };
    std::mutex m;
    std::condition_variable cv;
    std::vector<Job> q;
    std::atomic<int> pending{0};
    int nthreads = 0;

    bool try_pop(Job *j)
    {
        if (pending.load(std::memory_order_acquire) <= 0) return false;
        std::lock_guard<std::mutex> lk(m);
        if (q.empty()) return false;
        *j = q.back();
        q.pop_back();
        pending.fetch_sub(1, std::memory_order_relaxed);
        return true;
    }
    void worker()
    {
        for (;;) {
            Job j{};
            bool have = false;
            // chunks of one frame follow each other within microseconds: poll for a while before going to sleep (a
            // condition-variable wake-up costs more than the 256 KB copy it announces)
            for (int spin = 0; spin < 4000 && !have; spin++) {
                have = try_pop(&j);
                if (!have) __builtin_ia32_pause();
            }
            if (!have) {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return !q.empty(); });
                j = q.back();
                q.pop_back();
                pending.fetch_sub(1, std::memory_order_relaxed);
            }
            memcpy(j.d, j.s, j.n);
            j.left->fetch_sub(1, std::memory_order_release);
        }
    }

public:
    static CopyPool &get()
    {
        static CopyPool *pool = [] {
            CopyPool *p = new CopyPool();      // leaked on purpose: its threads must not be joined at process exit
            const char *e = getenv("ACGPU_COPY_THREADS");
            int n = e ? atoi(e) : 3;
            const int cores = (int)std::thread::hardware_concurrency();
            if (n > cores - 1) n = cores - 1;
            if (n < 0) n = 0;
            p->nthreads = n;
            for (int i = 0; i < n; i++) std::thread([p] { p->worker(); }).detach();
            return p;
        }();
        return *pool;
    }
    // copies n bytes with the helpers' aid; the caller takes the last piece itself and returns when all are done
    void copy(uint8_t *d, const uint8_t *s, size_t n, bool parallel)
    {
        // smallest piece worth handing to a helper ($ACGPU_COPY_PIECE_KB).  64 KB: a 1080p chroma plane (518 KB) and the last,
        // partial slot of a plane are shared out too -- with 256 KB they were copied by the caller alone and a lone caller on
        // pageable 1080p frames made 1.9-2.5 k conversions/s instead of 3.0 k (profiles/r2_experiments.md section 5)
        static const size_t kPiece = [] { const char *e = getenv("ACGPU_COPY_PIECE_KB"); const int v = e ? atoi(e) : 64; return (size_t)(v < 16 ? 16 : v) << 10; }();
        if (!parallel || nthreads == 0 || n < 2 * kPiece) { memcpy(d, s, n); return; }
        const int pieces = (int)std::min<size_t>((size_t)nthreads + 1, n / kPiece);
        const size_t each = (n / (size_t)pieces + 63) & ~(size_t)63;
        std::atomic<int> left{pieces - 1};
        {
            std::lock_guard<std::mutex> lk(m);
            for (int i = 0; i + 1 < pieces; i++) q.push_back(Job{d + (size_t)i * each, s + (size_t)i * each, each, &left});
            pending.fetch_add(pieces - 1, std::memory_order_release);
        }
        cv.notify_all();
        const size_t done = (size_t)(pieces - 1) * each;
        memcpy(d + done, s + done, n - done);
        while (left.load(std::memory_order_acquire) > 0) {
            // help with whatever is still queued (another caller's pieces count too) instead of spinning idle
            Job j{};
            if (try_pop(&j)) { memcpy(j.d, j.s, j.n); j.left->fetch_sub(1, std::memory_order_release); }
        }
    }
};

bool ensure_ring(DevCtx *c)
{
    if (c->ring) return true;
    if (!check(cudaHostAlloc(&c->ring, DevCtx::kRingSlots * DevCtx::kRingSlotBytes, cudaHostAllocDefault), "cudaHostAlloc(ring)")) return false;
    for (cudaEvent_t &e : c->ring_ev)
        if (!check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return false;
    return true;
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
    // Automatic order: the tensor-map staged form of tier 3 where it is measured faster (YUV420P -> RGB on wide frames),
    // the vectorised tier, else generic.  The other tier-3 variants (bulk / tensor stores) are selectable only: they
    // measured slower than tier 2 (profiles/r1_experiments.md, profiles/r2_tma_tensor_maps.md).
    const int force = tls.force_tier;
    tls.err[0] = 0;      // a tier that declines a call leaves this empty; a failed launch leaves its CUDA error
    if (force == 3) {
        if (convert_tma(a)) { tls.last_tier = 3; return true; }
        if (!tls.err[0]) set_error("tier 3 (bulk stores) does not cover this pair/size/alignment");
        return false;
    }
    if (force == 0 && convert_tma_auto(a)) { tls.last_tier = 3; return true; }
    if (tls.err[0]) return false;
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

}  // namespace

StagedCall::StagedCall() : others(g_staged_calls.fetch_add(1)) {}
StagedCall::~StagedCall() { g_staged_calls.fetch_sub(1); }

bool staged_h2d(DevCtx *c, uint8_t *d, size_t dpitch, const uint8_t *h, size_t hpitch, size_t width, size_t rows, cudaStream_t st,
                int host_kind, bool lone_caller)
{
    if (!width || !rows) return true;
    if (host_kind != PK_HOST)       // page-locked: the copy engine reads the caller's memory directly
        return check(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width, rows, cudaMemcpyHostToDevice, st), "H2D");
    if (!ensure_ring(c)) return false;
    CopyPool &pool = CopyPool::get();
    for (size_t r = 0; r < rows; r++)
        for (size_t off = 0; off < width; off += DevCtx::kRingSlotBytes) {
            const size_t n = std::min(DevCtx::kRingSlotBytes, width - off);
            const int slot = c->ring_next;
            c->ring_next = (slot + 1) % DevCtx::kRingSlots;
            uint8_t *b = c->ring + (size_t)slot * DevCtx::kRingSlotBytes;
            if (!check(cudaEventSynchronize(c->ring_ev[slot]), "ring wait")) return false;       // its previous DMA has drained
            pool.copy(b, h + r * hpitch + off, n, lone_caller);
            if (!check(cudaMemcpyAsync(d + r * dpitch + off, b, n, cudaMemcpyHostToDevice, st), "H2D")
                || !check(cudaEventRecord(c->ring_ev[slot], st), "ring record"))
                return false;
        }
    return true;
}

bool staged_d2h(DevCtx *c, uint8_t *h, size_t hpitch, const uint8_t *d, size_t dpitch, size_t width, size_t rows, cudaStream_t st,
                int host_kind, bool lone_caller)
{
    if (!width || !rows) return true;
    if (host_kind != PK_HOST)
        return check(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, rows, cudaMemcpyDeviceToHost, st), "D2H");
    if (!ensure_ring(c)) return false;
    CopyPool &pool = CopyPool::get();
    // chunk list in order; DMAs run kRingSlots - 1 chunks ahead of the host copies
    struct Chunk { size_t r, off, n; int slot; };
    const size_t per_row = (width + DevCtx::kRingSlotBytes - 1) / DevCtx::kRingSlotBytes, total = per_row * rows;
    auto chunk_at = [&](size_t i) {
        Chunk k;
        k.r = i / per_row;
        k.off = (i - k.r * per_row) * DevCtx::kRingSlotBytes;
        k.n = std::min(DevCtx::kRingSlotBytes, width - k.off);
        k.slot = (int)((c->ring_next + i) % DevCtx::kRingSlots);
        return k;
    };
    auto issue = [&](size_t i) {
        const Chunk k = chunk_at(i);
        // the slot's previous use was a host copy that has completed (d2h) or an upload whose event is waited here
        if (!check(cudaEventSynchronize(c->ring_ev[k.slot]), "ring wait")) return false;
        return check(cudaMemcpyAsync(c->ring + (size_t)k.slot * DevCtx::kRingSlotBytes, d + k.r * dpitch + k.off, k.n, cudaMemcpyDeviceToHost, st), "D2H")
            && check(cudaEventRecord(c->ring_ev[k.slot], st), "ring record");
    };
    const size_t ahead = DevCtx::kRingSlots - 1;
    for (size_t i = 0; i < total && i < ahead; i++)
        if (!issue(i)) return false;
    for (size_t i = 0; i < total; i++) {
        const Chunk k = chunk_at(i);
        if (!check(cudaEventSynchronize(c->ring_ev[k.slot]), "D2H wait")) return false;
        pool.copy(h + k.r * hpitch + k.off, c->ring + (size_t)k.slot * DevCtx::kRingSlotBytes, k.n, lone_caller);
        if (i + ahead < total && !issue(i + ahead)) return false;
    }
    c->ring_next = (int)((c->ring_next + total) % DevCtx::kRingSlots);
    return true;
}

namespace {

// One frame through ac_imgconvert's legacy signature; planes may live on the host or on the device.
bool convert_one(Image si, int sfmt, Image di, int dfmt, int w, int h)
{
    DevCtx *c = ctx();
    if (!c) return false;
    size_t ssz[3], dsz[3];
    int snp, dnp;
    plane_sizes(sfmt, w, h, ssz, &snp);
    plane_sizes(dfmt, w, h, dsz, &dnp);
    // one driver query per plane: every plane of an image must live on the same side (a host/device mix would be staged
    // wrongly), and the pageable / page-locked distinction decides whether the bounce buffer is used
    PtrKind sk = classify(si.p[0]), dk = classify(di.p[0]);
    for (int p = 1; p < snp; p++) {
        const PtrKind k = classify(si.p[p]);
        if ((k == PK_DEVICE) != (sk == PK_DEVICE)) { set_error("ac_imgconvert: source planes mix host and device memory"); return false; }
        if (k == PK_HOST) sk = sk == PK_DEVICE ? sk : PK_HOST;      // one pageable plane makes the image pageable
    }
    for (int p = 1; p < dnp; p++) {
        const PtrKind k = classify(di.p[p]);
        if ((k == PK_DEVICE) != (dk == PK_DEVICE)) { set_error("ac_imgconvert: destination planes mix host and device memory"); return false; }
        if (k == PK_HOST) dk = dk == PK_DEVICE ? dk : PK_HOST;
    }
    const bool src_host = sk != PK_DEVICE, dst_host = dk != PK_DEVICE;

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
    // Pageable caller memory travels through the thread's ring of pinned slots (staged_h2d / staged_d2h): a lone caller --
    // the usual transcode run has one frame thread -- gets its host copies spread over the helper threads; concurrent
    // callers (src/frame_threads.c:174-228) each copy their own, pipelined with their DMAs.
    struct Busy {
        bool on;
        int  others;
        explicit Busy(bool o) : on(o), others(o ? g_staged_calls.fetch_add(1) : 0) {}
        ~Busy() { if (on) g_staged_calls.fetch_sub(1); }
    } busy(src_host || dst_host);
    const bool lone = busy.others == 0;
    if (src_host)
        for (int p = 0; p < snp; p++) {
            a.src.p[p] = c->arena + soff[p];
            if (!staged_h2d(c, a.src.p[p], ssz[p], si.p[p], ssz[p], ssz[p], 1, c->stream, sk, lone)) return false;
        }
    if (dst_host) {
        const bool preload = !overwrites_whole_dest(sfmt, dfmt, w, h);
        for (int p = 0; p < dnp; p++) {
            a.dst.p[p] = c->arena + doff[p];
            if (preload && !staged_h2d(c, a.dst.p[p], dsz[p], di.p[p], dsz[p], dsz[p], 1, c->stream, dk, lone)) return false;
        }
    }
    if (!run_convert(a)) return false;
    if (dst_host)
        for (int p = 0; p < dnp; p++)
            if (!staged_d2h(c, di.p[p], dsz[p], a.dst.p[p], dsz[p], dsz[p], 1, c->stream, dk, lone)) return false;
    // callers that also copy through the ring do better spinning at every thread count measured
    return wait_stream(c, (sk == PK_HOST || dk == PK_HOST) ? 0 : busy.others, "ac_imgconvert");
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
}  // namespace

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

namespace {

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
    // libtcvideo calls this once per ROW (tcvideo.c:229-246, 706-715): an unmodified caller -- one that never called an
    // acgpu_* entry point and so cannot hold device memory -- pays no driver query per copy.
    if (!g_cuda_aware.load(std::memory_order_relaxed)) return memmove(dest, src, size);
    if (!bind_device()) fatal("ac_memcpy");                  // the query below would otherwise bind this thread to device 0
    const PtrKind kd = classify(dest), ks = classify(src);
    if (kd != PK_DEVICE && ks != PK_DEVICE) return memmove(dest, src, size);
    DevCtx *c = ctx();
    if (!c) fatal("ac_memcpy");
    const uint8_t *s = static_cast<const uint8_t *>(src);
    uint8_t *d = static_cast<uint8_t *>(dest);
    bool ok;
    if (kd == PK_DEVICE && ks == PK_DEVICE && s < d + size && d < s + size) {
        // overlapping device ranges: memmove semantics through the thread's temporary (a plain device copy is undefined)
        ok = ensure_arena(c, size) && arena_acquire(c, c->stream)
          && check(cudaMemcpyAsync(c->arena, src, size, cudaMemcpyDeviceToDevice, c->stream), "ac_memcpy")
          && check(cudaMemcpyAsync(dest, c->arena, size, cudaMemcpyDeviceToDevice, c->stream), "ac_memcpy");
    } else {
        ok = check(cudaMemcpyAsync(dest, src, size, cudaMemcpyDefault, c->stream), "ac_memcpy");
    }
    if (!ok || !check(cudaStreamSynchronize(c->stream), "ac_memcpy")) fatal("ac_memcpy");
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
    g_cuda_aware.store(true, std::memory_order_relaxed);
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
    g_cuda_aware.store(true, std::memory_order_relaxed);
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

// ---- frame-buffer plumbing -----------------------------------------------------------------------
// transcode allocates every frame buffer through tc_bufalloc (libtcutil/memutils.c:89-112: page-aligned malloc;
// libtc/tcframes.c:214-229 calls it twice per video frame, for vframe_list_t.internal_video_buf_0 / _1).  The same contract
// on page-locked memory makes every legacy ac_* / tcv_* call on those buffers a direct DMA instead of a staged copy.
// Layout, as tc_bufalloc's own (memutils.c:93-106): [base ... | kind | base pointer | page-aligned user buffer].  Small
// cudaHostAlloc blocks are sub-allocated at 256- or 512-byte granularity, so the page alignment is made here.
static const uint64_t kBufPinned = 0x6163677075627566ull, kBufPlain = 0x6163677075626d61ull;

void *acgpu_bufalloc(size_t size)
{
    const long pg = sysconf(_SC_PAGESIZE);
    const size_t page = pg > 0 ? (size_t)pg : 4096, total = (size ? size : 1) + page + 16;
    void *base = nullptr;
    uint64_t kind = kBufPinned;
    if (!(device_usable(nullptr) && bind_device() && cudaHostAlloc(&base, total, cudaHostAllocPortable) == cudaSuccess)) {
        cudaGetLastError();
        // no device (the library will refuse to convert anyway): still hand out what tc_bufalloc promises, a page-aligned buffer
        base = malloc(total);
        kind = kBufPlain;
        if (!base) return nullptr;
    }
    uint8_t *user = reinterpret_cast<uint8_t *>(align_up(reinterpret_cast<uintptr_t>(base) + 16, page));
    reinterpret_cast<void **>(user)[-1] = base;
    reinterpret_cast<uint64_t *>(user)[-2] = kind;
    return user;
}

void acgpu_buffree(void *ptr)
{
    if (!ptr) return;
    void *base = reinterpret_cast<void **>(ptr)[-1];
    const uint64_t kind = reinterpret_cast<uint64_t *>(ptr)[-2];
    if (kind == kBufPinned) cudaFreeHost(base);
    else if (kind == kBufPlain) free(base);
    else fprintf(stderr, "libacgpu: acgpu_buffree(%p): not a buffer from acgpu_bufalloc\n", ptr);
}

// An existing buffer (one tc_bufalloc already returned, a decoder's own frame): page-lock it ONCE, when it is created --
// not per call, which costs more than the copy it saves.  Whole pages are locked (the range is widened to page bounds).
int acgpu_host_register(void *ptr, size_t size)
{
    if (!ptr || !size) { set_error("acgpu_host_register: empty range"); return 0; }
    if (!bind_device()) return 0;
    const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE), a = (uintptr_t)ptr & ~(page - 1), e = ((uintptr_t)ptr + size + page - 1) & ~(page - 1);
    return check(cudaHostRegister(reinterpret_cast<void *>(a), e - a, cudaHostRegisterPortable), "acgpu_host_register") ? 1 : 0;
}

int acgpu_host_unregister(void *ptr)
{
    if (!ptr) return 0;
    const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
    return check(cudaHostUnregister(reinterpret_cast<void *>((uintptr_t)ptr & ~(page - 1))), "acgpu_host_unregister") ? 1 : 0;
}

// 0 = pageable host memory, 1 = page-locked host memory, 2 = device memory: which path a legacy call on `ptr` takes.
int acgpu_pointer_kind(const void *ptr) { return bind_device() ? (int)classify(ptr) : 0; }

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
    g_cuda_aware.store(true, std::memory_order_relaxed);
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
    if (nframes > 1 && (src_frame_pitch == 0 || dest_frame_pitch == 0)) {
        // planes are separate pointers here, so "tightly packed" has no single meaning: a batch needs explicit pitches
        set_error("acgpu_imgconvert_batch: frame pitch 0 with %d frames", nframes);
        return 0;
    }
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
    // ~32 MiB of the larger side per chunk in four slots: big enough to amortise launches and reach full PCIe rate, small
    // enough that the fill / drain bubbles of the pipeline stay short (profiles/r2_experiments.md section 6: slots x chunk)
    size_t per = pipe_chunk_bytes() / (sp > dp ? sp : dp);
    if (per < 1) per = 1;
    if (per > (size_t)nframes) per = nframes;
    const size_t slot_bytes = per * (sp + dp);
    for (int s = 0; s < pipe_slots(); s++) {
        if (!c->pipe_stream[s] && !check(cudaStreamCreateWithFlags(&c->pipe_stream[s], cudaStreamNonBlocking), "pipe stream")) return 0;
        if (c->pipe_cap[s] < slot_bytes) {
            if (c->pipe_buf[s]) { cudaStreamSynchronize(c->pipe_stream[s]); cudaFree(c->pipe_buf[s]); c->pipe_buf[s] = nullptr; c->pipe_cap[s] = 0; }
            if (!check(cudaMalloc(&c->pipe_buf[s], slot_bytes), "cudaMalloc(pipeline slot)")) return 0;
            c->pipe_cap[s] = slot_bytes;
        }
    }
    int chunk = 0;
    for (int f0 = 0; f0 < nframes; f0 += (int)per, chunk++) {
        const int s = chunk % pipe_slots();
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
    for (int s = 0; s < pipe_slots(); s++)
        if (!check(cudaStreamSynchronize(c->pipe_stream[s]), "acgpu_imgconvert_frames_host")) return 0;
    return 1;
}

int acgpu_imgconvert_frames_host_multi(const uint8_t *src_frames, ImageFormat srcfmt, uint8_t *dest_frames,
                                       ImageFormat destfmt, int width, int height, int nframes, int ndevices)
{
    const int sf = srcfmt == IMG_YV12 ? IMG_YUV420P : (int)srcfmt, df = destfmt == IMG_YV12 ? IMG_YUV420P : (int)destfmt;
    if (describe(sf).kind == K_NONE || describe(df).kind == K_NONE) { set_error("unknown format pair"); return 0; }
    if (width <= 0 || height <= 0 || nframes <= 0) return 1;
    const size_t sfb = frame_bytes(sf, width, height), dfb = frame_bytes(df, width, height);
    return run_on_devices("acgpu_imgconvert_frames_host_multi", ndevices, nframes, [=](int, int f0, int f1) {
        return acgpu_imgconvert_frames_host(src_frames + (size_t)f0 * sfb, srcfmt, dest_frames + (size_t)f0 * dfb, destfmt,
                                            width, height, f1 - f0) == 1;
    });
}

void acgpu_shutdown(void)
{
    // Stops the per-device host threads of the *_multi calls (each gives its streams and buffers back as it exits) while
    // the CUDA runtime is still alive.  Optional: without it they are simply abandoned at process exit.
    stop_device_workers();
}

// ---- row operations -----------------------------------------------------------------------------------
int acgpu_rowops_run(const uint8_t *src, size_t spitch, uint8_t *dest, size_t dpitch, const acgpu_rowop *ops,
                     int nops, int row_bytes, int nframes, acgpu_stream_t stream)
{
    DevCtx *c = ctx();
    if (!c) return 0;
    if (nops <= 0 || row_bytes <= 0 || nframes <= 0) return 1;
    if (!ops) { set_error("acgpu_rowops_run: null op list"); return 0; }
    if (nframes > 1 && (spitch == 0 || dpitch == 0)) { set_error("acgpu_rowops_run: frame pitch 0 with %d frames", nframes); return 0; }
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

}  // extern "C"
