// host_ctx.h -- host-side state shared by libacgpu's two host translation units (not installed):
// host_api.cu (aclib entry points, device/memory helpers, batched conversion, row operations) and
// host_tcv.cu (the frame-granular libtcvideo entry points).  Everything mutable is per caller thread: callers are N
// concurrent frame threads (src/frame_threads.c:174-228).
#pragma once

#include "acgpu_internal.h"

#include <functional>
#include <vector>

namespace acgpu {

constexpr int kMaxDev = 16;
constexpr int kPipeSlots = 6;          // buffers of the host-frame pipelines; pipe_slots() of them are used
int pipe_slots();                      // $ACGPU_PIPE_SLOTS, default 4
size_t pipe_chunk_bytes();             // $ACGPU_PIPE_CHUNK_MB, default 32: bytes of the larger side per pipeline chunk

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
    cudaEvent_t order_ev = nullptr;      // orders the private stream after a caller-named stream (see order_after)
    uint8_t *plane_stage = nullptr;      // device copies of HOST planes handed to the frame-granular entry points (its own
    size_t   plane_stage_cap = 0;        // buffer: those entry points may use the arena as their temporary meanwhile)
    // pageable caller memory travels through a ring of pinned slots: chunk k+1 is copied by the host while chunk k is DMA'd
    static constexpr int kRingSlots = 4;
    static constexpr size_t kRingSlotBytes = (size_t)1 << 20;
    uint8_t *ring = nullptr;
    cudaEvent_t ring_ev[kRingSlots] = {nullptr, nullptr, nullptr, nullptr};   // slot's last DMA (either direction) is done
    int      ring_next = 0;
    std::vector<Blob> blobs;
    cudaStream_t pipe_stream[kPipeSlots] = {};
    uint8_t *pipe_buf[kPipeSlots] = {};
    size_t   pipe_cap[kPipeSlots] = {};
};

bool process_exiting();      // true once exit() has begun: destructors then leave CUDA alone

struct ThreadCtx {
    int      device = -1;
    DevCtx   dev[kMaxDev];
    char     err[512] = {0};
    uint64_t launches = 0;
    int      last_tier = 0;
    int      force_tier = 0;
    int      device_only = 0;    // > 0 while a chain runs: every pointer handed to the entry points is device memory

    // A caller thread that exits gives its stream, staging buffers and cached tables back.  Best effort: at process
    // exit the CUDA runtime may already be gone, in which case these calls fail harmlessly.
    ~ThreadCtx()
    {
        if (process_exiting()) return;
        for (int d = 0; d < kMaxDev; d++) {
            DevCtx &c = dev[d];
            if (!c.stream && !c.arena && !c.plane_stage && !c.ring && c.blobs.empty() && !c.pipe_stream[0]) continue;
            if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaDeviceSynchronize();     // batched calls may have run on caller-supplied streams that still read our tables
            for (int s = 0; s < kPipeSlots; s++) {
                if (c.pipe_stream[s]) { cudaStreamSynchronize(c.pipe_stream[s]); cudaStreamDestroy(c.pipe_stream[s]); }
                if (c.pipe_buf[s]) cudaFree(c.pipe_buf[s]);
            }
            for (Blob &b : c.blobs) cudaFree(b.dptr);
            if (c.arena) cudaFree(c.arena);
            if (c.plane_stage) cudaFree(c.plane_stage);
            if (c.ring) cudaFreeHost(c.ring);
            for (cudaEvent_t e : c.ring_ev) if (e) cudaEventDestroy(e);
            if (c.sleep_ev) cudaEventDestroy(c.sleep_ev);
            if (c.arena_ev) cudaEventDestroy(c.arena_ev);
            if (c.order_ev) cudaEventDestroy(c.order_ev);
            if (c.stream) cudaStreamDestroy(c.stream);
            cudaGetLastError();
        }
    }
};

extern thread_local ThreadCtx tls;

DevCtx      *ctx();                                            // binds the thread's device, creates its stream on first use
cudaStream_t pick_stream(DevCtx *c, acgpu_stream_t s);         // the caller's stream, or the thread's own
bool  order_after(DevCtx *c, acgpu_stream_t caller);           // private stream waits for what the caller's stream holds so far
bool  is_device_pointer(const void *p);                        // device / managed memory (else pageable or page-locked host)
int   pointer_kind(const void *p);                             // 0 pageable host, 1 page-locked host, 2 device
// Host <-> device transfers of the staged (host-pointer) calls on stream `st`.  Page-locked memory is DMA'd directly and
// asynchronously.  Pageable memory goes through the thread's ring of pinned slots in 1 MB chunks -- the host copy of one
// chunk overlaps the DMA of the previous one -- and, when the caller is the only thread inside a staged call, the host
// copies are spread over a small pool of helper threads (one core copies ~10 GB/s, PCIe moves 55).  `rows` chunks of
// `width` bytes, `hpitch` / `dpitch` apart.  d2h with a pageable destination returns after the data has landed.
bool  staged_h2d(DevCtx *c, uint8_t *d, size_t dpitch, const uint8_t *h, size_t hpitch, size_t width, size_t rows, cudaStream_t st,
                 int host_kind, bool lone_caller);
bool  staged_d2h(DevCtx *c, uint8_t *h, size_t hpitch, const uint8_t *d, size_t dpitch, size_t width, size_t rows, cudaStream_t st,
                 int host_kind, bool lone_caller);
// callers inside staged calls right now (this one included after enter): RAII
struct StagedCall {
    int others;
    StagedCall();
    ~StagedCall();
};
bool  ensure_arena(DevCtx *c, size_t bytes);
bool  arena_acquire(DevCtx *c, cudaStream_t st);               // orders uses of the arena on different streams
bool  arena_release(DevCtx *c, cudaStream_t st);
void *device_blob(DevCtx *c, const void *host, size_t bytes, cudaStream_t st);   // cached device copy of a small host table

// one long-lived host thread per device; job(device, first_frame, end_frame) runs there with the device selected (host_chain.cu)
int   run_on_devices(const char *who, int ndevices, int nframes, const std::function<bool(int, int, int)> &job);
void  stop_device_workers();

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// grid.y carries the frame index (<= 65535): longer batches are cut into launches of 32768 frames
template <class F>
int per_frame_chunk(int nframes, F launch)
{
    for (int f0 = 0; f0 < nframes; f0 += 32768)
        if (!launch(f0, nframes - f0 < 32768 ? nframes - f0 : 32768)) return 0;
    return 1;
}

}  // namespace acgpu
