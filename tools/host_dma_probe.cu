// host_dma_probe.cu -- what can this box's host side deliver to N GPUs at once?  (profiling aid, not part of the library)
//
// The end-to-end numbers of bench.py are bounded by pinned-memory copies.  On the 8-GPU box the aggregate stalls near
// 115 GB/s of host traffic, far below 8 PCIe Gen5 links; this probe separates the candidates:
//   --mode threads|procs      one process driving every GPU from its own thread, or one process per GPU (fork before CUDA)
//   --mem hostalloc|wc|register|hugetlb|thp
//                             cudaHostAlloc, or malloc'd / MAP_HUGETLB / madvise(MADV_HUGEPAGE) memory + cudaHostRegister
//   --dir h2d|d2h|both        one direction alone or both at once (two streams per GPU)
//   --region MB               pinned bytes per GPU and direction that the copies walk through
//   --chunk MB                bytes per copy call
//   --numa                    bind each GPU's thread to the CPUs of the GPU's NUMA node before allocating (first touch)
// Every GPU starts behind one barrier; rates are bytes / the common wall-clock window.  One JSON line on stdout.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

struct Shared {
    pthread_barrier_t start, stop;
    double h2d_gbs[16], d2h_gbs[16], window[16];
    int numa_node[16], failed[16];
};

static int ngpu = 1, use_procs = 0, use_numa = 0;
static const char *mem_kind = "hostalloc", *dir_kind = "both";
static size_t region = 512u << 20, chunk = 16u << 20;
static double seconds = 2.0;
static Shared *sh;

static double now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "gpu %d: %s: %s\n", g, #x, cudaGetErrorString(e_)); sh->failed[g] = 1; goto out; } } while (0)

static int gpu_numa_node(int g)
{
    char bdf[32] = {0}, path[128];
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), g) != cudaSuccess) return -1;
    for (char *p = bdf; *p; p++) if (*p >= 'A' && *p <= 'Z') *p += 32;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    return node;
}

static void bind_to_node(int node)
{
    char path[128], list[4096] = {0};
    if (node < 0) return;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return;
    if (!fgets(list, sizeof(list), f)) { fclose(f); return; }
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(NULL, ",\n")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) CPU_SET(c, &set); }
        else if (sscanf(tok, "%d", &a) == 1) CPU_SET(a, &set);
    }
    sched_setaffinity(0, sizeof(set), &set);
}

static void *host_mem(size_t bytes, int g, int *registered, int is_input = 0)
{
    void *p = NULL;
    *registered = 0;
    if (!strcmp(mem_kind, "hostalloc") || !strcmp(mem_kind, "wc")) {
        /* wc: the UPLOAD buffers are write-combined (the CPU only writes them; the device reads them without cache snoops) */
        const unsigned flags = cudaHostAllocPortable | (!strcmp(mem_kind, "wc") && is_input ? cudaHostAllocWriteCombined : 0);
        if (cudaHostAlloc(&p, bytes, flags) != cudaSuccess) return NULL;
        memset(p, g + 1, bytes);
        return p;
    }
    if (!strcmp(mem_kind, "hugetlb")) {
        p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (p == MAP_FAILED) { perror("mmap(MAP_HUGETLB)"); return NULL; }
    } else {
        p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return NULL;
        if (!strcmp(mem_kind, "thp")) madvise(p, bytes, MADV_HUGEPAGE);
    }
    memset(p, g + 1, bytes);                    /* first touch on this thread's node */
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return NULL; }
    *registered = 1;
    return p;
}

static void *gpu_worker(void *arg)
{
    const int g = (int)(intptr_t)arg;
    const bool do_in = strcmp(dir_kind, "d2h") != 0, do_out = strcmp(dir_kind, "h2d") != 0;
    uint8_t *hin = NULL, *hout = NULL, *din = NULL, *dout = NULL;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int reg_in = 0, reg_out = 0;
    size_t bytes_in = 0, bytes_out = 0, pos_in = 0, pos_out = 0;
    double t0 = 0, t1 = 0;
    bool at_start = false, at_stop = false;
    CK(cudaSetDevice(g));
    sh->numa_node[g] = gpu_numa_node(g);
    if (use_numa) bind_to_node(sh->numa_node[g]);
    CK(cudaFree(0));
    if (do_in)  { hin = (uint8_t *)host_mem(region, g, &reg_in, 1);   if (!hin)  { sh->failed[g] = 1; goto out; } CK(cudaMalloc(&din, chunk)); }
    if (do_out) { hout = (uint8_t *)host_mem(region, g, &reg_out); if (!hout) { sh->failed[g] = 1; goto out; } CK(cudaMalloc(&dout, chunk)); }
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    // warm-up
    if (do_in) CK(cudaMemcpyAsync(din, hin, chunk, cudaMemcpyHostToDevice, s_in));
    if (do_out) CK(cudaMemcpyAsync(hout, dout, chunk, cudaMemcpyDeviceToHost, s_out));
    CK(cudaStreamSynchronize(s_in));
    CK(cudaStreamSynchronize(s_out));
    pthread_barrier_wait(&sh->start);
    at_start = true;
    t0 = now();
    // keep a few copies queued per direction; stop issuing when the window is over
    while (now() - t0 < seconds) {
        for (int k = 0; k < 4; k++) {
            if (do_in)  { CK(cudaMemcpyAsync(din, hin + pos_in, chunk, cudaMemcpyHostToDevice, s_in));     pos_in += chunk;  if (pos_in + chunk > region) pos_in = 0;   bytes_in += chunk; }
            if (do_out) { CK(cudaMemcpyAsync(hout + pos_out, dout, chunk, cudaMemcpyDeviceToHost, s_out)); pos_out += chunk; if (pos_out + chunk > region) pos_out = 0; bytes_out += chunk; }
        }
        CK(cudaStreamSynchronize(s_in));
        CK(cudaStreamSynchronize(s_out));
    }
    t1 = now();
    sh->window[g] = t1 - t0;
    sh->h2d_gbs[g] = bytes_in / (t1 - t0) / 1e9;
    sh->d2h_gbs[g] = bytes_out / (t1 - t0) / 1e9;
out:
    if (!at_start) pthread_barrier_wait(&sh->start);
    if (!at_stop) pthread_barrier_wait(&sh->stop);
    if (hin)  { if (reg_in) cudaHostUnregister(hin); else cudaFreeHost(hin); }
    if (hout) { if (reg_out) cudaHostUnregister(hout); else cudaFreeHost(hout); }
    if (din) cudaFree(din);
    if (dout) cudaFree(dout);
    return NULL;
}

int main(int argc, char **argv)
{
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) ngpu = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--mode") && i + 1 < argc) use_procs = !strcmp(argv[++i], "procs");
        else if (!strcmp(argv[i], "--mem") && i + 1 < argc) mem_kind = argv[++i];
        else if (!strcmp(argv[i], "--dir") && i + 1 < argc) dir_kind = argv[++i];
        else if (!strcmp(argv[i], "--region") && i + 1 < argc) region = (size_t)atol(argv[++i]) << 20;
        else if (!strcmp(argv[i], "--chunk") && i + 1 < argc) chunk = (size_t)atol(argv[++i]) << 20;
        else if (!strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = atof(argv[++i]);
        else if (!strcmp(argv[i], "--numa")) use_numa = 1;
        else { fprintf(stderr, "host_dma_probe: bad argument %s\n", argv[i]); return 2; }
    }
    if (ngpu < 1 || ngpu > 16 || chunk == 0 || region < chunk) { fprintf(stderr, "host_dma_probe: bad sizes\n"); return 2; }
    sh = (Shared *)mmap(NULL, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(sh, 0, sizeof(*sh));
    pthread_barrierattr_t ba;
    pthread_barrierattr_init(&ba);
    pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->start, &ba, (unsigned)ngpu);
    pthread_barrier_init(&sh->stop, &ba, (unsigned)ngpu);
    if (use_procs) {
        for (int g = 0; g < ngpu; g++) {
            pid_t pid = fork();
            if (pid == 0) { gpu_worker((void *)(intptr_t)g); _exit(0); }
        }
        while (wait(NULL) > 0) {}
    } else {
        pthread_t th[16];
        for (int g = 0; g < ngpu; g++) pthread_create(&th[g], NULL, gpu_worker, (void *)(intptr_t)g);
        for (int g = 0; g < ngpu; g++) pthread_join(th[g], NULL);
    }
    double in = 0, out = 0;
    int failed = 0;
    for (int g = 0; g < ngpu; g++) { in += sh->h2d_gbs[g]; out += sh->d2h_gbs[g]; failed += sh->failed[g]; }
    printf("{\"gpus\": %d, \"mode\": \"%s\", \"mem\": \"%s\", \"dir\": \"%s\", \"numa_bind\": %d, \"region_mb\": %zu, \"chunk_mb\": %zu, "
           "\"failed\": %d, \"h2d_gbs_total\": %.2f, \"d2h_gbs_total\": %.2f, \"host_gbs_total\": %.2f, \"per_gpu\": [",
           ngpu, use_procs ? "procs" : "threads", mem_kind, dir_kind, use_numa, region >> 20, chunk >> 20, failed, in, out, in + out);
    for (int g = 0; g < ngpu; g++)
        printf("%s{\"h2d\": %.2f, \"d2h\": %.2f, \"numa\": %d}", g ? ", " : "", sh->h2d_gbs[g], sh->d2h_gbs[g], sh->numa_node[g]);
    printf("]}\n");
    return failed ? 1 : 0;
}
