/*
 * emu_cuda.cpp -- host-thread CUDA emulator (see emu_cuda.h).  TEST INFRASTRUCTURE ONLY.
 *
 * One std::thread per CUDA thread, all blocks of a launch alive at once (so atomics between
 * blocks behave), block barrier = std::barrier, warp collectives = per-warp barrier plus an
 * exchange array.  Launches are synchronous.
 */
#include "emu_cuda.h"
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace {
struct WarpCtx {
    std::barrier<> bar;
    uint32_t slot[32];
    explicit WarpCtx(int n) : bar(n) {}
};
struct BlockCtx {
    std::barrier<> bar;
    std::vector<std::unique_ptr<WarpCtx>> warps;
    unsigned char *smem;
    std::vector<uint32_t> tmem;
    std::atomic<int> orAcc[2];
    explicit BlockCtx(int n) : bar(n), smem(nullptr), tmem(128 * 512, 0xDEADBEEFu) { orAcc[0] = 0; orAcc[1] = 0; }
};
thread_local BlockCtx *tl_block = nullptr;
thread_local WarpCtx *tl_warp = nullptr;
}  // namespace

unsigned char *emu_dyn_smem() { return tl_block->smem; }
void emu_yield() { std::this_thread::yield(); }
uint32_t *emu_tmem_cell(uint32_t taddr, int k)
{
    const unsigned lane = ((taddr >> 16) & 0xffffu) + (threadIdx.x & 31u), col = (taddr & 0xffffu) + (unsigned)k;
    if (lane >= 128u || col >= 512u) abort();            /* out-of-range tensor-memory access */
    return &tl_block->tmem[(size_t)lane * 512 + col];
}
void __syncthreads() { tl_block->bar.arrive_and_wait(); }
void __syncwarp(unsigned) { tl_warp->bar.arrive_and_wait(); }
int __syncthreads_or(int pred)
{
    BlockCtx *b = tl_block;
    if (pred) b->orAcc[0].store(1);
    b->bar.arrive_and_wait();
    const int r = b->orAcc[0].load();
    b->bar.arrive_and_wait();
    if (threadIdx.x == 0) b->orAcc[0].store(0);
    b->bar.arrive_and_wait();
    return r;
}

unsigned __ballot_sync(unsigned, int pred)
{
    WarpCtx *w = tl_warp;
    const int lane = threadIdx.x & 31;
    w->slot[lane] = pred ? 1u : 0u;
    w->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (w->slot[i] & 1u) << i;
    w->bar.arrive_and_wait();
    return r;
}

uint32_t emu_shfl(uint32_t v, int srcLane)
{
    WarpCtx *w = tl_warp;
    const int lane = threadIdx.x & 31;
    w->slot[lane] = v;
    w->bar.arrive_and_wait();
    const uint32_t r = w->slot[srcLane & 31];
    w->bar.arrive_and_wait();
    return r;
}

unsigned __match_any_sync(unsigned, unsigned v)
{
    WarpCtx *w = tl_warp;
    const int lane = threadIdx.x & 31;
    w->slot[lane] = v;
    w->bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (w->slot[i] == v) r |= 1u << i;
    w->bar.arrive_and_wait();
    return r;
}

void emu_launch_impl(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem)
{
    /* one launch at a time, process-wide: host threads that launch concurrently (the re-entrancy tests) would
     * otherwise multiply the number of live std::threads */
    static std::mutex launchMutex;
    std::lock_guard<std::mutex> lock(launchMutex);
    const int nthr = (int)block.x;
    const int nblk = (int)(grid.x * grid.y);
    if (nthr % 32 != 0) abort();
    /* blocks run in waves of a few at a time (like CTAs on a small GPU); no kernel here needs all blocks
     * of a grid to be resident at once */
    const int wave = nthr > 512 ? 2 : 4;
    for (int b0 = 0; b0 < nblk; b0 += wave) {
        const int nb = nblk - b0 < wave ? nblk - b0 : wave;
        std::vector<std::unique_ptr<BlockCtx>> blocks;
        std::vector<std::thread> threads;
        for (int b = 0; b < nb; ++b) {
            auto bc = std::make_unique<BlockCtx>(nthr);
            for (int w = 0; w < nthr / 32; ++w) {
                bc->warps.push_back(std::make_unique<WarpCtx>(32));
                memset(bc->warps.back()->slot, 0, sizeof(bc->warps.back()->slot));
            }
            bc->smem = (unsigned char *)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
            memset(bc->smem, 0xCD, smem);          /* shared memory starts undefined */
            blocks.push_back(std::move(bc));
        }
        for (int b = 0; b < nb; ++b)
            for (int t = 0; t < nthr; ++t) {
                BlockCtx *bc = blocks[b].get();
                const int gb = b0 + b;
                threads.emplace_back([=, &body]() {
                    threadIdx = dim3(t); blockIdx = dim3(gb % grid.x, gb / grid.x); blockDim = block; gridDim = grid;
                    tl_block = bc; tl_warp = bc->warps[t / 32].get();
                    body();
                    /* a thread that returns early must not leave the others stuck at a barrier */
                    bc->warps[t / 32]->bar.arrive_and_drop();
                    bc->bar.arrive_and_drop();
                });
            }
        for (auto &th : threads) th.join();
        for (auto &bc : blocks) free(bc->smem);
    }
}

/* ---- runtime API ------------------------------------------------------------------------ */
struct emu_event { std::chrono::steady_clock::time_point t; };
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "host-thread emulator");
    p->multiProcessorCount = 2; p->sharedMemPerBlockOptin = 232448; p->major = 10; p->minor = 0;
    p->totalGlobalMem = (size_t)1 << 34;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, ((n + 255) / 256 + 1) * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "emulator"; }
