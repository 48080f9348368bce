/*
 * emu_cuda.h -- a tiny host-thread emulator of the CUDA subset the scbonita_b200 kernels use.
 *
 * TEST INFRASTRUCTURE ONLY.  The build container has no GPU, so the CPU test-suite compiles
 * the *unchanged* kernel and host-API sources (scbonita_b200/csrc) with g++ -DSCB_EMU against
 * this header to exercise their logic: every CUDA thread becomes a std::thread, __syncthreads
 * is a std::barrier, warp collectives rendezvous on a per-warp barrier.  It is slow (tiny
 * problem sizes only), lives under tests/, is built into tests/emu/_build/, and is never
 * looked up by the product loader (scbonita_b200/_lib.py) -- the product has no CPU path.
 */
#ifndef SCB_EMU_CUDA_H
#define SCB_EMU_CUDA_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

unsigned char *emu_dyn_smem();
void emu_yield();
/* tensor memory of the emulated SM: 128 lanes x 512 columns of 32 bits per block; taddr = lane << 16 | column */
uint32_t *emu_tmem_cell(uint32_t taddr, int k);
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
void __syncthreads();
int __syncthreads_or(int pred);
void __syncwarp(unsigned mask = 0xffffffffu);
unsigned __ballot_sync(unsigned mask, int pred);
unsigned __match_any_sync(unsigned mask, unsigned value);   /* all 32 lanes must call it */
uint32_t emu_shfl(uint32_t v, int srcLane);
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int laneMask) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ laneMask); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return emu_shfl(v, src & 31); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
template <class T> static inline T __ldg(const T *p) { return *p; }

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

/* ---- runtime API subset ---------------------------------------------------------------- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t sharedMemPerBlockOptin; int major, minor; size_t totalGlobalMem; };

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }

/* ---- launch ------------------------------------------------------------------------------ */
#include <functional>
void emu_launch_impl(const std::function<void()> &body, dim3 grid, dim3 block, size_t smem);
template <class F, class... Args>
static inline void emu_launch(F kernel, dim3 grid, dim3 block, size_t smem, Args... args)
{
    emu_launch_impl([=]() { kernel(args...); }, grid, block, smem);
}

#endif
