/*
 * scb_microbench.inl -- on-box peaks of the three resources the simulator kernel K2 lives on, so that the
 * roofline denominators in bench.py are measured numbers, not data-sheet figures ("measure it with
 * a LOP3 microbenchmark on the box ... same for POPC and shared-memory bandwidth"):
 *     kind 0  LOP3   32-bit three-input logic ops          (result: thread-level ops per second)
 *     kind 1  POPC   32-bit population counts              (ops per second)
 *     kind 2  LDS    conflict-free 128-bit shared loads    (bytes per second)
 * Every CTA slot of every SM is filled, each thread runs independent chains (8 accumulators) so that the pipes,
 * not latency, bound the loop.  Included by scb_api.cu; not compiled into the host-thread emulator.
 */
#ifndef SCB_EMU
namespace {

#define SCB_MB_ITERS 2048

__global__ void __launch_bounds__(512) scb_k_mb_lop3(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < SCB_MB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)      /* table 0x1E = a ^ (b | c) over a neighbouring chain: two steps do not fold into one */
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x1E;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(i & 1 ? a : b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[0] = s;
}

__global__ void __launch_bounds__(512) scb_k_mb_popc(uint32_t *out)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i * 0x9E3779B9u;
    for (int it = 0; it < SCB_MB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 0x12345u) out[0] = s;
}

__global__ void __launch_bounds__(1024) scb_k_mb_lds(uint32_t *out)
{
    extern __shared__ __align__(16) unsigned char mbsm[];
    uint4 *sm = reinterpret_cast<uint4 *>(mbsm);
    for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) sm[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 16u;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int it = 0; it < SCB_MB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {    /* 8 rows of 16 KB: a warp reads 512 consecutive bytes per instruction; volatile, or ptxas
                                            hoists the loop-invariant loads out of the loop */
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                         : "r"(base + (uint32_t)i * 16384u));
            acc0 ^= v0; acc1 ^= v1; acc2 ^= v2; acc3 ^= v3;
        }
    }
    if ((acc0 ^ acc1 ^ acc2 ^ acc3) == 0x12345u) out[0] = acc0;
}

}  // namespace

extern "C" int scb_microbench(scb_ctx *c, int kind, double *perSecond)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!perSecond) return fail(SCB_ERR_INVALID, "null argument");
    if (kind < 0 || kind > 2) return fail(SCB_ERR_INVALID, "unknown microbenchmark %d", kind);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if ((rc = c->dFlush.ensure(4096))) return rc;
    uint32_t *out = reinterpret_cast<uint32_t *>(c->dFlush.p);
    cudaStream_t st = c->stream;
    const int ldsSmem = 8 * 16384;
    if (kind == 2) CK(cudaFuncSetAttribute(scb_k_mb_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, ldsSmem));
    const unsigned grid = (unsigned)c->numSM * (kind == 2 ? 1u : 4u);
    const unsigned block = kind == 2 ? 1024u : 512u;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                    /* first repetition is the warm-up */
        CK(cudaEventRecord(c->ev[6], st));
        if (kind == 0) scb_k_mb_lop3<<<grid, block, 0, st>>>(out, 0x0F0F1234u + rep, 0x33CC5678u);
        else if (kind == 1) scb_k_mb_popc<<<grid, block, 0, st>>>(out);
        else scb_k_mb_lds<<<grid, block, ldsSmem, st>>>(out);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev[7], st));
        CK(cudaEventSynchronize(c->ev[7]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
        const double work = (double)grid * block * SCB_MB_ITERS * 8.0 * (kind == 2 ? 16.0 : 1.0);
        if (rep > 0 && ms > 0.f) best = std::max(best, work / (ms * 1e-3));
    }
    *perSecond = best;
    return SCB_OK;
}
#else
extern "C" int scb_microbench(scb_ctx *, int, double *)
{
    return fail(SCB_ERR_UNSUPPORTED, "microbenchmarks need the CUDA build");
}
#endif
