/*
 * scb_kernels.cuh -- sm_100a kernels of the rule-inference fitness hot path.
 *
 *   K0  scb_k_compile    reference-format rule tables (or upstream triples) -> per-individual programs of 3-input
 *                        truth tables (scb_compile.h), sorted by table; one CTA per (replaced) individual.  Also the
 *                        REDUCED programs: constants (knocked nodes, their copies, the AND-terms they erase) propagated
 *                        to depths 2, 5, 8, 12 and to the fixed point, each exact from step depth + 2 on; and every
 *                        program as the IMAGE the simulator's warps consume (runs of nodes with the same table, cut into
 *                        cost-balanced pieces, one per warp)
 *   K1  scb_k_pack       int32 / uint8 binarized matrix -> node-major 32-cells-per-word S_0
 *   K1b scb_k_pack_csr   the same straight from a CSR matrix
 *   K2  scb_k_sim<K,TM>  the simulator: persistent CTAs, one (individual, cell tile) per work item, state
 *                        ping-pong in shared memory, one LOP3 per node per 32 cells inside loops specialised per truth
 *                        table (generated PTX, one brx.idx per run), snapshot in tensor memory (TM), static schedule of
 *                        compares / snapshots / program switches, resolver items and handed-back tiles taken from the
 *                        same work loop
 *   K3  scb_resolve_item / scb_k_resolve   exact second pass for the tiles K2 could not prove "attractor found" for
 *                        (literally getAttractCycle2 on packed words), target state in tensor memory inside K2;
 *       scb_k_fill       the reference's fill-forward across chunks: a lookup of the recorded error column
 *                        of the nearest earlier found cell, one warp per item
 *   KG  scb_k_generic    the general simulator: any [7][3] table, any node count, state in global memory; the path
 *                        for input K0 / K2 cannot hold (exact, slow), incl. SURVEY Q8 from 1000 nodes on
 *   K4  scb_k_traj + scb_k_window   trajectories, attractor windows and their states (C `cluster` and Python
 *                        `__cluster` semantics); scb_k_dedupe_*: the distinct states in first-seen order
 *   K5  scb_k_hamming<NW4>  cell x attractor Hamming distance + first-minimum argmin (carry-save popcount);
 *       scb_k_pack_attr / scb_k_restride  attractors into the kernel's bit rows
 *       scb_k_gather     the generation's fitness all-gather over peer memory (last kernel of the generation's graph)
 *       scb_k_scatter_rows  replaced individuals of a resident population into their slots
 *   importanceScore (simulator.c:555-755) has no kernel of its own: K0 compiles its single-literal rules, K2/K3 run
 *   in the SCB_MODE_IMPORTANCE output shape (per-node counts over fixed-point cells), one program per knock array
 *
 * What K2/K3 compute (reference simulator.c:350-485): per cell the synchronous trajectory
 * S_0..S_99; found <=> S_99 equals some earlier S_j; error bits e = S_98 xor S_0 when found, else the previous
 * cell's e.  The kernels never store the trajectory: "found" is proven in flight by any observed repeat
 * S_t == S_j (j < t <= 99), which implies transient + period <= 99; cells without such a proof go to K3.
 *
 * Data layout in shared memory (per CTA): two state buffers [N][32*K] uint32 (row = node, lane l owns words
 * l*K .. l*K+K-1, so every access is a conflict-free LDS/STS.(32*K)), followed by one program image and a few
 * 32*K-word masks.
 *
 * The file also compiles with g++ -DSCB_EMU against tests/emu (host-thread emulator, test infrastructure).
 */
#ifndef SCB_KERNELS_CUH
#define SCB_KERNELS_CUH
#include "scb_types.h"
#include "scb_compile.h"

#ifdef SCB_EMU
#define SCB_DYN_SMEM(name) unsigned char *name = emu_dyn_smem()
#else
#define SCB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

/* ------------------------------------------------------------------------------------------
 * small device helpers
 * ---------------------------------------------------------------------------------------- */
template <int K> struct ScbVec { uint32_t v[K]; };

template <int K> SCB_DEV ScbVec<K> scb_ld(const unsigned char *p)
{
    ScbVec<K> r;
    if (K == 4) { uint4 t = *reinterpret_cast<const uint4 *>(p); r.v[0] = t.x; r.v[1 % K] = t.y; r.v[2 % K] = t.z; r.v[3 % K] = t.w; }
    else if (K == 2) { uint2 t = *reinterpret_cast<const uint2 *>(p); r.v[0] = t.x; r.v[1 % K] = t.y; }
    else { r.v[0] = *reinterpret_cast<const uint32_t *>(p); }
    return r;
}
template <int K> SCB_DEV void scb_st(unsigned char *p, const ScbVec<K> &r)
{
    if (K == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(r.v[0], r.v[1 % K], r.v[2 % K], r.v[3 % K]);
    else if (K == 2) *reinterpret_cast<uint2 *>(p) = make_uint2(r.v[0], r.v[1 % K]);
    else *reinterpret_cast<uint32_t *>(p) = r.v[0];
}
/* Shared-memory accesses of the hot loop go through 32-bit shared-window addresses (ld.shared / st.shared):
 * with generic pointers the compiler re-derives the window base (S2R + LEA) and does 64-bit address
 * arithmetic for every access.  Under the emulator an "address" is simply a host pointer. */
#if defined(__CUDA_ARCH__)
typedef uint32_t scb_saddr;
SCB_DEV scb_saddr scb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int K> SCB_DEV ScbVec<K> scb_lds(scb_saddr a)
{
    ScbVec<K> r;
    if (K == 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1 % K]), "=r"(r.v[2 % K]), "=r"(r.v[3 % K]) : "r"(a));
    else if (K == 2) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1 % K]) : "r"(a));
    else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.v[0]) : "r"(a));
    return r;
}
template <int K> SCB_DEV void scb_sts(scb_saddr a, const ScbVec<K> &r)
{
    if (K == 4) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(r.v[0]), "r"(r.v[1 % K]), "r"(r.v[2 % K]), "r"(r.v[3 % K]) : "memory");
    else if (K == 2) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(r.v[0]), "r"(r.v[1 % K]) : "memory");
    else asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(r.v[0]) : "memory");
}
SCB_DEV uint4 scb_lds_desc(scb_saddr a)
{
    uint4 d;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(a));
    return d;
}
#else
typedef unsigned char *scb_saddr;
SCB_DEV scb_saddr scb_smem_addr(const void *p) { return reinterpret_cast<unsigned char *>(const_cast<void *>(p)); }
template <int K> SCB_DEV ScbVec<K> scb_lds(scb_saddr a) { return scb_ld<K>(a); }
template <int K> SCB_DEV void scb_sts(scb_saddr a, const ScbVec<K> &r) { scb_st<K>(a, r); }
SCB_DEV uint4 scb_lds_desc(scb_saddr a) { return *reinterpret_cast<const uint4 *>(a); }
#endif

/* ---- TMA bulk copies of the S_0 tile: one cp.async.bulk (SASS UBLKCP) per node row, completion on an mbarrier ---- */
#if defined(__CUDA_ARCH__)
SCB_DEV void scb_mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SCB_DEV void scb_mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
SCB_DEV void scb_bulk_g2s(uint32_t dstSmem, const void *srcGlobal, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dstSmem), "l"(srcGlobal), "r"(bytes), "r"(mbar) : "memory");
}
SCB_DEV void scb_mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nSCB_WAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SCB_WAIT_DONE;\n"
                 "bra SCB_WAIT_LOOP;\nSCB_WAIT_DONE:\n}" :: "r"(mbar), "r"(parity) : "memory");
}
SCB_DEV void scb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

/* ---- 16-byte asynchronous global -> shared copies (LDGSTS): the next reduced program is fetched into its staging
 * buffer while the steps of the current one run ---- */
SCB_DEV void scb_cp_async16(void *smemDst, const void *gsrc)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smemDst)), "l"(gsrc) : "memory");
#else
    *reinterpret_cast<uint4 *>(smemDst) = *reinterpret_cast<const uint4 *>(gsrc);
#endif
}
SCB_DEV void scb_cp_async_commit()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
SCB_DEV void scb_cp_async_wait()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

/* ---- tensor memory as a scratchpad --------------------------------------------------------------------
 * The simulator uses no tensor cores, so the SM's 256 KB of TMEM (128 lanes x 512 columns x 32 bit) is free.
 * K2 keeps its snapshot S_s there instead of in global memory: a warp owns the 32 lanes of its quarter
 * (warp % 4), lane l of the warp maps to TMEM lane 32*(warp%4)+l, and every node the warp updates gets K
 * consecutive columns.  tcgen05.st / tcgen05.ld (.32x32b.xK) then move exactly one state row segment
 * (K words per lane) between registers and TMEM, on a datapath that does not compete with LDS/STS. */
#define SCB_TMEM_OFF 0xffffffffu
#if defined(__CUDA_ARCH__)
SCB_DEV void scb_tmem_alloc(uint32_t smemSlotAddr, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemSlotAddr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SCB_DEV void scb_tmem_dealloc(uint32_t base, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(ncols) : "memory");
}
SCB_DEV void scb_tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SCB_DEV void scb_tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int K> SCB_DEV void scb_tmem_st(uint32_t taddr, const ScbVec<K> &r)
{
    if (K == 4) asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(r.v[0]), "r"(r.v[1 % K]), "r"(r.v[2 % K]), "r"(r.v[3 % K]) : "memory");
    else if (K == 2) asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "r"(r.v[0]), "r"(r.v[1 % K]) : "memory");
    else asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr), "r"(r.v[0]) : "memory");
}
/* asynchronous: the registers may only be read after scb_tmem_wait_ld, which is tied to them as an operand */
template <int K> SCB_DEV void scb_tmem_ld(uint32_t taddr, ScbVec<K> &r)
{
    if (K == 4) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1 % K]), "=r"(r.v[2 % K]), "=r"(r.v[3 % K]) : "r"(taddr) : "memory");
    else if (K == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r.v[0]), "=r"(r.v[1 % K]) : "r"(taddr) : "memory");
    else asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r.v[0]) : "r"(taddr) : "memory");
}
template <int K> SCB_DEV void scb_tmem_wait_ld(ScbVec<K> &r)
{
    if (K == 4) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r.v[0]), "+r"(r.v[1 % K]), "+r"(r.v[2 % K]), "+r"(r.v[3 % K]) :: "memory");
    else if (K == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r.v[0]), "+r"(r.v[1 % K]) :: "memory");
    else asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r.v[0]) :: "memory");
}
SCB_DEV void scb_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#elif defined(SCB_EMU)
/* host-thread emulator: tensor memory is a per-block array (tests/emu) */
SCB_DEV void scb_tmem_alloc(uint32_t, uint32_t) {}
SCB_DEV void scb_tmem_dealloc(uint32_t, uint32_t) {}
SCB_DEV void scb_tmem_fence_before() {}
SCB_DEV void scb_tmem_fence_after() {}
template <int K> SCB_DEV void scb_tmem_st(uint32_t taddr, const ScbVec<K> &r) { for (int k = 0; k < K; ++k) *emu_tmem_cell(taddr, k) = r.v[k]; }
template <int K> SCB_DEV void scb_tmem_ld(uint32_t taddr, ScbVec<K> &r) { for (int k = 0; k < K; ++k) r.v[k] = *emu_tmem_cell(taddr, k); }
template <int K> SCB_DEV void scb_tmem_wait_ld(ScbVec<K> &) {}
SCB_DEV void scb_tmem_wait_st() {}
#else
SCB_DEV void scb_tmem_alloc(uint32_t, uint32_t) {}
SCB_DEV void scb_tmem_dealloc(uint32_t, uint32_t) {}
SCB_DEV void scb_tmem_fence_before() {}
SCB_DEV void scb_tmem_fence_after() {}
template <int K> SCB_DEV void scb_tmem_st(uint32_t, const ScbVec<K> &) {}
template <int K> SCB_DEV void scb_tmem_ld(uint32_t, ScbVec<K> &) {}
template <int K> SCB_DEV void scb_tmem_wait_ld(ScbVec<K> &) {}
SCB_DEV void scb_tmem_wait_st() {}
#endif

/* global memory, L2 only (the snapshot scratch is written and re-read by the same thread) */
template <int K> SCB_DEV ScbVec<K> scb_ldcg(const unsigned char *p)
{
#if defined(__CUDA_ARCH__)
    ScbVec<K> r;
    if (K == 4) { uint4 t = __ldcg(reinterpret_cast<const uint4 *>(p)); r.v[0] = t.x; r.v[1 % K] = t.y; r.v[2 % K] = t.z; r.v[3 % K] = t.w; }
    else if (K == 2) { uint2 t = __ldcg(reinterpret_cast<const uint2 *>(p)); r.v[0] = t.x; r.v[1 % K] = t.y; }
    else { r.v[0] = __ldcg(reinterpret_cast<const uint32_t *>(p)); }
    return r;
#else
    return scb_ld<K>(p);
#endif
}
template <int K> SCB_DEV void scb_stcg(unsigned char *p, const ScbVec<K> &r)
{
#if defined(__CUDA_ARCH__)
    if (K == 4) __stcg(reinterpret_cast<uint4 *>(p), make_uint4(r.v[0], r.v[1 % K], r.v[2 % K], r.v[3 % K]));
    else if (K == 2) __stcg(reinterpret_cast<uint2 *>(p), make_uint2(r.v[0], r.v[1 % K]));
    else __stcg(reinterpret_cast<uint32_t *>(p), r.v[0]);
#else
    scb_st<K>(p, r);
#endif
}
/* read-only global (S_0 rows) */
template <int K> SCB_DEV ScbVec<K> scb_ldg(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    ScbVec<K> r;
    if (K == 4) { uint4 t = __ldg(reinterpret_cast<const uint4 *>(p)); r.v[0] = t.x; r.v[1 % K] = t.y; r.v[2 % K] = t.z; r.v[3 % K] = t.w; }
    else if (K == 2) { uint2 t = __ldg(reinterpret_cast<const uint2 *>(p)); r.v[0] = t.x; r.v[1 % K] = t.y; }
    else { r.v[0] = __ldg(p); }
    return r;
#else
    return scb_ld<K>(reinterpret_cast<const unsigned char *>(p));
#endif
}

SCB_DEV unsigned scb_ld_volatile(const unsigned int *p)
{
#if defined(SCB_EMU)
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);          /* the host emulation needs a real atomic load */
#else
    return *reinterpret_cast<const volatile unsigned int *>(p);
#endif
}
SCB_DEV void scb_backoff()
{
#if defined(__CUDA_ARCH__)
    __nanosleep(256);
#elif defined(SCB_EMU)
    emu_yield();
#endif
}

SCB_DEV long long scb_clock()
{
#if defined(__CUDA_ARCH__) && defined(SCB_PHASE_CLOCKS)
    return clock64();          /* per-phase cycle counters of K2 (build with -DSCB_PHASE_CLOCKS); off by default */
#else
    return 0;
#endif
}

SCB_DEV unsigned scb_warp_sum(unsigned v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* ---- truth-table application: one LOP3 per word with the table as an immediate ----------
 * The table is per-node runtime data, LOP3's table is an immediate.  A C++ switch over the 256
 * tables compiles to a 7-deep compare/branch tree; the generated inline PTX (scb_lut_ptx.inc)
 * uses brx.idx instead, which ptxas lowers to one LDC (jump table in the constant bank) + BRX. */
#if defined(__CUDA_ARCH__)
#include "scb_lut_ptx.inc"
#endif

SCB_HD inline unsigned scb_expand2(unsigned l4)
{
    return ((l4 & 1u) ? 0x03u : 0u) | ((l4 & 2u) ? 0x0Cu : 0u) | ((l4 & 4u) ? 0x30u : 0u) | ((l4 & 8u) ? 0xC0u : 0u);
}
SCB_HD inline unsigned scb_compress2(unsigned l8)
{
    return ((l8 >> 0) & 1u) | (((l8 >> 2) & 1u) << 1) | (((l8 >> 4) & 1u) << 2) | (((l8 >> 6) & 1u) << 3);
}

/* three-input node: any of the 256 tables (lut8, lop3 convention) */
template <int K> SCB_DEV ScbVec<K> scb_apply3(unsigned lut, const ScbVec<K> &a, const ScbVec<K> &b, const ScbVec<K> &c)
{
    ScbVec<K> r;
#if defined(__CUDA_ARCH__)
    if (K == 4)
        asm(SCB_LUT3_PTX_K4 : "=r"(r.v[0]), "=r"(r.v[1 % K]), "=r"(r.v[2 % K]), "=r"(r.v[3 % K])
            : "r"(a.v[0]), "r"(a.v[1 % K]), "r"(a.v[2 % K]), "r"(a.v[3 % K]), "r"(b.v[0]), "r"(b.v[1 % K]), "r"(b.v[2 % K]),
              "r"(b.v[3 % K]), "r"(c.v[0]), "r"(c.v[1 % K]), "r"(c.v[2 % K]), "r"(c.v[3 % K]), "r"(lut));
    else if (K == 2)
        asm(SCB_LUT3_PTX_K2 : "=r"(r.v[0]), "=r"(r.v[1 % K])
            : "r"(a.v[0]), "r"(a.v[1 % K]), "r"(b.v[0]), "r"(b.v[1 % K]), "r"(c.v[0]), "r"(c.v[1 % K]), "r"(lut));
    else
        asm(SCB_LUT3_PTX_K1 : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]), "r"(c.v[0]), "r"(lut));
#else
    for (int k = 0; k < K; ++k) r.v[k] = scb_lut_apply(lut, a.v[k], b.v[k], c.v[k]);
#endif
    return r;
}
/* two-input node: lut4 bit (x0<<1)|x1 */
template <int K> SCB_DEV ScbVec<K> scb_apply2(unsigned l4, const ScbVec<K> &a, const ScbVec<K> &b)
{
    ScbVec<K> r;
#if defined(__CUDA_ARCH__)
    if (K == 4)
        asm(SCB_LUT2_PTX_K4 : "=r"(r.v[0]), "=r"(r.v[1 % K]), "=r"(r.v[2 % K]), "=r"(r.v[3 % K])
            : "r"(a.v[0]), "r"(a.v[1 % K]), "r"(a.v[2 % K]), "r"(a.v[3 % K]), "r"(b.v[0]), "r"(b.v[1 % K]), "r"(b.v[2 % K]),
              "r"(b.v[3 % K]), "r"(l4));
    else if (K == 2)
        asm(SCB_LUT2_PTX_K2 : "=r"(r.v[0]), "=r"(r.v[1 % K]) : "r"(a.v[0]), "r"(a.v[1 % K]), "r"(b.v[0]), "r"(b.v[1 % K]), "r"(l4));
    else
        asm(SCB_LUT2_PTX_K1 : "=r"(r.v[0]) : "r"(a.v[0]), "r"(b.v[0]), "r"(l4));
#else
    for (int k = 0; k < K; ++k) r.v[k] = scb_lut_apply(scb_expand2(l4), a.v[k], b.v[k], b.v[k]);
#endif
    return r;
}

/* ---- the program image --------------------------------------------------------------------------------------
 * K0 sorts the non-constant nodes of a program by truth table and lays the result out the way the simulator's
 * warps consume it (scb_layout_program), once per individual and program, so that K2 only has to copy the image
 * into shared memory (cp.async, while the previous program still runs) and can switch programs by swapping two
 * buffers.  Warp w of S owns a contiguous piece of the sorted list, cut so that every warp moves about the same
 * number of state rows per step (a node costs arity + 1 rows), as RUNS of nodes with the same table:
 *     header {SCB_NOIN | table, SCB_NOIN | n, SCB_NOIN, 0}   then n node entries   ...   header {SCB_NOIN | SCB_CLS_END, ..}
 * Image:  [0] = {n3, n2, n1, n0} (nodes by arity; n0 = constants, which are NOT in the lists)
 *         [1 + w/2] = {offset, entries} of warp w (x, y for even w; z, w for odd w), offset in 16-byte entries from
 *                     the image start, entries without the end header
 *         then the lists, each followed by its end header and two entries of slack (the descriptor prefetch of the
 *         step code reads that far).
 * Node entry: x, y, z = byte offsets of the input rows within a state buffer, SCB_NOIN (sign bit) for an input the
 * table does not depend on; w = byte offset of the node's own row | table << 24.
 * Plain and compare steps run a list through scb_step_runs: one jump per RUN into a loop specialised for that table
 * (no per-node dispatch, no loads for unused inputs, no LOP3 for copies).  Snapshot steps walk the same list entry by
 * entry (scb_step), skipping the headers.  Constant rows are written at steps 1 and 2 straight from the global
 * program (scb_const_rows). */
#define SCB_NOIN 0x80000000u
#define SCB_DST_MASK 0xFFFFFu
#define SCB_CLS_END 256u
#define SCB_PROG_MAX_WARPS 24
#define SCB_LIST_TAIL 3           /* end header + two entries of slack per warp */
SCB_HD inline int scb_img_hdr(int S) { return 1 + (S + 1) / 2; }
/* entries of an image for S warps: header, every node, every run (at most one per distinct table, plus one more
 * for every piece boundary that cuts a run), the tails */
SCB_HD inline size_t scb_img_entries(int N, int S) { return (size_t)scb_img_hdr(S) + (size_t)N + (size_t)(N < 255 ? N : 255) + (size_t)S * (1 + SCB_LIST_TAIL); }
SCB_HD inline size_t scb_img_entries_max(int N) { return scb_img_entries(N, SCB_PROG_MAX_WARPS); }
/* shared-memory carve-up of the simulator for S warps, identical formula on host and device: state buffers, ONE
 * program image (the next one waits in registers, SCB_STAGE_REGS uint4 per thread), masks, scalars.  Kept tight:
 * two CTAs of K = 2 tiles share an SM for a 200-node network (see pick_geometry). */
#define SCB_STAGE_REGS 3
SCB_HD inline size_t scb_smem_bytes(int N, int K, int nbuf, int S)
{
    return (size_t)nbuf * N * 128 * K + scb_img_entries(N, S) * 16 + (size_t)8 * 32 * K * 4 + 512;
}

/* {offset, entries} of the piece `warp` of S executes.  Pieces are consecutive stretches of the table-sorted list, so
 * neighbouring pieces run the same specialised loops; the four warps of a scheduler (warp % 4) take four NEIGHBOURING
 * pieces, so that a scheduler's small instruction cache sees few different loops (measured on the trace build: with
 * piece = warp, warps with identical static work differed by 60 % in time). */
SCB_DEV void scb_img_warp(const uint4 *img, int warp, int S, int &off, int &nw)
{
    const int piece = (S & 3) ? warp : (warp & 3) * (S >> 2) + (warp >> 2);
    const uint4 h = img[1 + (piece >> 1)];
    off = (int)((piece & 1) ? h.z : h.x);
    nw = (int)((piece & 1) ? h.w : h.y);
}

/* table-sorted program g (fields of scb_scatter_by_table) -> image for S warps and rows of ROWB bytes, in global or
 * shared memory.  scr: 3 x 32 ints of shared scratch.  Contains barriers: all threads of the CTA call it. */
SCB_DEV void scb_layout_program(uint4 *img, int *scr, const uint4 *g, int4 cnt, int N, int S, uint32_t ROWB, int tid, int nthr)
{
    const int e1 = cnt.x + cnt.y + cnt.z;
    const unsigned total = (unsigned)(4 * cnt.x + 3 * cnt.y + 2 * cnt.z);
    int *first = scr, *last = scr + 32, *offs = scr + 64;
    if (tid <= S) first[tid] = e1;
    __syncthreads();
#define SCB_WARP_OF(d) ((int)((((d).x >> 16) * (unsigned)S) / total))
    for (int j = tid; j < e1; j += nthr) {
        const int w = SCB_WARP_OF(g[j]);
        if (j == 0 || SCB_WARP_OF(g[j - 1]) != w) first[w] = j;
    }
    __syncthreads();
    if (tid < S) {
        int e = e1;                                       /* the piece ends where the next non-empty one starts */
        for (int w = tid + 1; w < S; ++w) if (first[w] < e1) { e = first[w]; break; }
        last[tid] = e;
    }
    __syncthreads();
    /* entries of every piece: its nodes + its runs (the piece's first node always starts one); one thread per piece
     * (the two reads of g are L2 round trips: done side by side, not one piece after the other) */
    int myN = 0;
    if (tid < S) {
        const int f = first[tid], l = last[tid];
        myN = f < e1 ? (l - f) + (int)((g[l - 1].z >> 16) - (g[f].z >> 16)) + 1 : 0;
        offs[tid] = myN;
    }
    __syncthreads();
    int myO = scb_img_hdr(S);
    if (tid < S) for (int w = 0; w < tid; ++w) myO += offs[w] + SCB_LIST_TAIL;      /* S <= 24 additions from shared memory */
    __syncthreads();
    if (tid < S) {
        offs[tid] = myO;
        unsigned *h = reinterpret_cast<unsigned *>(img + 1 + (tid >> 1)) + ((tid & 1) ? 2 : 0);
        h[0] = (unsigned)myO; h[1] = (unsigned)myN;
        img[myO + myN] = make_uint4(SCB_NOIN | SCB_CLS_END, SCB_NOIN, SCB_NOIN, 0u);
        img[myO + myN + 1] = make_uint4(SCB_NOIN | SCB_CLS_END, SCB_NOIN, SCB_NOIN, 0u);
        img[myO + myN + 2] = make_uint4(SCB_NOIN | SCB_CLS_END, SCB_NOIN, SCB_NOIN, 0u);
    }
    if (tid == 0) img[0] = make_uint4((unsigned)cnt.x, (unsigned)cnt.y, (unsigned)cnt.z, (unsigned)cnt.w);
    __syncthreads();
    for (int j = tid; j < e1; j += nthr) {
        const uint4 d = g[j];
        const int w = SCB_WARP_OF(d);
        const int s0 = first[w], end = last[w];
        const unsigned lut = d.w >> 24, ar = (d.y >> 16) & 3u;
        const bool runStart = j == s0 || (g[j - 1].w >> 24) != lut;
        const int idx = (j - s0) + (int)((d.z >> 16) - (g[s0].z >> 16)) + 1;
        uint4 o;
        o.x = (d.x & 0xFFFFu) * ROWB;
        o.y = ar >= 2u ? (d.y & 0xFFFFu) * ROWB : SCB_NOIN;
        o.z = ar >= 3u ? (d.z & 0xFFFFu) * ROWB : SCB_NOIN;
        o.w = ((d.w & 0xFFFFu) * ROWB) | (lut << 24);
        uint4 *mine = img + offs[w];
        mine[idx] = o;
        if (runStart) {
            int len = 1;
            while (j + len < end && (g[j + len].w >> 24) == lut) ++len;
            mine[idx - 1] = make_uint4(SCB_NOIN | lut, SCB_NOIN | (unsigned)len, SCB_NOIN, 0u);
        }
    }
#undef SCB_WARP_OF
    __syncthreads();
}

#ifndef SCB_SIM_ONLY   /* the simulator's own translation units (scb_sim_tu.cu) leave the other kernels out */
/* ------------------------------------------------------------------------------------------
 * K0: rule compiler.  One CTA per individual, one thread per node; the program is written sorted
 * by arity (stable in node order, so the layout is deterministic).
 * The individual's two [N][7][3] tables are first staged in shared memory with coalesced loads (a thread's
 * own 2 x 84 bytes would otherwise be fetched as scattered 4-byte reads).
 * dynamic shared memory: N * 16 (descriptors) + N (arity) + 64 + 2 * N * 21 * 4 (tables)
 * ---------------------------------------------------------------------------------------- */
/* scratch of the table sort at the end of K0's shared memory: bins[SCB_SORT_BINS] counts -> starts,
 * cost[SCB_SORT_BINS] rows before each bin, run[SCB_SORT_BINS] non-empty bins before each bin, rank[N] (uint16) */
#define SCB_SORT_BINS 258          /* 256 tables, then the two constants */
/* (every part a multiple of 16 bytes: the sorted copy of the program at the end is read as uint4) */
SCB_HD inline size_t scb_sort_smem_bytes(int N) { return (((size_t)3 * SCB_SORT_BINS * 4 + 15) & ~(size_t)15) + (size_t)((N + 7) / 8) * 16 + 32 + 96 * 4 + (size_t)N * 16; }
SCB_HD inline size_t scb_compile_smem_bytes(int N)
{
    return (size_t)N * 16 + (size_t)((N + 15) / 16) * 16 + 64 + (((size_t)2 * N * SCB_MAX_TERMS * SCB_MAX_IN * 4 + 15) & ~(size_t)15) + scb_sort_smem_bytes(N);
}

/* number of inputs a table in lop3 convention depends on */
SCB_HD inline int scb_lut_arity(unsigned lut)
{
    return ((((lut >> 4) ^ lut) & 0x0Fu) != 0) + ((((lut >> 2) ^ lut) & 0x33u) != 0) + ((((lut >> 1) ^ lut) & 0x55u) != 0);
}

/* The program as K2 wants it: the non-constant nodes sorted by truth table (stable in node order, so the layout is
 * deterministic), the constants behind them.  Nodes with the same table form a RUN that K2 executes in a loop
 * specialised for that table; a warp is given a contiguous, cost-balanced piece of the sorted list, so the spare
 * upper halves of the descriptor words carry what scb_load_program needs to cut the list without a scan of its own:
 *   x >> 16  state rows (loads + store) moved by all earlier nodes of the list
 *   y >> 16  arity
 *   z >> 16  number of runs that start before this node's run
 * Stable counting sort: within a warp __match_any_sync ranks equal keys, the warps add their counts to the bins one
 * after the other.  All threads of the CTA call it (blockDim.x a multiple of 32). */
SCB_DEV void scb_scatter_by_table(const uint4 *sdv, const unsigned char *sarv, int N, uint4 *out, unsigned int *scr,
                                  int tid, int nthr, uint4 *outShared = nullptr)
{
    unsigned int *bins = scr, *cost = scr + SCB_SORT_BINS, *run = scr + 2 * SCB_SORT_BINS;
    unsigned short *rank = reinterpret_cast<unsigned short *>(scr + 3 * SCB_SORT_BINS);
    const int lane = tid & 31, warp = tid >> 5, W = nthr >> 5;
    for (int k = tid; k < SCB_SORT_BINS; k += nthr) bins[k] = 0u;
    __syncthreads();
    for (int base = 0; base < N; base += nthr) {
        const int i = base + tid;
        unsigned key = 0xFFFFu;                                   /* lanes beyond N: a key no node has */
        if (i < N) key = sarv[i] == 0 ? 256u + ((sdv[i].w >> 24) ? 1u : 0u) : (sdv[i].w >> 24);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs((int)peers) - 1;
        const unsigned before = (unsigned)__popc(peers & ((1u << lane) - 1u));
        for (int w = 0; w < W; ++w) {
            if (w == warp) {
                unsigned b = 0;
                if (lane == leader && i < N) { b = bins[key]; bins[key] = b + (unsigned)__popc(peers); }
                b = __shfl_sync(0xffffffffu, b, leader);
                if (i < N) rank[i] = (unsigned short)(b + before);
            }
            __syncthreads();
        }
    }
    if (warp == 0) {
        /* exclusive prefixes over the bins (node position, rows moved, runs): every lane sums its 9 consecutive bins,
         * one warp scan over the 32 partial sums, then the lane walks its bins again */
        const int PER = (SCB_SORT_BINS + 31) / 32;
        const int k0 = lane * PER;
        unsigned sn = 0, sc = 0, sr = 0;
        for (int k = k0; k < k0 + PER && k < SCB_SORT_BINS; ++k) {
            const unsigned n = bins[k];
            sn += n;
            if (k < 256) { sc += n * (unsigned)(scb_lut_arity((unsigned)k) + 1); sr += n != 0u ? 1u : 0u; }
        }
        unsigned pn = sn, pc = sc, pr = sr;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned tn = __shfl_sync(0xffffffffu, pn, (lane - o) & 31), tc = __shfl_sync(0xffffffffu, pc, (lane - o) & 31), tr = __shfl_sync(0xffffffffu, pr, (lane - o) & 31);
            if (lane >= o) { pn += tn; pc += tc; pr += tr; }
        }
        unsigned cN = pn - sn, cC = pc - sc, cR = pr - sr;
        for (int k = k0; k < k0 + PER && k < SCB_SORT_BINS; ++k) {
            const unsigned n = bins[k];
            const unsigned rows = k < 256 ? n * (unsigned)(scb_lut_arity((unsigned)k) + 1) : 0u;
            const unsigned isrun = (n != 0u && k < 256) ? 1u : 0u;
            bins[k] = cN; cost[k] = cC; run[k] = cR;
            cN += n; cC += rows; cR += isrun;
        }
    }
    __syncthreads();
    for (int i = tid; i < N; i += nthr) {
        uint4 d = sdv[i];
        const unsigned ar = sarv[i];
        const unsigned key = ar == 0 ? 256u + ((d.w >> 24) ? 1u : 0u) : (d.w >> 24);
        const unsigned r = rank[i];
        if (ar) {
            d.x |= (cost[key] + r * (ar + 1u)) << 16;
            d.y |= ar << 16;
            d.z |= run[key] << 16;
        }
        out[bins[key] + r] = d;
        if (outShared) outShared[bins[key] + r] = d;      /* the layout that follows reads the sorted program from here, not from L2 */
    }
    __syncthreads();
}

#if defined(SCB_TRACE) && defined(__CUDA_ARCH__)
#define SCB_K0_MARK(m) do { if (blockIdx.x == 0 && threadIdx.x == 0) A.diag->trace[(m)] = (unsigned)clock64(); } while (0)
#else
#define SCB_K0_MARK(m) do { } while (0)
#endif
__global__ void __launch_bounds__(256, 3) scb_k_compile(ScbCompileArgs A)
{
    SCB_DYN_SMEM(smraw);
    const int N = A.N, p = A.slots ? A.slots[blockIdx.x] : (int)blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    SCB_K0_MARK(0);
    if (tid == 0 && A.fitness) { A.fitness[p] = 0ull; A.cost[p] = 0u; A.cost[A.P + p] = 0u; }
    uint4 *sd = reinterpret_cast<uint4 *>(smraw);
    unsigned int *cn = reinterpret_cast<unsigned int *>(sd + N);          /* [0..3] class counts, [4..7] flag counts */
    unsigned char *sar = reinterpret_cast<unsigned char *>(cn + 8);
    const int32_t *ind = A.individuals + (long long)p * A.indLen;
    const long long toff = A.tablesPerIndividual ? (long long)p * N * (SCB_MAX_TERMS * SCB_MAX_IN) : 0;
    const int TBL = N * SCB_MAX_TERMS * SCB_MAX_IN;
    int32_t *an = reinterpret_cast<int32_t *>(smraw + (size_t)N * 16 + (size_t)((N + 15) / 16) * 16 + 64);
    int32_t *ai = an + TBL;
    if (tid < 8) cn[tid] = 0u;
    __syncthreads();
    if (A.compact) {
        /* the GA keeps only each node's upstream triple per individual: 24 bytes per node over the
         * bus instead of 168; the tables the reference would have marshalled are rebuilt here */
        const long long uoff = A.tablesPerIndividual ? (long long)p * N * SCB_MAX_IN : 0;
        for (int i = tid; i < N; i += nthr) {
            const int nt = scb_expand_terms(A.an + uoff + i * SCB_MAX_IN, A.ai + uoff + i * SCB_MAX_IN,
                                            an + i * (SCB_MAX_TERMS * SCB_MAX_IN), ai + i * (SCB_MAX_TERMS * SCB_MAX_IN));
            if (nt != A.andLen[i]) atomicAdd(&cn[7], 1u);        /* triple and andLenList disagree */
        }
    } else
    if ((TBL & 3) == 0) {                                     /* 16-byte copies when an individual's tables are 16-byte multiples */
        const int4 *gan = reinterpret_cast<const int4 *>(A.an + toff), *gai = reinterpret_cast<const int4 *>(A.ai + toff);
        int4 *san = reinterpret_cast<int4 *>(an), *sai = reinterpret_cast<int4 *>(ai);
        for (int i = tid; i < TBL / 4; i += nthr) { san[i] = gan[i]; sai[i] = gai[i]; }
    } else
    for (int i = tid; i < TBL; i += nthr) { an[i] = A.an[toff + i]; ai[i] = A.ai[toff + i]; }
    __syncthreads();
    SCB_K0_MARK(1);
    for (int i = tid; i < N; i += nthr) {
        const ScbNodeFn f = scb_compile_node(i, N, A.indLen, ind, A.andLen, A.parse, an, ai, A.ko + (long long)p * A.koStride, A.ki, A.sem);
        sd[i] = make_uint4((unsigned)f.in[0], (unsigned)f.in[1], (unsigned)f.in[2], (unsigned)i | (f.lut << 24));
        sar[i] = (unsigned char)f.arity;
        atomicAdd(&cn[3 - f.arity], 1u);
        if (f.flags & SCB_NF_UNSUPPORTED) atomicAdd(&cn[4], 1u);
        if (f.flags & SCB_NF_EMPTY) atomicAdd(&cn[5], 1u);
        if (f.flags & SCB_NF_BADIDX) atomicAdd(&cn[6], 1u);
        if (f.flags & SCB_NF_TERMS) atomicAdd(&cn[7], 1u);
    }
    __syncthreads();
    SCB_K0_MARK(2);
    const int c3 = (int)cn[0], c2 = (int)cn[1], c1 = (int)cn[2], c0 = (int)cn[3];
    unsigned int *wc = reinterpret_cast<unsigned int *>(smraw + scb_compile_smem_bytes(N) - scb_sort_smem_bytes(N));
    /* the images the simulator copies into its shared memory: image 0 = the full program, 1 + k = reduced program k */
    int *lscr = reinterpret_cast<int *>(smraw + scb_compile_smem_bytes(N) - 96 * 4 - (size_t)N * 16);
    uint4 *ssort = reinterpret_cast<uint4 *>(smraw + scb_compile_smem_bytes(N) - (size_t)N * 16);
    uint4 *imgs = A.imgs ? A.imgs + (long long)p * (SCB_NPROG + 1) * A.imgStride : nullptr;
    scb_scatter_by_table(sd, sar, N, A.descs + (long long)p * A.Npad, wc, tid, nthr, ssort);
    SCB_K0_MARK(3);
    if (imgs) scb_layout_program(imgs, lscr, ssort, make_int4(c3, c2, c1, c0), N, A.imgWarps, A.imgRowBytes, tid, nthr);
    SCB_K0_MARK(4);
    /* ---- reduced programs by constant propagation.  c(t) = nodes whose value in S_t is the same constant for every
     * cell: c(1) = the constant rules (knock-outs, knock-ins, constant-TRUE literals), c(t+1) = rules that are
     * constant once the inputs in c(t) are fixed.  Knowledge only grows, so c(D) stays valid for every t >= D.  A
     * program restricted by c(D) (constant inputs folded into the tables, constant nodes no longer computed) is
     * exact for every step from D + 2 on: then both state buffers of K2 hold the constant rows.  The always-knocked
     * node 0 and its zero-input copies (Q2) erase most AND-terms of a typical network, a few more with
     * every round.  Programs are emitted at the fixed depths SCB_STAGE_DEPTHS and at the fixed point (or the cap).
     * The dead term tables' shared memory is reused. */
    unsigned char *cst = reinterpret_cast<unsigned char *>(an);
    unsigned char *pend = cst + ((N + 15) / 16) * 16;
    unsigned char *sar2 = pend + ((N + 15) / 16) * 16;
    uint4 *sd2 = reinterpret_cast<uint4 *>(sar2 + ((N + 15) / 16) * 16);
    int *chg = reinterpret_cast<int *>(sd2 + N);                 /* [0] changed, [1..4] class counts */
    const int stageDepth[SCB_NPROG - 1] = SCB_STAGE_DEPTHS;
    uint4 *progs = A.progs + (long long)p * SCB_NPROG * (N + 1);
    unsigned char *progFrom = A.progFrom + (long long)p * 8;
    for (int i = tid; i < N; i += nthr) cst[i] = sar[i] == 0 ? ((sd[i].w >> 24) ? 1 : 0) : 2;
    if (tid < 8) { chg[tid] = 0; progFrom[tid] = 0; }
    __syncthreads();
    int D = 1, stage = 0;
    for (;;) {
        /* knowledge here is c(D) */
        bool last = D >= SCB_STEADY_MAX_DEPTH;
        if (!last) {                                                       /* c(D + 1) from c(D) */
            int grew = 0;
            for (int i = tid; i < N; i += nthr) {
                unsigned char v = 2;
                if (cst[i] == 2) {
                    const uint4 d = sd[i];
                    ScbNodeFn f;
                    f.arity = sar[i]; f.in[0] = (int)d.x; f.in[1] = (int)d.y; f.in[2] = (int)d.z;
                    f.lut = d.w >> 24; f.flags = 0;
                    const unsigned rl = scb_restrict_lut(f, cst);
                    if (rl == 0u || rl == 0xFFu) v = rl ? 1 : 0;
                }
                pend[i] = v;                                               /* read back by this thread only */
                grew |= v != 2;
            }
            last = __syncthreads_or(grew) == 0;                            /* fixed point: c(D + 1) == c(D) */
        }
        const bool emitStage = !last && stage < SCB_NPROG - 1 && D == stageDepth[stage];
        if (last || emitStage) {
            /* the program restricted by c(D) */
            if (tid < 4) chg[1 + tid] = 0;
            __syncthreads();
            for (int i = tid; i < N; i += nthr) {
                const uint4 d = sd[i];
                ScbNodeFn r;
                if (cst[i] != 2) r = scb_const_fn(cst[i], 0);
                else {
                    ScbNodeFn f;
                    f.arity = sar[i]; f.in[0] = (int)d.x; f.in[1] = (int)d.y; f.in[2] = (int)d.z;
                    f.lut = d.w >> 24; f.flags = 0;
                    r = scb_restrict_fn(f, cst);
                    /* constant under c(D) but not in c(D): constant only from S_(D+1) on -- such a node keeps its
                     * full rule in this program (there are none at the fixed point) */
                    if (r.arity == 0) r = f;
                }
                sd2[i] = make_uint4((unsigned)r.in[0], (unsigned)r.in[1], (unsigned)r.in[2], (unsigned)i | (r.lut << 24));
                sar2[i] = (unsigned char)r.arity;
                atomicAdd(reinterpret_cast<unsigned int *>(&chg[1 + 3 - r.arity]), 1u);
            }
            __syncthreads();
            const int s3 = chg[1], s2 = chg[2], s1 = chg[3], s0 = chg[4];
            const int slot = last ? SCB_NPROG - 1 : stage;
            uint4 *out2 = progs + (long long)slot * (N + 1);
            SCB_K0_MARK(8 + 4 * slot);
            scb_scatter_by_table(sd2, sar2, N, out2, wc, tid, nthr, ssort);
            SCB_K0_MARK(9 + 4 * slot);
            if (imgs) scb_layout_program(imgs + (long long)(1 + slot) * A.imgStride, lscr, ssort, make_int4(s3, s2, s1, s0), N, A.imgWarps, A.imgRowBytes, tid, nthr);
            SCB_K0_MARK(10 + 4 * slot);
            if (tid == 0) {
                out2[N] = make_uint4((unsigned)s3, (unsigned)s2, (unsigned)s1, (unsigned)s0);   /* the counts travel with the program */
                progFrom[slot] = (unsigned char)(D + 2);
                if (last && A.pred) A.pred[p] = (unsigned)(4 * s3 + 3 * s2 + 2 * s1);
            }
            if (emitStage) ++stage;
        }
        if (last) break;
        /* every thread is past a barrier that follows the reads of cst above */
        for (int i = tid; i < N; i += nthr) if (pend[i] != 2) cst[i] = pend[i];
        ++D;
        __syncthreads();
    }
    SCB_K0_MARK(5);
#if defined(SCB_TRACE) && defined(__CUDA_ARCH__)
    if (blockIdx.x == 0 && tid == 0) { A.diag->trace[6] = (unsigned)D; A.diag->trace[7] = (unsigned)nthr; }
#endif
    if (tid == 0) {
        A.counts[p] = make_int4(c3, c2, c1, c0);
        if (cn[4]) atomicAdd(&A.diag->unsupported_nodes, (unsigned long long)cn[4]);
        if (cn[5]) atomicAdd(&A.diag->empty_rule_nodes, (unsigned long long)cn[5]);
        if (cn[6]) atomicAdd(&A.diag->bad_index_nodes, (unsigned long long)cn[6]);
        if (cn[7]) atomicAdd(&A.diag->term_overflow_nodes, (unsigned long long)cn[7]);
    }
}

/* order[rank] = individual, by predicted cost, costliest first (ties by index): one CTA, every thread ranks its
 * individuals against all others (P <= SCB_ORDER_MAX_P values in shared memory) */
#define SCB_ORDER_MAX_P 4096
__global__ void scb_k_order(const unsigned int *pred, int32_t *order, int P)
{
    SCB_DYN_SMEM(smraw);
    unsigned int *sp = reinterpret_cast<unsigned int *>(smraw);
    for (int i = threadIdx.x; i < P; i += blockDim.x) sp[i] = pred[i];
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const unsigned int v = sp[i];
        int rank = 0;
        for (int j = 0; j < P; ++j) rank += (sp[j] > v || (sp[j] == v && j < i)) ? 1 : 0;
        order[rank] = i;
    }
}

/* rows of a resident population replaced in place: row j of `src` ([n][rowInts]) becomes row slots[j] of `dst` */
__global__ void scb_k_scatter_rows(int32_t *dst, const int32_t *src, const int32_t *slots, int rowInts, int n)
{
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        int32_t *d = dst + (long long)slots[j] * rowInts;
        const int32_t *s = src + (long long)j * rowInts;
        for (int i = threadIdx.x; i < rowInts; i += blockDim.x) d[i] = s[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * K1: bit-pack.  grid = (ceil(Wpad/32/warpsPerBlock), N); a warp produces 32 consecutive
 * words of one row: 32 coalesced 32-element reads, one ballot each.
 * ---------------------------------------------------------------------------------------- */
__global__ void scb_k_pack(ScbPackArgs A)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.y;
    const int w0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32;
    if (w0 >= A.Wpad) return;
    const int32_t *s4 = reinterpret_cast<const int32_t *>(A.src) + (long long)row * A.rowStride;
    const unsigned char *s1 = reinterpret_cast<const unsigned char *>(A.src) + (long long)row * A.rowStride;
    uint32_t mine = 0;
    unsigned bad = 0;
    for (int k = 0; k < 32; ++k) {
        const long long c = (long long)(w0 + k) * 32 + lane;
        int v = 0;
        if (c < A.C) v = (A.elemBytes == 4) ? s4[c] : (int)s1[c];
        bad += (v != 0 && v != 1);
        const unsigned b = __ballot_sync(0xffffffffu, v != 0);
        if (k == lane) mine = b;
    }
    if (w0 + lane < A.Wpad) A.X[(long long)row * A.Wpad + w0 + lane] = mine;
    bad = scb_warp_sum(bad);
    if (lane == 0 && bad) atomicAdd(&A.diag->nonbinary_cells, (unsigned long long)bad);
}

/* K1b: bit-pack straight from a CSR matrix (the reference keeps binMat as scipy CSR and densifies it on every
 * evaluation, ruleMaker.py:1109-1111).  One thread per stored element of the selected rows: the bit of
 * (row i, column c) is OR-ed into X[i][c/32].  X must be zeroed first. */
struct ScbPackCsrArgs {
    const int32_t *rowOfNnz;      /* [nnz] node index of each stored element */
    const int32_t *col;           /* [nnz] */
    const double *val;            /* [nnz] or null (structure only: stored => 1) */
    long long nnz;
    int C, Wpad;
    uint32_t *X;
    ScbDiag *diag;
};

__global__ void scb_k_pack_csr(ScbPackCsrArgs A)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.nnz) return;
    const int c = A.col[e];
    const double v = A.val ? A.val[e] : 1.0;
    if (c < 0 || c >= A.C) return;                  /* columns beyond lenSamples are not part of the data set */
    if (v != 0.0 && v != 1.0) { atomicAdd(&A.diag->nonbinary_cells, 1ull); return; }
    if (v == 1.0) atomicOr(&A.X[(long long)A.rowOfNnz[e] * A.Wpad + (c >> 5)], 1u << (c & 31));
}

#endif  /* SCB_SIM_ONLY */

/* ------------------------------------------------------------------------------------------
 * shared pieces of K2 / K3
 * ---------------------------------------------------------------------------------------- */
/* the constant rows of a program (entries e1 .. e1 + n0 of the GLOBAL program) into `dst` (lane-adjusted), and
 * into the global snapshot when zg != null */
template <int K>
SCB_DEV void scb_const_rows(scb_saddr dst, const uint4 *gconst, int n0, int warp, int S, unsigned char *zg)
{
    const uint32_t ROWB = 128u * K;
    for (int j = warp; j < n0; j += S) {
        const uint4 dc = gconst[j];
        ScbVec<K> r;
        const uint32_t val = (dc.w >> 24) ? 0xffffffffu : 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) r.v[k] = val;
        const uint32_t off = (dc.w & 0xFFFFFFu) * ROWB;
        scb_sts<K>(dst + off, r);
        if (zg) scb_stcg<K>(zg + off, r);
    }
}

/* load a state row unless the descriptor marks the input unused (sign bit of `off`) */
template <int K> SCB_DEV void scb_lds_if(ScbVec<K> &r, scb_saddr base, uint32_t off)
{
#if defined(__CUDA_ARCH__)
    const uint32_t a = base + off;
    if (K == 4) asm volatile("{\n.reg .pred p;\nsetp.ge.s32 p, %5, 0;\n@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
                             : "+r"(r.v[0]), "+r"(r.v[1 % K]), "+r"(r.v[2 % K]), "+r"(r.v[3 % K]) : "r"(a), "r"(off));
    else if (K == 2) asm volatile("{\n.reg .pred p;\nsetp.ge.s32 p, %3, 0;\n@p ld.shared.v2.u32 {%0,%1}, [%2];\n}"
                                  : "+r"(r.v[0]), "+r"(r.v[1 % K]) : "r"(a), "r"(off));
    else asm volatile("{\n.reg .pred p;\nsetp.ge.s32 p, %2, 0;\n@p ld.shared.u32 %0, [%1];\n}" : "+r"(r.v[0]) : "r"(a), "r"(off));
#else
    if ((int32_t)off >= 0) r = scb_ld<K>(base + off);
#endif
}

/* one synchronous update S_t -> S_{t+1} for the warp's share of the nodes (its list `wl` of `nw` entries, run
 * headers included), entry by entry with a table dispatch per node: the form the flagged steps use.
 *   src, dst : lane-adjusted addresses of the two state buffers
 *   zg       : lane-adjusted pointer into the global snapshot (ZCMP / ZSNAP without tensor memory)
 *   tsm      : lane-adjusted address of the shared-memory target state (TCMP)
 * d2 / dz accumulate, per word, the cells whose new state differs from S_{t-1} ...
 *
 * Software pipeline, two register sets: while node k is computed (table dispatch, LOP3, store), the rows of node
 * k + 1 are already in flight and the descriptor of node k + 3 is being fetched, so that a warp never waits for a
 * full shared-memory round trip per node (the first version issued descriptor -> rows -> table serially: 9 warp
 * cycles per instruction with 5 warps per scheduler). */
template <int K, bool PLAIN, bool TM = false, unsigned FIXED = 0u>
SCB_DEV void scb_step(scb_saddr src, scb_saddr dst, scb_saddr wl, int nw, unsigned flagsRt, unsigned char *zg, scb_saddr tsm,
                      ScbVec<K> &d2, ScbVec<K> &dz, bool active = true, uint32_t tm = SCB_TMEM_OFF)
{
    /* TM: the snapshot lives in tensor memory; tm is the address of this warp's first slot and the warp's k-th
     * list entry uses columns tm + K*k.  Otherwise it lives in global scratch (zg). */
    /* `active` is false for lanes whose words lie entirely beyond the last cell (the padding of the last
     * tile): they skip the step, which cuts the shared-memory wavefronts of a mostly empty tile.
     * PLAIN steps (no compare, no snapshot: the majority) are a separate instantiation so that their loop
     * carries no flag tests */
    /* FIXED != 0: the flags are a compile-time constant (the resolver's compare-with-target steps) */
    const unsigned flags = PLAIN ? 0u : (FIXED ? FIXED : flagsRt);
    /* lanes without valid cells sit out the plain steps; compare / snapshot steps keep the warp converged
     * because the tensor-memory instructions are warp-collective (their garbage is masked by the valid mask) */
    if (nw <= 0 || ((PLAIN || !TM) && !active)) return;
    uint32_t tmc = tm;
    ScbVec<K> A0, B0, C0, A1, B1, C1, Z0, Z1;
#pragma unroll
    for (int k = 0; k < K; ++k) { A0.v[k] = B0.v[k] = C0.v[k] = A1.v[k] = B1.v[k] = C1.v[k] = Z0.v[k] = Z1.v[k] = 0u; }
    uint32_t W0, W1 = 0u, X0, X1 = SCB_NOIN;
    const bool zglob = !TM && (flags & SCB_SF_ZCMP);
    /* rows (and the global snapshot row) of the node described by D into register set s */
#define SCB_ISSUE(s, D)                                                                           \
    do {                                                                                          \
        scb_lds_if<K>(A##s, src, (D).x); scb_lds_if<K>(B##s, src, (D).y); scb_lds_if<K>(C##s, src, (D).z); \
        W##s = (D).w; X##s = (D).x;                                                               \
        if (zglob && (int32_t)(D).x >= 0) Z##s = scb_ldcg<K>(zg + ((D).w & SCB_DST_MASK));        \
    } while (0)
    /* run headers (sign bit in x) are skipped: they take no tensor-memory slot either */
#define SCB_COMPUTE(s)                                                                            \
    if ((int32_t)X##s >= 0) {                                                                     \
        const uint32_t dsto = W##s & SCB_DST_MASK;                                                \
        ScbVec<K> o2, zc = Z##s;                                                                  \
        if (flags & SCB_SF_CMP2) o2 = scb_lds<K>(dst + dsto);                                     \
        if (flags & SCB_SF_TCMP) zc = scb_lds<K>(tsm + dsto);                                     \
        if (TM && (flags & SCB_SF_ZCMP)) scb_tmem_ld<K>(tmc, zc);                                 \
        const ScbVec<K> r = scb_apply3<K>(W##s >> 24, A##s, B##s, C##s);                          \
        if (flags & SCB_SF_CMP2) {                                                                \
            _Pragma("unroll") for (int k = 0; k < K; ++k) d2.v[k] |= r.v[k] ^ o2.v[k]; }          \
        if (flags & (SCB_SF_ZCMP | SCB_SF_TCMP)) {                                                \
            if (TM && (flags & SCB_SF_ZCMP)) scb_tmem_wait_ld<K>(zc);                             \
            _Pragma("unroll") for (int k = 0; k < K; ++k) dz.v[k] |= r.v[k] ^ zc.v[k]; }          \
        scb_sts<K>(dst + dsto, r);                                                                \
        if (flags & SCB_SF_ZSNAP) {                                                               \
            if (TM) scb_tmem_st<K>(tmc, r);                                                       \
            else scb_stcg<K>(zg + dsto, r);                                                       \
        }                                                                                         \
        if (TM) tmc += (uint32_t)K;                                                               \
    }

    uint4 D0 = scb_lds_desc(wl), D1 = scb_lds_desc(wl + 16);
    SCB_ISSUE(0, D0);
    D0 = scb_lds_desc(wl + 32);
    scb_saddr dp = wl;
    for (int k = 0;;) {
        SCB_ISSUE(1, D1);
        D1 = scb_lds_desc(dp + 48);
        SCB_COMPUTE(0);
        if (++k >= nw) break;
        SCB_ISSUE(0, D0);
        D0 = scb_lds_desc(dp + 64);
        SCB_COMPUTE(1);
        if (++k >= nw) break;
        dp += 32;
    }
    if (TM && (flags & SCB_SF_ZSNAP)) scb_tmem_wait_st();
#undef SCB_COMPUTE
#undef SCB_ISSUE
}

/* One step for the warp's list through the RUN interpreter (generated PTX, scb_lut_ptx.inc / gen_lut_ptx.py): one
 * table dispatch per run into a loop specialised for that table.  FLAGS = 0 (plain: most steps), SCB_SF_CMP2,
 * SCB_SF_ZCMP (snapshot / resolver target in tensor memory, list order) or both; the compare accumulators d2 / dz stay
 * in registers.  Lanes without valid cells sit out the steps without a tensor-memory compare; those keep the warp
 * converged (tcgen05.ld is warp-collective, the garbage of those lanes is masked by the valid mask). */
/* the head of a warp's list (first run header, first node entry): constant over the steps of a program, so the step
 * code gets it in registers instead of waiting for a shared-memory round trip at the start of every step */
struct ScbListHead { uint32_t h0, h1; uint4 d; };
SCB_DEV ScbListHead scb_list_head(const uint4 *list)
{
    ScbListHead r;
    r.h0 = list[0].x; r.h1 = list[0].y; r.d = list[1];
    return r;
}

template <int K, unsigned FLAGS>
SCB_DEV void scb_step_runs(scb_saddr src, scb_saddr dst, scb_saddr wl, const ScbListHead &hd, bool active, ScbVec<K> &d2, ScbVec<K> &dz, uint32_t tm)
{
    if (!(FLAGS & SCB_SF_ZCMP) && !active) return;
#if defined(__CUDA_ARCH__)
#define SCB_RUNS_IN "r"(wl), "r"(src), "r"(dst), "r"(tm), "r"(hd.h0), "r"(hd.h1), "r"(hd.d.x), "r"(hd.d.y), "r"(hd.d.z), "r"(hd.d.w)
#define SCB_RUNS_ASM(NAME)                                                                                             \
    do {                                                                                                               \
        if (K == 4) asm volatile(NAME##_K4 : "+r"(d2.v[0]), "+r"(d2.v[1 % K]), "+r"(d2.v[2 % K]), "+r"(d2.v[3 % K]),   \
                                 "+r"(dz.v[0]), "+r"(dz.v[1 % K]), "+r"(dz.v[2 % K]), "+r"(dz.v[3 % K])                \
                                 : SCB_RUNS_IN : "memory");                                                            \
        else if (K == 2) asm volatile(NAME##_K2 : "+r"(d2.v[0]), "+r"(d2.v[1 % K]), "+r"(dz.v[0]), "+r"(dz.v[1 % K])   \
                                      : SCB_RUNS_IN : "memory");                                                       \
        else asm volatile(NAME##_K1 : "+r"(d2.v[0]), "+r"(dz.v[0]) : SCB_RUNS_IN : "memory");                          \
    } while (0)
    if (FLAGS == 0u) SCB_RUNS_ASM(SCB_RUNS_PTX);
    else if (FLAGS == SCB_SF_CMP2) SCB_RUNS_ASM(SCB_RUNS_CMP2_PTX);
    else if (FLAGS == SCB_SF_ZCMP) SCB_RUNS_ASM(SCB_RUNS_ZCMP_PTX);
    else SCB_RUNS_ASM(SCB_RUNS_CMP2_ZCMP_PTX);
#undef SCB_RUNS_ASM
#undef SCB_RUNS_IN
#else
    const uint4 *e = reinterpret_cast<const uint4 *>(wl);
    uint32_t tmc = tm;
    for (;;) {
        const uint4 h = *e++;
        const unsigned cls = h.x & 0x1ffu, cnt = h.y & 0xffffu;
        if (cls == SCB_CLS_END) break;
        for (unsigned i = 0; i < cnt; ++i) {
            const uint4 d = *e++;
            if ((d.w >> 24) != cls) abort();                       /* a node in a run of another table */
            const ScbVec<K> a = scb_ld<K>(src + d.x);
            const ScbVec<K> b = (int32_t)d.y >= 0 ? scb_ld<K>(src + d.y) : a, c = (int32_t)d.z >= 0 ? scb_ld<K>(src + d.z) : a;
            ScbVec<K> r, o, t;
            if (FLAGS & SCB_SF_CMP2) o = scb_ld<K>(dst + (d.w & SCB_DST_MASK));
            if (FLAGS & SCB_SF_ZCMP) { scb_tmem_ld<K>(tmc, t); tmc += (uint32_t)K; }
            for (int k = 0; k < K; ++k) {
                r.v[k] = scb_lut_apply(cls, a.v[k], b.v[k], c.v[k]);
                if (FLAGS & SCB_SF_CMP2) d2.v[k] |= r.v[k] ^ o.v[k];
                if (FLAGS & SCB_SF_ZCMP) dz.v[k] |= r.v[k] ^ t.v[k];
            }
            scb_st<K>(dst + (d.w & SCB_DST_MASK), r);
        }
    }
#endif
}

/* index of the first set bit above `after` in the 128-bit mask (lo, hi); 1000 if there is none */
SCB_DEV int scb_next_bit(unsigned long long lo, unsigned long long hi, int after)
{
    int s = after + 1;
    if (s < 64) {
        const unsigned long long m = lo >> s;
        if (m) return s + __ffsll((long long)m) - 1;
        s = 64;
    }
    if (s < 128) {
        const unsigned long long m = hi >> (s - 64);
        if (m) return s + __ffsll((long long)m) - 1;
    }
    return 1000;
}

/* valid-cell mask of word w (cells w*32 .. w*32+31 below C) */
SCB_DEV uint32_t scb_valid_mask(long long w, int C)
{
    const long long lo = w * 32;
    if (lo >= C) return 0u;
    if (lo + 32 <= C) return 0xffffffffu;
    return (1u << (unsigned)(C - lo)) - 1u;
}

/* ------------------------------------------------------------------------------------------
 * K2: the simulator.  Persistent CTAs pull (individual, tile) items from an atomic counter.
 * A tile is 32*K words (= 1024*K cells) of every node.  Control flow is uniform per CTA.
 *
 * Early exit, all bit-exact:
 *   E2  S_t == S_{t-2} for every valid cell of the tile (period <= 2 from t-2 on):
 *       S_98 = S_t if 98-t is even, else S_{t-1}; all cells found.
 *   E3  S_t == snapshot S_s for every valid cell (tile period L | t-s from s on):
 *       S_98 is (98-t) mod L steps ahead; all cells found.
 *   R   otherwise cells are individually proven "found" by any repeat seen in a compare; once all
 *       are, the tile runs plain steps to S_98.  Cells still unproven after the last compare
 *       at t = 99 are queued, as 1024-cell chunks, for K3, and K2 leaves them out of its sums.
 * ---------------------------------------------------------------------------------------- */
template <int KR, bool TTM = false> SCB_DEV void scb_resolve_item(const ScbSimArgs &A, unsigned char *smraw, unsigned it, uint32_t tm = SCB_TMEM_OFF, bool useImg = false);

#ifndef SCB_SIM_MAX_WARPS
#define SCB_SIM_MAX_WARPS 16     /* 512 threads: 128 registers per thread.  At 640 threads (96 registers) the control code between
                                   * the steps spills, and with 229 KB of the SM's L1 given to shared memory a spill reload is
                                   * an L2 round trip (measured: 2.94 ms against 3.19 ms per C3 generation) */
#endif
#if defined(SCB_TRACE) && defined(__CUDA_ARCH__)
#define SCB_TRACE_MARK(step, which, fl)                                                                          \
    do {                                                                                                         \
        if (blockIdx.x == 0 && tilesDone == 0 && lane == 0 && (step) < 100) {                                    \
            A.diag->trace[((step) * 24 + warp) * 4 + (which)] = (unsigned)clock64();                             \
            if (warp == 0 && (which) == 0) A.diag->traceFlags[(step)] = (fl);                                    \
        }                                                                                                        \
    } while (0)
#else
#define SCB_TRACE_MARK(step, which, fl) do { } while (0)
#endif
#if defined(SCB_TRACE) && defined(__CUDA_ARCH__)
SCB_DEV unsigned scb_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (unsigned)t; }
#define SCB_EV_LOG(kind, t0, extra)                                                                              \
    do {                                                                                                         \
        if (threadIdx.x == 0) {                                                                                  \
            const unsigned e_ = atomicAdd(&A.diag->evCount, 1u);                                                 \
            if (e_ < 16384u) {                                                                                   \
                A.diag->evLog[e_ * 4 + 0] = (unsigned)(kind) | (blockIdx.x << 8);                                \
                A.diag->evLog[e_ * 4 + 1] = (t0); A.diag->evLog[e_ * 4 + 2] = scb_gtime();                       \
                A.diag->evLog[e_ * 4 + 3] = (extra);                                                             \
            }                                                                                                    \
        }                                                                                                        \
    } while (0)
#define SCB_EV_T0(name) const unsigned name = scb_gtime()
#else
#define SCB_EV_LOG(kind, t0, extra) do { } while (0)
#define SCB_EV_T0(name) do { } while (0)
#endif
template <int K, bool TM>
__global__ void __launch_bounds__(SCB_SIM_MAX_WARPS * 32, 1) scb_k_sim(ScbSimArgs A)
{
    SCB_DYN_SMEM(smraw);
    const int N = A.N;
    const uint32_t ROWB = 128u * K, BUFB = (uint32_t)N * ROWB;
    const int WT = 32 * K;                              /* words per tile */
    unsigned char *sm = smraw;
    uint4 *desc = reinterpret_cast<uint4 *>(smraw + 2u * BUFB);
    /* one program image; the next program waits in registers (fetched right after a switch, stored at the next one) */
    const uint32_t IMGB = (uint32_t)scb_img_entries(N, (int)(blockDim.x >> 5)) * 16u;
    /* compare accumulators {d2[WT], dz[WT]} (word lane * K + k of the tile sits at [k * 32 + lane]: a warp's atomics hit
     * 32 different banks), then scratch of the epilogue: [2 * WT ...) the error column of each chunk's last valid cell */
    uint32_t *cm = reinterpret_cast<uint32_t *>(smraw + 2u * BUFB + IMGB);
    uint32_t *vm = cm + 6 * WT;
    uint32_t *Rm = vm + WT;                              /* cells a repeat has been seen for (found) */
    int *sh = reinterpret_cast<int *>(Rm + WT);          /* [0] item, [4..4+K) chunk flags, [20..28) compare flags of the folding warps */
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31;
    int warp = tid >> 5, S = nthr >> 5;
    uint32_t laneOff = (uint32_t)lane * 4u * K;

    unsigned char *zg = reinterpret_cast<unsigned char *>(A.zscratch + (unsigned long long)blockIdx.x * A.zstride) + laneOff;
    scb_saddr sa = scb_smem_addr(sm), sdesc = scb_smem_addr(desc);
#if defined(__CUDA_ARCH__)
    /* opaque to the compiler from here on: otherwise it re-derives these from the special registers (S2R tid,
     * SR_CgaCtaId, the shared window base) in every step instead of keeping five registers (measured: +14 % instructions) */
    asm volatile("" : "+r"(warp), "+r"(S), "+r"(laneOff), "+r"(sa), "+r"(sdesc));
#endif
    /* the steady program must be in use before the first snapshot is taken: the tensor-memory slots of a snapshot
     * follow the program's node order */
    /* snapshot in tensor memory: warp 0 allocates the columns, every warp derives the address of its slots */
    uint32_t tm = SCB_TMEM_OFF;
    uint32_t tmemBase = 0u;
    if (TM) {
#if defined(__CUDA_ARCH__)
        if (warp == 0) scb_tmem_alloc((uint32_t)__cvta_generic_to_shared(&sh[14]), (uint32_t)A.tmemCols);
        scb_tmem_fence_before();
        __syncthreads();
        scb_tmem_fence_after();
        tmemBase = (uint32_t)sh[14];
#endif
        tm = tmemBase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(K * (warp >> 2) * A.tmemSlotsPerWarp);
    }
#if defined(__CUDA_ARCH__)
    /* mbarrier of the TMA tile loads: 8 bytes in the reserved tail of the launch's dynamic shared memory */
    const uint32_t mbarAddr = (uint32_t)__cvta_generic_to_shared(smraw + A.smemBytes - 32);
    uint32_t tmaParity = 0u;
    if (A.useTma) {
        if (tid == 0) scb_mbar_init(mbarAddr, (uint32_t)(S < N ? S : N));
        __syncthreads();
    }
#endif
    const unsigned total = (unsigned)A.nEval * (unsigned)A.tiles;
    unsigned stepsDone = 0, tilesDone = 0, rowsMoved = 0;     /* per CTA: a few thousand steps, a few million rows */
    long long cycPro = 0, cycStep = 0, cycEpi = 0;

    /* ---- the work loop.  Thread 0 picks the next piece of work, in this order:
     *   1. a published resolver item: the longest piece of work there is (98 compare steps) and the launch ends when
     *      the last one does.  Claimed with a compare-and-swap on the cursor, so a CTA never holds a ticket for an item
     *      that does not exist yet.  The tile ticket the CTA holds at that moment is HANDED BACK (retItems): it is early
     *      in the cost order, i.e. expensive, and would otherwise wait for the whole item (measured: those tiles were
     *      the last to end, 250 us after the median CTA had left).
     *   2. a handed-back tile;
     *   3. the ticket it holds (requested while the previous tile ran: the atomic's latency is hidden), else a new one;
     *   4. nothing left: it waits for resolver items / handed-back tiles until every tile is done and both queues are
     *      empty (all producers are running CTAs of this launch, so waiting cannot deadlock). */
    int pending = -1;                                   /* thread 0: the ticket held */
    if (tid == 0) pending = (int)atomicAdd(A.workCounter, 1u);
    for (;;) {
        if (tid == 0) {
            int tileItem = -1, resItem = -1, done = 0;
            for (;;) {
                if (A.fuseResolve) {
                    const unsigned cur = scb_ld_volatile(A.resolveCursor);
                    unsigned cnt = scb_ld_volatile(A.resolveCount);
                    if (cnt > A.resolveCap) cnt = A.resolveCap;
                    if (cur < cnt && atomicCAS(A.resolveCursor, cur, cur + 1u) == cur) {
                        while (scb_ld_volatile(&A.readyFlags[cur]) == 0u) scb_backoff();      /* its producer is about to publish it */
                        __threadfence();
                        resItem = (int)cur;
                        if (pending >= 0 && (unsigned)pending < total) {
                            const unsigned r = atomicAdd(A.retCount, 1u);
                            atomicExch(&A.retItems[r], (unsigned)pending + 1u);
                            pending = -1;
                        }
                        break;
                    }
                    const unsigned rc = scb_ld_volatile(A.retCursor), rn = scb_ld_volatile(A.retCount);
                    if (rc < rn && atomicCAS(A.retCursor, rc, rc + 1u) == rc) {
                        unsigned v;
                        while ((v = scb_ld_volatile(&A.retItems[rc])) == 0u) scb_backoff();
                        tileItem = (int)(v - 1u);
                        break;
                    }
                }
                if (pending < 0) pending = (int)atomicAdd(A.workCounter, 1u);
                if ((unsigned)pending < total) { tileItem = pending; pending = -1; break; }
                if (!A.fuseResolve) { done = 1; break; }
                if (scb_ld_volatile(A.tilesDone) >= total) {
                    __threadfence();
                    unsigned cnt = scb_ld_volatile(A.resolveCount);
                    if (cnt > A.resolveCap) cnt = A.resolveCap;
                    if (scb_ld_volatile(A.resolveCursor) >= cnt && scb_ld_volatile(A.retCursor) >= scb_ld_volatile(A.retCount)) { done = 1; break; }
                }
                scb_backoff();
            }
            sh[12] = tileItem; sh[15] = resItem; sh[13] = done;
            /* The next ticket is requested now and looked at when this piece of work is over -- except among the last
             * gridDim tickets: a ticket held by a CTA that is busy with a long tile cannot be taken by one that has run
             * dry (measured at 300 tiles on 296 CTAs: the launch ended on second tiles queued behind 98-step ones). */
            if (tileItem >= 0 && pending < 0 && (unsigned)tileItem + gridDim.x < total) pending = (int)atomicAdd(A.workCounter, 1u);
        }
        __syncthreads();
        if (sh[13]) break;
        {
            const int it = sh[15];
            if (it >= 0) {
                SCB_EV_T0(evt0);
                if (TM && A.resolveGroup == K) scb_resolve_item<K, true>(A, smraw, (unsigned)it, tm, true);
                else if (A.resolveGroup == 2) scb_resolve_item<2>(A, smraw, (unsigned)it, SCB_TMEM_OFF, false);
                else scb_resolve_item<1>(A, smraw, (unsigned)it, SCB_TMEM_OFF, false);
                SCB_EV_LOG(1, evt0, (unsigned)it);
                __syncthreads();                            /* the resolver's scratch overlaps sh[] */
                continue;
            }
        }
        const unsigned item = (unsigned)sh[12];
        /* With a cost order (scb_pop_set_order) the list is individual-major, costliest individual first: classic
         * longest-processing-time scheduling, the launch ends on the cheapest individuals' short tiles.  Without one it
         * is tile-major: the long tiles of a hard individual are spread over the whole launch and the last items handed
         * out are the cheap, partly filled tail tiles. */
        int p, tile;
        if (A.order) { p = A.order[item / (unsigned)A.tiles]; tile = (int)(item % (unsigned)A.tiles); }
        else { p = (int)(item % (unsigned)A.nEval); tile = (int)(item / (unsigned)A.nEval); }
        const long long c0 = scb_clock();
        SCB_EV_T0(evTile0);
        const int w0 = A.wordBase + tile * WT;
        int4 cnt = A.counts[p];
        const uint4 *gfull = A.descs + (long long)p * A.Npad;
        const int e1full = cnt.x + cnt.y + cnt.z, n0full = cnt.w;      /* the constants of the FULL program are written at steps 1, 2 */
        /* K0 laid every program out as the image this CTA's warps consume: the full program is copied into the image
         * buffer now; the next reduced program (program k may compute steps >= from[k]) is fetched into registers
         * (`stage`, three uint4 per thread) right after a switch and stored over the buffer at the next one, between two
         * barriers.  All switches happen before the first snapshot, whose tensor-memory slots follow the program's
         * node order. */
        const uint4 *imgBase = A.imgs + (long long)p * (SCB_NPROG + 1) * A.imgStride;
        uint4 stage[SCB_STAGE_REGS];                        /* this thread's entries of the next program image */
        for (int j = tid; j < A.imgStride; j += nthr) scb_cp_async16(&desc[j], &imgBase[j]);
        scb_cp_async_commit();
        scb_cp_async_wait();
        __syncthreads();
        int nw, woff;
        scb_img_warp(desc, warp, S, woff, nw);
        scb_saddr wl = sdesc + 16u * (uint32_t)woff;
        ScbListHead hd = scb_list_head(desc + woff);
        unsigned long long fromBytes = A.useSteady ? *reinterpret_cast<const unsigned long long *>(A.progFrom + (long long)p * 8) : 0ull;
        int switchAt = 1000, nextProg = 0;
        unsigned rowsPerStep = (unsigned)(4 * cnt.x + 3 * cnt.y + 2 * cnt.z);
        const unsigned rowsAtTileStart = rowsMoved;
        int segStart = 0;
#define SCB_NEXT_PROGRAM(after)                                                                           \
        do {                                                                                              \
            switchAt = 1000;                                                                              \
            for (; nextProg < SCB_NPROG; ++nextProg) {                                                    \
                const int from = (int)((fromBytes >> (8 * nextProg)) & 0xffull);                          \
                if (from > (after) && from <= A.firstSnap) { switchAt = from; break; }                    \
            }                                                                                             \
            if (switchAt < 1000) {                                                                        \
                const uint4 *src = imgBase + (long long)(1 + nextProg) * A.imgStride;                     \
                _Pragma("unroll") for (int r = 0; r < SCB_STAGE_REGS; ++r) {                              \
                    const int j = tid + r * nthr;                                                         \
                    if (j < A.imgStride) stage[r] = src[j];                                               \
                }                                                                                         \
                ++nextProg;                                                                               \
            }                                                                                             \
        } while (0)
        SCB_NEXT_PROGRAM(0);
        bool tileByTma = false;
#if defined(__CUDA_ARCH__)
        if (A.useTma) {
            /* S_0 tile by TMA: one bulk copy per node row (a 128*K-byte segment of the packed matrix), issued by one
             * thread, landing in buffer A while the other threads load the program; completion on the mbarrier */
            tileByTma = true;
            if (lane == 0 && warp < N) {
                /* every warp's leader announces and issues the copies of its own rows (warp, warp + S, ...) */
                const int nrows = (N - warp + S - 1) / S;
                scb_fence_proxy_async();
                scb_mbar_expect_tx(mbarAddr, (uint32_t)nrows * ROWB);
                const uint32_t *xs = A.X + w0;
                for (int row = warp; row < N; row += S)
                    scb_bulk_g2s(sa + (uint32_t)row * ROWB, xs + (long long)row * A.Wpad, ROWB, mbarAddr);
            }
        }
#endif
        if (!tileByTma) {   /* S_0 tile: four rows in flight per warp */
            const uint32_t *xs = A.X + w0 + lane * K;
            int row = warp;
            for (; row + 3 * S < N; row += 4 * S) {
                const ScbVec<K> v0 = scb_ldg<K>(xs + (long long)row * A.Wpad);
                const ScbVec<K> v1 = scb_ldg<K>(xs + (long long)(row + S) * A.Wpad);
                const ScbVec<K> v2 = scb_ldg<K>(xs + (long long)(row + 2 * S) * A.Wpad);
                const ScbVec<K> v3 = scb_ldg<K>(xs + (long long)(row + 3 * S) * A.Wpad);
                scb_st<K>(sm + (uint32_t)row * ROWB + laneOff, v0);
                scb_st<K>(sm + (uint32_t)(row + S) * ROWB + laneOff, v1);
                scb_st<K>(sm + (uint32_t)(row + 2 * S) * ROWB + laneOff, v2);
                scb_st<K>(sm + (uint32_t)(row + 3 * S) * ROWB + laneOff, v3);
            }
            for (; row < N; row += S)
                scb_st<K>(sm + (uint32_t)row * ROWB + laneOff, scb_ldg<K>(xs + (long long)row * A.Wpad));
        }
        if (tid < WT) { vm[tid] = scb_valid_mask((long long)w0 + tid, A.C); Rm[tid] = 0u; }
        for (int i = tid; i < 6 * WT; i += nthr) cm[i] = 0u;
#if defined(__CUDA_ARCH__)
        if (tileByTma) { scb_mbar_wait(mbarAddr, tmaParity); tmaParity ^= 1u; }
#endif
        __syncthreads();

        int laneActiveI = ((long long)w0 + lane * K) * 32 < A.C ? 1 : 0;      /* the lane holds a valid cell */
#if defined(__CUDA_ARCH__)
        asm volatile("" : "+r"(laneActiveI));             /* kept in a register: otherwise the 64-bit test is redone in every step */
#endif
        const bool laneActive = laneActiveI != 0;
        const long long c1 = scb_clock();
        uint32_t srcOff = 0u, dstOff = BUFB, finOff = 0u;
        int t = 0, target = SCB_LAST, zs = -1, ncmp = 0, searchLeft = 0;
        bool allres = false, allfound = false, plainOnly = false, haveFin = false;
        /* The schedule of the flagged steps is static (host-built 128-bit masks: period-<=2 checks, snapshots, snapshot
         * compares); the tile only delays its first check.  Every warp runs this control code, so it is kept short. */
        unsigned long long chkLo = A.chkLo, chkHi = A.chkHi;
        {   /* no period-<=2 check before the constants have spread: until then the nodes they will reach still carry
             * shifting cell data, so the tile cannot have settled (the checks are an optimisation, results never depend
             * on when they run) */
            int fromK = A.chkFromProg >= 0 ? (int)((fromBytes >> (8 * A.chkFromProg)) & 0xffull) : 0;
            if (fromK == 0 && A.chkFromProg >= 0) fromK = (int)((fromBytes >> (8 * (SCB_NPROG - 1))) & 0xffull);   /* that depth was not reached */
            if (fromK > 0) {
                if (fromK >= 64) { chkLo = 0ull; chkHi &= ~((1ull << (fromK >= 99 ? 35 : fromK - 64)) - 1ull); }
                else chkLo &= ~((1ull << fromK) - 1ull);
            }
        }
        const unsigned long long evLo = chkLo | A.snapLo | A.zcmpLo, evHi = chkHi | A.snapHi | A.zcmpHi;
        int nextEvent = scb_next_bit(evLo, evHi, 0);        /* most steps are plain: the first step with a flag */
        while (t < target) {
            int tn = t + 1;                             /* this iteration computes S_tn */
            if (tn == switchAt) {
                /* the image waiting in registers replaces the program in use (every warp is past the barrier of the
                 * previous step, so nobody reads the old one any more) */
#pragma unroll
                for (int r = 0; r < SCB_STAGE_REGS; ++r) {
                    const int j = tid + r * nthr;
                    if (j < A.imgStride) desc[j] = stage[r];
                }
                __syncthreads();
                rowsMoved += (unsigned)(t - segStart) * rowsPerStep;
                segStart = t;
                const uint4 c4 = desc[0];
                cnt = make_int4((int)c4.x, (int)c4.y, (int)c4.z, (int)c4.w);
                rowsPerStep = (unsigned)(4 * cnt.x + 3 * cnt.y + 2 * cnt.z);
                scb_img_warp(desc, warp, S, woff, nw);
                wl = sdesc + 16u * (uint32_t)woff;
                hd = scb_list_head(desc + woff);
                SCB_NEXT_PROGRAM(tn);
            }
            ScbVec<K> d2, dz;
#pragma unroll
            for (int k = 0; k < K; ++k) { d2.v[k] = 0u; dz.v[k] = 0u; }
            if (tn <= 2) scb_const_rows<K>(sa + dstOff + laneOff, gfull + e1full, n0full, warp, S, nullptr);
            if (tn < nextEvent) {
                /* a run of plain steps: nothing but update, barrier, swap until the next event, program switch or
                 * the target (most of a tile's steps) */
                int until = nextEvent < switchAt ? nextEvent : switchAt;
                if (until > target + 1) until = target + 1;
                if (tn <= 2) until = tn + 1;            /* the constant rows go into both buffers */
#pragma unroll 1
                for (; tn < until; ++tn) {
                    SCB_TRACE_MARK(tn, 0, rowsPerStep << 16);
                    scb_step_runs<K, 0u>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, hd, laneActive, d2, dz, tm);
                    SCB_TRACE_MARK(tn, 1, 0u);
                    __syncthreads();
                    const uint32_t x = srcOff; srcOff = dstOff; dstOff = x;
                }
                t = until - 1;
                continue;
            }
            unsigned f = 0u;
            if (!allres) {
                /* a cell of period L matches the snapshot at every multiple of L, so the first compares after a
                 * snapshot can be skipped without losing any cell (the gap is part of the static mask) */
                const unsigned sh6 = (unsigned)tn & 63u;
                const unsigned long long c = tn < 64 ? chkLo : chkHi, z = tn < 64 ? A.zcmpLo : A.zcmpHi, sn = tn < 64 ? A.snapLo : A.snapHi;
                f = (((c >> sh6) & 1ull) ? SCB_SF_CMP2 : 0u) | (((z >> sh6) & 1ull) ? SCB_SF_ZCMP : 0u) | (((sn >> sh6) & 1ull) ? SCB_SF_ZSNAP : 0u);
            } else if (zs >= 0 && searchLeft > 0) {
                f = SCB_SF_ZCMP; --searchLeft;        /* all cells proven: look for a tile-wide period */
            }
            /* compare steps run through the interpreter too (snapshot in tensor memory); snapshot steps, the global
             * snapshot fallback and the odd step without a flag take the entry-by-entry form */
            SCB_TRACE_MARK(tn, 0, f | 0x100u | (rowsPerStep << 16));
            if (f == SCB_SF_CMP2) scb_step_runs<K, SCB_SF_CMP2>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, hd, laneActive, d2, dz, tm);
            else if (TM && f == SCB_SF_ZCMP) scb_step_runs<K, SCB_SF_ZCMP>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, hd, laneActive, d2, dz, tm);
            else if (TM && f == (SCB_SF_CMP2 | SCB_SF_ZCMP)) scb_step_runs<K, SCB_SF_CMP2 | SCB_SF_ZCMP>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, hd, laneActive, d2, dz, tm);
            else scb_step<K, false, TM>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, nw, f, zg, sa, d2, dz, laneActive, tm);
            SCB_TRACE_MARK(tn, 1, 0u);
            /* word lane * K + k of the tile is kept at [k * 32 + lane]: a warp's atomics hit 32 different banks */
            if (f & SCB_SF_CMP2) {
#pragma unroll
                for (int k = 0; k < K; ++k) if (d2.v[k]) atomicOr(&cm[k * 32 + lane], d2.v[k]);
            }
            if (f & SCB_SF_ZCMP) {
#pragma unroll
                for (int k = 0; k < K; ++k) if (dz.v[k]) atomicOr(&cm[WT + k * 32 + lane], dz.v[k]);
            }
            __syncthreads();
            SCB_TRACE_MARK(tn, 2, 0u);
            t = tn;
            { const uint32_t x = srcOff; srcOff = dstOff; dstOff = x; }   /* src = S_t, dst = S_{t-1} */
            if (f & (SCB_SF_CMP2 | SCB_SF_ZCMP)) {
                /* the first K warps fold one word per thread and vote; their flags go to a slot per warp and compare
                 * (double buffered by compare parity), so no atomics on one address */
                int *fslot = sh + 20 + 4 * (ncmp & 1);
                if (tid < WT) {
                    const uint32_t v = vm[tid];
                    const int slot = (tid % K) * 32 + tid / K;
                    const uint32_t m2 = (f & SCB_SF_CMP2) ? cm[slot] : 0xffffffffu;
                    const uint32_t mz = (f & SCB_SF_ZCMP) ? cm[WT + slot] : 0xffffffffu;
                    const uint32_t r = Rm[tid] | ~m2 | ~mz;    /* a repeat was seen for these cells */
                    Rm[tid] = r;
                    cm[slot] = 0u; cm[WT + slot] = 0u;
                    const unsigned b1 = __ballot_sync(0xffffffffu, (m2 & v) != 0u), b2 = __ballot_sync(0xffffffffu, (mz & v) != 0u);
                    const unsigned b4 = __ballot_sync(0xffffffffu, (~r & v) != 0u);
                    if (lane == 0) fslot[warp] = (int)((b1 ? 1u : 0u) | (b2 ? 2u : 0u) | (b4 ? 4u : 0u));
                }
                __syncthreads();
                unsigned fl = 0u;
#pragma unroll
                for (int k = 0; k < K; ++k) fl |= (unsigned)fslot[k];
                SCB_TRACE_MARK(tn, 3, 0u);
                ++ncmp;
                allres = (fl & 4u) == 0u;
                if ((f & SCB_SF_CMP2) && !(fl & 1u)) {
                    /* E2: S_t == S_{t-2} on every valid cell */
                    allfound = true; haveFin = true;
                    finOff = ((SCB_LAST - 1 - t) & 1) ? dstOff : srcOff;
                    break;
                }
                if ((f & SCB_SF_ZCMP) && !(fl & 2u) && t <= SCB_LAST - 1 && !plainOnly) {
                    /* E3: the whole tile is back at snapshot S_zs; period L divides t - zs */
                    const int L = t - zs;
                    target = t + (SCB_LAST - 1 - t) % L;
                    allfound = true; plainOnly = true;
                } else if (allres && target == SCB_LAST) {
                    target = SCB_LAST - 1;               /* S_99 is no longer needed */
                    searchLeft = A.periodSearch;
                }
            }
            if (f & SCB_SF_ZSNAP) zs = tn;
            if (plainOnly) nextEvent = 1000;
            else if (allres) nextEvent = (zs >= 0 && searchLeft > 0) ? tn + 1 : 1000;
            else nextEvent = scb_next_bit(evLo, evHi, tn);
        }
        /* loop ran out at t == target (98: src = S_98; 99: dst = S_98; E3: src = S_target == S_98),
         * or stopped early at t > target when the last cell resolved at t == 99 */
        if (!haveFin) finOff = (t == SCB_LAST) ? dstOff : srcOff;
        if (allres) allfound = true;
#if defined(SCB_TRACE) && defined(__CUDA_ARCH__)
        if (tid == 0) {
            atomicAdd(&A.diag->exitHist[t < 100 ? t : 99], 1u);
            atomicAdd(&A.diag->exitHist[haveFin ? 100 : plainOnly ? 101 : allfound ? 102 : 103], 1u);
        }
#endif
        stepsDone += (unsigned)t; ++tilesDone;
        rowsMoved += (unsigned)(t - segStart) * rowsPerStep;   /* state rows moved through shared memory (loads + stores) */
        /* cost of the tile in row units: its rows plus ~128 rows' worth of fixed time per step (barrier, dispatch) */
        if (A.cost && tid == 0) {
            /* cost of the tile in state-row units; a tile left to the resolver is followed by 98 compare steps of the
             * same width, which cost rather more than the tile's own */
            unsigned tc = rowsMoved - rowsAtTileStart + 128u * (unsigned)t;
            if (!allfound) tc += tc + (tc >> 2);
            atomicAdd(&A.cost[p], tc);
            atomicMax(&A.cost[A.P + p], tc);
        }
#undef SCB_NEXT_PROGRAM
        const long long c2 = scb_clock();

        if (A.mode == SCB_MODE_IMPORTANCE) {
            /* importanceScore adds rows res[0]..res[1] divided by the window length as integers (simulator.c:641-648):
             * only cells with a window of length 1 (S_98 == S_99) contribute.  The two state buffers hold S_99 and
             * S_98 (t == 99), or S_98 and S_97 of cells that are all proven periodic from step 97 or earlier, where
             * S_97 == S_98 <=> S_98 == S_99: fixed point <=> the buffers agree on every node.  Such a cell is found
             * (state 99 repeats state 98), so it needs no resolver; dmz keeps the mask for the epilogue. */
            ScbVec<K> df;
#pragma unroll
            for (int k = 0; k < K; ++k) df.v[k] = 0u;
            for (int row = warp; row < N; row += S) {
                const ScbVec<K> a = scb_ld<K>(sm + (uint32_t)row * ROWB + laneOff), b = scb_ld<K>(sm + BUFB + (uint32_t)row * ROWB + laneOff);
#pragma unroll
                for (int k = 0; k < K; ++k) df.v[k] |= a.v[k] ^ b.v[k];
            }
            uint32_t *dm2 = cm, *dmz = cm + WT;               /* word-indexed here; the compare buffers are no longer needed */
            __syncthreads();
            for (int i = tid; i < 2 * WT; i += nthr) cm[i] = 0u;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k) if (df.v[k]) atomicOr(&dm2[lane * K + k], df.v[k]);
            __syncthreads();
            if (tid < WT) {
                const uint32_t fp = ~dm2[tid] & vm[tid];
                dmz[tid] = fp; Rm[tid] |= fp;
            }
            __syncthreads();
        }

        /* ---- chunk flags: a 32-word chunk with an unproven valid cell goes to K3 ---------- */
        if (warp == 0) {
            bool bad = false;
            if (!allfound) {
#pragma unroll
                for (int k = 0; k < K; ++k) bad = bad || ((~Rm[lane * K + k] & vm[lane * K + k]) != 0u);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, bad);
            const int LPC = 32 / K;                          /* lanes per chunk */
            const int KR = A.resolveGroup;                   /* chunks per resolver item (1 or 2, <= K) */
            if (lane < K) {
                const unsigned msk = (LPC == 32) ? 0xffffffffu : (((1u << LPC) - 1u) << (lane * LPC));
                const int cb = (bal & msk) != 0u;
                const int q = w0 / 32 + lane;                  /* chunk index within the individual */
                sh[4 + lane] = cb;
                sh[8 + lane] = -1;
                if (!cb && q < A.chunksPerInd) {
                    /* every valid cell of the chunk is found: its last valid cell can feed a later fill-forward */
                    const long long rem = (long long)A.C - (long long)q * 1024;
                    A.chunkLast[(long long)p * A.chunksPerInd + q] = (int)(rem >= 1024 ? 1023 : rem - 1);
                }
            }
            __syncwarp();
            if (lane < K / KR) {
                /* one resolver item per group of KR adjacent chunks that holds an unproven chunk */
                unsigned bad = 0u;
                for (int j = 0; j < KR; ++j) if (sh[4 + lane * KR + j]) bad |= 1u << j;
                sh[16 + lane] = -1;
                if (bad) {
                    const unsigned idx = atomicAdd(A.resolveCount, 1u);
                    if (idx < A.resolveCap) {
                        A.resolveItems[idx] = make_uint2((unsigned)p, (unsigned)(w0 / 32 + lane * KR) | (bad << 24));
                        sh[16 + lane] = (int)idx;
                    }
                    if (idx < A.resolveTCap) for (int j = 0; j < KR; ++j) sh[8 + lane * KR + j] = (int)idx;
                }
            }
        }
        __syncthreads();
        /* hand S_99 of the queued groups to the resolver (t == 99 here, so src holds S_99) */
        if (!allfound) {
            const int GW = 32 * A.resolveGroup;               /* words per group */
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int w = lane * K + k, j = w / 32;
                const int slot = sh[8 + j];
                if (slot < 0) continue;
                uint32_t *T = A.resolveT + (unsigned long long)slot * N * GW + (w % GW);
                for (int row = warp; row < N; row += S)
                    T[(long long)row * GW] = *reinterpret_cast<const uint32_t *>(sm + srcOff + (uint32_t)row * ROWB + (uint32_t)w * 4u);
            }
            __syncthreads();          /* the cellwise epilogue reuses the S_99 buffer */
            /* publish the queued groups: everything a resolver needs (item, S_99) is in memory */
            if (A.fuseResolve && tid < K / A.resolveGroup && sh[16 + tid] >= 0) {
                __threadfence();
                atomicExch(&A.readyFlags[sh[16 + tid]], 1u);
            }
        }
        if (A.fuseResolve && tid == 0) { __threadfence(); atomicAdd(A.tilesDone, 1u); }

        if (A.mode == SCB_MODE_IMPORTANCE) {
            /* per node: fixed-point cells with the node set (every chunk is counted here; the resolver only
             * decides whether the remaining cells have a window at all) */
            const unsigned char *fin = sm + finOff + laneOff;
            ScbVec<K> fpm;
#pragma unroll
            for (int k = 0; k < K; ++k) fpm.v[k] = cm[WT + lane * K + k];
            for (int j = warp; j < N; j += S) {
                const ScbVec<K> s = scb_ld<K>(fin + (uint32_t)j * ROWB);
                unsigned c = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) c += (unsigned)__popc(s.v[k] & fpm.v[k]);
                c = scb_warp_sum(c);
                if (lane == 0 && c) atomicAdd(&A.nodewise[(long long)p * N + j], (int)c);
            }
            __syncthreads();
            cycPro += c1 - c0; cycStep += c2 - c1; cycEpi += scb_clock() - c2;
            continue;
        }
        /* ---- errors: e = S_98 xor S_0 on valid cells of proven chunks ---------------------- */
        ScbVec<K> ok;
#pragma unroll
        for (int k = 0; k < K; ++k) ok.v[k] = sh[4 + (lane * K + k) / 32] ? 0u : vm[lane * K + k];
        /* The error column of every proven chunk's last valid cell is kept (count, and node bits in the per-node mode):
         * a later chunk that starts with not-found cells repeats it (the reference's fill-forward), so the
         * fill-forward kernel has nothing to simulate.  A lane's K words lie in one chunk; lm marks the cell. */
        ScbVec<K> lm;
        const int myChunk = (lane * K) / 32;
        uint32_t *lastAcc = cm + 2 * WT;                    /* [K] counts, then [K][NB] node bits; zero since the tile began */
        const int NB = (N + 31) >> 5;
        {
            const int q = w0 / 32 + myChunk;
            const long long rem = (long long)A.C - (long long)q * 1024;
            const int lc = rem >= 1024 ? 1023 : (int)rem - 1;
            const bool have = q < A.chunksPerInd && !sh[4 + myChunk];
#pragma unroll
            for (int k = 0; k < K; ++k) lm.v[k] = (have && (lane * K + k) == myChunk * 32 + (lc >> 5)) ? 1u << (lc & 31) : 0u;
        }
        unsigned lastCnt = 0;
        const unsigned char *fin = sm + finOff + laneOff;
        unsigned char *oth = sm + (finOff ? 0u : BUFB) + laneOff;
        unsigned total32 = 0;
        constexpr int RF = 4;                                /* rows in flight per warp: the S_0 rows come from L2 */
        for (int j0 = warp; j0 < N; j0 += RF * S) {
            ScbVec<K> x[RF];
            int node[RF];
            bool live[RF];
#pragma unroll
            for (int r = 0; r < RF; ++r) {
                const int j = j0 + r * S;
                live[r] = j < N;
                node[r] = live[r] ? j : j0;
                x[r] = scb_ldg<K>(A.X + (long long)node[r] * A.Wpad + w0 + lane * K);
            }
#pragma unroll
            for (int r = 0; r < RF; ++r) {
                const uint32_t dsto = (uint32_t)node[r] * ROWB;
                const ScbVec<K> sv = scb_ld<K>(fin + dsto);
                ScbVec<K> e;
                unsigned c = 0;
                uint32_t l = 0u;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    e.v[k] = (sv.v[k] ^ x[r].v[k]) & ok.v[k];
                    c += (unsigned)__popc(e.v[k]);
                    l |= e.v[k] & lm.v[k];
                }
                if (!live[r]) { c = 0; l = 0u; }
                total32 += c;
                lastCnt += l ? 1u : 0u;
                if (A.mode == 1) {
                    if (l) atomicOr(&lastAcc[K + myChunk * NB + (node[r] >> 5)], 1u << (node[r] & 31));
                    c = scb_warp_sum(c);
                    if (lane == 0 && c) atomicAdd(&A.nodewise[(long long)p * N + node[r]], (int)c);
                } else if (A.mode == 2) {
                    if (live[r]) scb_st<K>(oth + dsto, e);
                }
            }
        }
        total32 = scb_warp_sum(total32);
        if (lane == 0 && total32) atomicAdd(&A.fitness[p], (unsigned long long)total32);
        if (lastCnt) atomicAdd(&lastAcc[myChunk], lastCnt);
        __syncthreads();
        if (tid < K && !sh[4 + tid] && w0 / 32 + tid < A.chunksPerInd) {
            A.chunkLastErr[(long long)p * A.chunksPerInd + w0 / 32 + tid] = (int)lastAcc[tid];
            lastAcc[tid] = 0u;
        }
        if (A.mode == 1)
            for (int i = tid; i < K * NB; i += nthr) {
                const int c = i / NB;
                if (!sh[4 + c] && w0 / 32 + c < A.chunksPerInd)
                    A.chunkLastBits[((long long)p * A.chunksPerInd + w0 / 32 + c) * NB + (i - c * NB)] = lastAcc[K + i];
                lastAcc[K + i] = 0u;
            }
        if (A.mode == 2) {
            const unsigned char *eb = sm + (finOff ? 0u : BUFB);
            for (int idx = tid; idx < WT * 32; idx += nthr) {
                const int w = idx >> 5, b = idx & 31;
                const long long c = ((long long)w0 + w) * 32 + b;
                if (c >= A.C || sh[4 + w / 32]) continue;
                int n1 = 0;
                for (int row = 0; row < N; ++row)
                    n1 += (int)((*reinterpret_cast<const uint32_t *>(eb + (uint32_t)row * ROWB + (uint32_t)w * 4u) >> b) & 1u);
                A.cellwise[(long long)p * A.C + c] = n1;
            }
        }
        __syncthreads();
        cycPro += c1 - c0; cycStep += c2 - c1; cycEpi += scb_clock() - c2;
        SCB_EV_LOG(0, evTile0, (unsigned)t | ((unsigned)p << 8) | (allfound ? 0u : 0x80000000u));
    }
#if defined(__CUDA_ARCH__)
    if (TM) {
        scb_tmem_fence_before();
        __syncthreads();
        if (warp == 0) scb_tmem_dealloc(tmemBase, (uint32_t)A.tmemCols);
    }
#endif
    if (tid == 0 && tilesDone) {
        atomicAdd(&A.diag->steps_executed, (unsigned long long)stepsDone);
        atomicAdd(&A.diag->rows_moved, (unsigned long long)rowsMoved);
        atomicAdd(&A.diag->tiles, (unsigned long long)tilesDone);
        atomicAdd(&A.diag->cyc_prologue, (unsigned long long)cycPro);
        atomicAdd(&A.diag->cyc_steps, (unsigned long long)cycStep);
        atomicAdd(&A.diag->cyc_epilogue, (unsigned long long)cycEpi);
    }
}

/* ------------------------------------------------------------------------------------------
 * K3: exact resolver for one 1024-cell chunk (32 words, K = 1), three shared-memory buffers
 * (A, B, T).  Restates the reference literally on packed words:
 *   pass a: 99 plain steps -> T = S_99
 *   pass b: found = OR_j (S_j == T), j = 0..98  (getAttractCycle2, simulator.c:306-331)
 *   e = S_98 xor S_0; found cells contribute e; a not-found cell repeats the previous found
 *   cell's e (the reference leaves error[]/totalError untouched, simulator.c:431-482, Q7),
 *   looking back into earlier chunks when the chunk starts with not-found cells.
 * ---------------------------------------------------------------------------------------- */
struct ScbResolveArgs {
    ScbSimArgs s;
    unsigned int *itemCursor;     /* persistent work counter of the launch */
};

/* simulate KR adjacent chunks (first chunk q) of the individual whose program is in `desc`.
 *   how = 0: pass a (99 plain steps -> T = S_99), then pass b (98 steps comparing every S_j with T)
 *   how = 1: T is read from Tglob ([N][32*KR] words, written by K2), pass b only
 *   how = 2: 98 plain steps only (Fm is not computed)
 * On return: buffer A rows hold S_98, buffer B rows hold e = (S_98 xor S_0) & valid for every node,
 * Fm[32*KR] = found & valid per word (how != 2), vm[32*KR] = valid masks.  All threads call it. */
/* program `which` (0 = full, 1 + k = reduced k) of individual p into the image buffer `desc`: K0's image when it was
 * cut for this CTA's geometry (the simulator's own launch), else laid out here from the table-sorted program.
 * All threads call it (barriers inside); returns the program's arity counts. */
template <int KR>
SCB_DEV int4 scb_fetch_program(const ScbSimArgs &A, uint4 *desc, int *scr, int p, int which, bool useImg, int S, int tid, int nthr)
{
    const int N = A.N;
    if (useImg) {
        const uint4 *src = A.imgs + ((long long)p * (SCB_NPROG + 1) + which) * A.imgStride;
        for (int j = tid; j < A.imgStride; j += nthr) desc[j] = src[j];
        __syncthreads();
        const uint4 c4 = desc[0];
        return make_int4((int)c4.x, (int)c4.y, (int)c4.z, (int)c4.w);
    }
    const uint4 *g = which == 0 ? A.descs + (long long)p * A.Npad : A.progs + ((long long)p * SCB_NPROG + (which - 1)) * (N + 1);
    int4 cnt;
    if (which == 0) cnt = A.counts[p];
    else { const uint4 c4 = g[N]; cnt = make_int4((int)c4.x, (int)c4.y, (int)c4.z, (int)c4.w); }
    scb_layout_program(desc, scr, g, cnt, N, S, 128u * KR, tid, nthr);
    return cnt;
}

template <int KR, bool TTM = false>
SCB_DEV void scb_resolve_chunk(const ScbSimArgs &A, unsigned char *sm, uint4 *desc, int *winfo, bool useImg, int4 cnt, int p,
                               uint32_t *dmz, uint32_t *Fm, uint32_t *vm, int q, int how,
                               const uint32_t *Tglob, int tid, int nthr, uint32_t *D = nullptr,
                               uint32_t tm = SCB_TMEM_OFF)
{
    /* Like the simulator, the passes switch to K0's reduced programs as soon as each one is exact (program k from step
     * from[k] on: the constants it no longer computes are then in both state buffers, and they equal T's -- T = S_99
     * lies beyond every from[k] -- so leaving them out of the compare with T is exact too).  With T in tensor memory
     * its slots follow the program's node order and are re-laid out from Tglob at every switch (how == 1 only). */
    const int4 cntFull = cnt;
    const unsigned long long fromBytes = (A.useSteady && !(TTM && how == 0))
        ? *reinterpret_cast<const unsigned long long *>(A.progFrom + (long long)p * 8) : 0ull;
    bool replaced = false;
    /* TTM: the target state T = S_99 is kept in tensor memory (this warp's slots start at tm, one slot of KR
     * columns per non-constant node in the order scb_step visits them) instead of a third shared-memory buffer,
     * so that a whole 4-chunk tile fits; the per-step fold then goes through dmz with one extra barrier. */
    /* D: [2][warps][32*KR] per-step difference words of every warp (double buffered by step parity), so that the
     * per-step "equal to T" reduction needs no atomics and no barrier of its own.
     */
    const int N = A.N;
    const uint32_t ROWB = 128u * KR, BUFB = (uint32_t)N * ROWB;
    const int GW = 32 * KR;
    const int lane = tid & 31, warp = tid >> 5, S = nthr >> 5;
    const uint32_t laneOff = (uint32_t)lane * 4u * KR;
    const int w0 = q * 32;
    unsigned char *bufA = sm, *bufB = sm + BUFB, *bufT = sm + 2u * BUFB;
    const scb_saddr sa = scb_smem_addr(sm), sdesc = scb_smem_addr(desc);
    ScbVec<KR> d2, dz;
    const uint4 *gfull = A.descs + (long long)p * A.Npad;
    const int e1full = cntFull.x + cntFull.y + cntFull.z, n0full = cntFull.w;
    int woff, nw;                                                 /* this warp's list: offset, entries (run headers included) */
    scb_img_warp(desc, warp, S, woff, nw);
    const uint4 *wlp = desc + woff;
    scb_saddr wl = sdesc + 16u * (uint32_t)woff;

    for (int pass = (how == 0 ? 0 : 1); pass < 2; ++pass) {
        const bool cmp = pass == 1 && how != 2;
        if (replaced) {                       /* the previous pass ended on a reduced program: start again with the full one */
            cnt = scb_fetch_program<KR>(A, desc, winfo, p, 0, useImg, S, tid, nthr);
            scb_img_warp(desc, warp, S, woff, nw);
            wlp = desc + woff; wl = sdesc + 16u * (uint32_t)woff;
            replaced = false;
        }
        int nextProg = 0, switchAt = 1000;
        for (; nextProg < SCB_NPROG; ++nextProg) {
            const int from = (int)((fromBytes >> (8 * nextProg)) & 0xffull);
            if (from > 0) { switchAt = from; break; }
        }
        for (int row = warp; row < N; row += S) {
            const ScbVec<KR> v = scb_ldg<KR>(A.X + (long long)row * A.Wpad + w0 + lane * KR);
            scb_st<KR>(bufA + (uint32_t)row * ROWB + laneOff, v);
            if (!TTM && pass == 1 && how == 1)
                scb_st<KR>(bufT + (uint32_t)row * ROWB + laneOff,
                           scb_ld<KR>(reinterpret_cast<const unsigned char *>(Tglob + (long long)row * GW + lane * KR)));
        }
        if (tid < GW) { vm[tid] = scb_valid_mask((long long)w0 + tid, A.C); dmz[tid] = 0u; dmz[GW + tid] = 0u; Fm[tid] = 0u; }
        if (TTM && pass == 1 && how == 1) {
            /* T from the simulator's hand-over buffer into this warp's tensor-memory slots */
            uint32_t tmc = tm;
            for (int k = 0; k < nw; ++k) {
                if ((int32_t)wlp[k].x < 0) continue;               /* run header */
                const int node = (int)((wlp[k].w & SCB_DST_MASK) / ROWB);
                scb_tmem_st<KR>(tmc, scb_ld<KR>(reinterpret_cast<const unsigned char *>(Tglob + (long long)node * GW + lane * KR)));
                tmc += (uint32_t)KR;
            }
            scb_tmem_wait_st();
        }
        __syncthreads();
        if (cmp) {
            /* j = 0: compare S_0 with T on every row */
            ScbVec<KR> d;
#pragma unroll
            for (int k = 0; k < KR; ++k) d.v[k] = 0u;
            if (TTM) {
                uint32_t tmc = tm;
                for (int i = 0; i < nw; ++i) {
                    if ((int32_t)wlp[i].x < 0) continue;           /* run header */
                    ScbVec<KR> y;
                    scb_tmem_ld<KR>(tmc, y);
                    tmc += (uint32_t)KR;
                    const ScbVec<KR> x = scb_ld<KR>(bufA + (wlp[i].w & SCB_DST_MASK) + laneOff);
                    scb_tmem_wait_ld<KR>(y);
#pragma unroll
                    for (int k = 0; k < KR; ++k) d.v[k] |= x.v[k] ^ y.v[k];
                }
                for (int j = warp; j < n0full; j += S) {            /* constant rows (the full program is loaded): T holds the constant */
                    const uint4 dc = gfull[e1full + j];
                    const ScbVec<KR> x = scb_ld<KR>(bufA + (dc.w & 0xFFFFFFu) * ROWB + laneOff);
                    const uint32_t cv = (dc.w >> 24) ? 0xffffffffu : 0u;
#pragma unroll
                    for (int k = 0; k < KR; ++k) d.v[k] |= x.v[k] ^ cv;
                }
            } else
            for (int row = warp; row < N; row += S) {
                const ScbVec<KR> x = scb_ld<KR>(bufA + (uint32_t)row * ROWB + laneOff), y = scb_ld<KR>(bufT + (uint32_t)row * ROWB + laneOff);
#pragma unroll
                for (int k = 0; k < KR; ++k) d.v[k] |= x.v[k] ^ y.v[k];
            }
#pragma unroll
            for (int k = 0; k < KR; ++k) if (d.v[k]) atomicOr(&dmz[lane * KR + k], d.v[k]);
            __syncthreads();
            if (tid < GW) { Fm[tid] |= ~dmz[tid]; dmz[tid] = 0u; }
            __syncthreads();
        }
        const int last = pass == 0 ? SCB_LAST : SCB_LAST - 1;
        uint32_t srcOff = 0u, dstOff = BUFB;
        for (int tn = 1; tn <= last; ++tn) {
            if (tn == switchAt) {
                /* every warp is past the barrier that ended step tn - 1: the program can be replaced in place */
                cnt = scb_fetch_program<KR>(A, desc, winfo, p, 1 + nextProg, useImg, S, tid, nthr);
                scb_img_warp(desc, warp, S, woff, nw);
                wlp = desc + woff; wl = sdesc + 16u * (uint32_t)woff;
                replaced = true;
                if (TTM && cmp) {
                    /* T into this warp's tensor-memory slots in the new program's node order */
                    uint32_t tmc = tm;
                    for (int k = 0; k < nw; ++k) {
                        if ((int32_t)wlp[k].x < 0) continue;       /* run header */
                        const int node = (int)((wlp[k].w & SCB_DST_MASK) / ROWB);
                        scb_tmem_st<KR>(tmc, scb_ld<KR>(reinterpret_cast<const unsigned char *>(Tglob + (long long)node * GW + lane * KR)));
                        tmc += (uint32_t)KR;
                    }
                    scb_tmem_wait_st();
                }
                const int prev = switchAt;
                switchAt = 1000;
                for (++nextProg; nextProg < SCB_NPROG; ++nextProg) {
                    const int from = (int)((fromBytes >> (8 * nextProg)) & 0xffull);
                    if (from > prev) { switchAt = from; break; }
                }
            }
            unsigned f = 0u;
            if (cmp) f |= TTM ? SCB_SF_ZCMP : SCB_SF_TCMP;
#pragma unroll
            for (int k = 0; k < KR; ++k) { d2.v[k] = 0u; dz.v[k] = 0u; }
            const bool act = true;
            /* the constant rows of the full program (it is the one in use at steps 1, 2): they equal T's, so the
             * compare leaves them out */
            if (tn <= 2) scb_const_rows<KR>(sa + dstOff + laneOff, gfull + e1full, n0full, warp, S, nullptr);
            if (TTM && cmp) {
                /* compare with T in tensor memory, fold through dmz */
                /* One barrier per step: the accumulator of step tn is folded (and cleared) by the first GW threads after
                 * the step's barrier while the other warps already add step tn + 1 to the other accumulator; it is used
                 * again at step tn + 2, which every warp enters through the barrier of tn + 1, after the fold. */
                scb_step_runs<KR, SCB_SF_ZCMP>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, scb_list_head(wlp), true, d2, dz, tm);
                uint32_t *acc = dmz + (tn & 1) * GW;
#pragma unroll
                for (int k = 0; k < KR; ++k) if (dz.v[k]) atomicOr(&acc[k * 32 + lane], dz.v[k]);
                __syncthreads();
                if (tid < GW) { uint32_t *a = &acc[(tid % KR) * 32 + tid / KR]; Fm[tid] |= ~*a; *a = 0u; }   /* word tid = lane tid / KR, k = tid % KR */
                const uint32_t x = srcOff; srcOff = dstOff; dstOff = x;
                continue;
            }
            if (f == 0u && TTM) scb_step_runs<KR, 0u>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, scb_list_head(wlp), act, d2, dz, tm);
            else if (f == 0u) scb_step<KR, true>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, nw, 0u, nullptr, sa + 2u * BUFB + laneOff, d2, dz, act);
            else scb_step<KR, false, false, SCB_SF_TCMP>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, nw, f, nullptr, sa + 2u * BUFB + laneOff, d2, dz, act);
            if (cmp) {
                uint32_t *Dw = D + ((size_t)(tn & 1) * S + warp) * GW + lane * KR;
#pragma unroll
                for (int k = 0; k < KR; ++k) Dw[k] = dz.v[k];
            }
            __syncthreads();
            if (cmp && tid < GW) {
                /* these threads fold step tn while the others already run step tn + 1; buffer (tn & 1) is not
                 * written again before the barrier of step tn + 1, which these threads take part in */
                const uint32_t *Dr = D + (size_t)(tn & 1) * S * GW + tid;
                uint32_t acc = 0u;
                for (int w = 0; w < S; ++w) acc |= Dr[(size_t)w * GW];
                Fm[tid] |= ~acc;
            }
            const uint32_t x = srcOff; srcOff = dstOff; dstOff = x;
        }
        if (pass == 0) {
            /* S_99 sits in buffer B (99 is odd): T = B */
            if (TTM) {
                uint32_t tmc = tm;
                for (int k = 0; k < nw; ++k) {
                    if ((int32_t)wlp[k].x < 0) continue;           /* run header */
                    scb_tmem_st<KR>(tmc, scb_ld<KR>(bufB + (wlp[k].w & SCB_DST_MASK) + laneOff));
                    tmc += (uint32_t)KR;
                }
                scb_tmem_wait_st();
            } else
            for (int row = warp; row < N; row += S)
                scb_st<KR>(bufT + (uint32_t)row * ROWB + laneOff, scb_ld<KR>(bufB + (uint32_t)row * ROWB + laneOff));
            __syncthreads();
        }
    }
    /* S_98 in buffer A (98 is even).  e rows -> buffer B */
    for (int row = warp; row < N; row += S) {
        const ScbVec<KR> s98 = scb_ld<KR>(bufA + (uint32_t)row * ROWB + laneOff);
        const ScbVec<KR> x0 = scb_ldg<KR>(A.X + (long long)row * A.Wpad + w0 + lane * KR);
        ScbVec<KR> e;
#pragma unroll
        for (int k = 0; k < KR; ++k) e.v[k] = (s98.v[k] ^ x0.v[k]) & vm[lane * KR + k];
        scb_st<KR>(bufB + (uint32_t)row * ROWB + laneOff, e);
    }
    if (tid < GW) Fm[tid] &= vm[tid];
    __syncthreads();
}

/* shared-memory carve-up of the resolver kernels (after the three state buffers) */
SCB_HD inline size_t scb_resolve_smem_bytes(int N, int KR, bool ttm = false, int S = SCB_PROG_MAX_WARPS)
{
    /* state buffers (three, or two when T lives in tensor memory) + program + dmz/Fm/vm + scalars +
     * srcc[1024*KR] int16 + (shared-memory T only) D[2][<=24 warps][32*KR] + slack */
    return (size_t)(ttm ? 2 : 3) * N * 128 * KR + scb_img_entries(N, S) * 16 + (size_t)4 * 32 * KR * 4 + 64 + 384 + (size_t)1024 * KR * 2
           + (ttm ? 0 : (size_t)2 * 24 * 32 * KR * 4) + 256;
}

/* K3 phase 1: exact found mask and error sums of every queued group of KR chunks; leading not-found cells that
 * must copy a cell of an earlier chunk are queued for phase 2.  Only the chunks K2 flagged are summed here. */
template <int KR, bool TTM>
SCB_DEV void scb_resolve_item(const ScbSimArgs &A, unsigned char *smraw, unsigned it, uint32_t tm, bool useImg)
{
    const int N = A.N;
    const uint32_t ROWB = 128u * KR, BUFB = (uint32_t)N * ROWB;
    const int GW = 32 * KR, GC = 1024 * KR;
    unsigned char *sm = smraw;
    uint4 *desc = reinterpret_cast<uint4 *>(smraw + (TTM ? 2u : 3u) * BUFB);
    uint32_t *dmz = reinterpret_cast<uint32_t *>(desc + scb_img_entries(N, TTM ? (int)(blockDim.x >> 5) : SCB_PROG_MAX_WARPS));
    uint32_t *Fm = dmz + 2 * GW;                           /* dmz: two accumulators, used in turn by the compare steps */
    uint32_t *vm = Fm + GW;
    int *sh = reinterpret_cast<int *>(vm + GW);            /* [2] pending cells, [3] not-found cells */
    int *winfo = sh + 16;                                  /* [96] scratch of scb_layout_program */
    short *srcc = reinterpret_cast<short *>(winfo + 96);   /* [GC] source cell of each cell */
    uint32_t *Dbuf = reinterpret_cast<uint32_t *>(srcc + GC);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, S = nthr >> 5;
    const unsigned char *bufB = sm + BUFB;
    {
        const int p = (int)A.resolveItems[it].x;
        const int q = (int)(A.resolveItems[it].y & 0xFFFFFFu);
        const unsigned flagged = A.resolveItems[it].y >> 24;       /* bit j: chunk q + j was left out by K2 */
        const int4 cnt = scb_fetch_program<KR>(A, desc, winfo, p, 0, useImg, S, tid, nthr);
        if (tid == 0) { sh[2] = 0; sh[3] = 0; }
        __syncthreads();
        const bool haveT = it < A.resolveTCap;
        scb_resolve_chunk<KR, TTM>(A, sm, desc, winfo, useImg, cnt, p, dmz, Fm, vm, q, haveT ? 1 : 0,
                                   A.resolveT + (unsigned long long)it * N * GW, tid, nthr, Dbuf, tm);

        /* source cell of every cell: itself when found, else the nearest earlier found cell of the group
         * (the reference's fill-forward), -1 when that lies in an earlier chunk, -2 padding */
        for (int i = tid; i < GC; i += nthr) {
            const int w = i >> 5, b = i & 31;
            int s = -2;
            if ((vm[w] >> b) & 1u) {
                if ((Fm[w] >> b) & 1u) s = i;
                else {
                    s = -1;
                    const uint32_t m = Fm[w] & ((1u << b) - 1u);
                    if (m) s = w * 32 + (31 - __clz((int)m));
                    else for (int ww = w - 1; ww >= 0; --ww)
                        if (Fm[ww]) { s = ww * 32 + (31 - __clz((int)Fm[ww])); break; }
                    if ((flagged >> (i >> 10)) & 1u) {
                        atomicAdd(reinterpret_cast<unsigned int *>(&sh[3]), 1u);
                        if (s == -1) atomicAdd(reinterpret_cast<unsigned int *>(&sh[2]), 1u);
                    }
                }
            }
            srcc[i] = (short)s;
        }
        __syncthreads();

        if (A.mode == SCB_MODE_IMPORTANCE) {
            /* the simulator has counted every fixed-point cell; what is left to decide is whether the other cells
             * of the flagged chunks have an attractor window at all (the reference divides by zero otherwise) */
            if (tid == 0) {
                if (sh[3] > 0) {
                    atomicAdd(&A.fitness[p], (unsigned long long)sh[3]);
                    atomicAdd(&A.diag->not_found_cells, (unsigned long long)sh[3]);
                }
                atomicAdd(&A.diag->resolver_chunks, (unsigned long long)__popc(flagged));
            }
            __syncthreads();
            return;
        }
        /* per-node sums over the flagged chunks: found cells plus the copies made by not-found cells */
        unsigned tot = 0;
        for (int j = warp; j < N; j += S) {
            const int node = j;
            const uint32_t dsto = (uint32_t)node * ROWB;
            unsigned c = 0;
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                const int w = lane * KR + k;
                if (!((flagged >> (w >> 5)) & 1u)) continue;
                const uint32_t e = *reinterpret_cast<const uint32_t *>(bufB + dsto + (uint32_t)w * 4u);
                c += (unsigned)__popc(e & Fm[w]);
                uint32_t nfb = ~Fm[w] & vm[w];
                while (nfb) {
                    const int b = __ffs((int)nfb) - 1;
                    nfb &= nfb - 1u;
                    const int sc = srcc[w * 32 + b];
                    if (sc >= 0) c += (*reinterpret_cast<const uint32_t *>(bufB + dsto + (uint32_t)(sc >> 5) * 4u) >> (sc & 31)) & 1u;
                }
            }
            tot += c;
            if (A.mode == 1) {
                c = scb_warp_sum(c);
                if (lane == 0 && c) atomicAdd(&A.nodewise[(long long)p * N + node], (int)c);
            }
        }
        tot = scb_warp_sum(tot);
        if (lane == 0 && tot) atomicAdd(&A.fitness[p], (unsigned long long)tot);
        if (A.mode == 2) {
            for (int i = tid; i < GC; i += nthr) {
                const int sc = srcc[i];
                if (sc < 0 || !((flagged >> (i >> 10)) & 1u)) continue;
                int n1 = 0;
                for (int row = 0; row < N; ++row)
                    n1 += (int)((*reinterpret_cast<const uint32_t *>(bufB + (uint32_t)row * ROWB + (uint32_t)(sc >> 5) * 4u) >> (sc & 31)) & 1u);
                A.cellwise[(long long)p * A.C + (long long)q * 1024 + i] = n1;
            }
        }
        /* the last found cell of every flagged chunk and its error column (read by the fill-forward of later chunks) */
        if (warp < KR && ((flagged >> warp) & 1u)) {
            const int j = warp;
            const unsigned any = __ballot_sync(0xffffffffu, Fm[j * 32 + lane] != 0u);
            int lastFound = -1;
            if (any) {
                const int w = 31 - __clz((int)any);
                const uint32_t fw = Fm[j * 32 + w];
                lastFound = w * 32 + (31 - __clz((int)fw));
                const uint32_t wofs = (uint32_t)(j * 32 + w) * 4u;
                const int b = lastFound & 31;
                const int NB = (N + 31) >> 5;
                unsigned cntl = 0;
                for (int itn = 0; itn < NB; ++itn) {
                    const int row = itn * 32 + lane;
                    const unsigned bit = row < N ? (*reinterpret_cast<const uint32_t *>(bufB + (uint32_t)row * ROWB + wofs) >> b) & 1u : 0u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit != 0u);
                    cntl += (unsigned)__popc(bal);
                    if (A.mode == 1 && lane == 0) A.chunkLastBits[((long long)p * A.chunksPerInd + q + j) * NB + itn] = bal;
                }
                if (lane == 0) A.chunkLastErr[(long long)p * A.chunksPerInd + q + j] = (int)cntl;
            }
            if (lane == 0) A.chunkLast[(long long)p * A.chunksPerInd + q + j] = lastFound;
        }
        if (tid == 0) {
            if (sh[2] > 0) {
                const unsigned idx = atomicAdd(A.fillCount, 1u);
                A.fillItems[idx] = make_uint4((unsigned)p, (unsigned)q, (unsigned)sh[2], 0u);
            }
            if (sh[3] > 0) atomicAdd(&A.diag->not_found_cells, (unsigned long long)sh[3]);
            atomicAdd(&A.diag->resolver_chunks, (unsigned long long)__popc(flagged));
        }
        __syncthreads();
    }
}

#ifndef SCB_SIM_ONLY
/* K3 phase 1 as a kernel of its own (used when the simulator does not resolve in its own tail) */
template <int KR>
__global__ void __launch_bounds__(768, 1) scb_k_resolve(ScbResolveArgs R)
{
    SCB_DYN_SMEM(smraw);
    const ScbSimArgs &A = R.s;
    int *slot = reinterpret_cast<int *>(smraw + scb_resolve_smem_bytes(A.N, KR) - 64);
    unsigned nItems = *A.resolveCount;
    if (nItems > A.resolveCap) nItems = A.resolveCap;
    for (;;) {
        if (threadIdx.x == 0) slot[0] = (int)atomicAdd(R.itemCursor, 1u);
        __syncthreads();
        const unsigned it = (unsigned)slot[0];
        __syncthreads();
        if (it >= nItems) break;
        scb_resolve_item<KR>(A, smraw, it);
    }
}

/* K3 phase 2: chunks that start with not-found cells copy the last found cell of the nearest earlier chunk that
 * has one (chunkLast, complete after phase 1).  Only that ONE cell's error column is needed, and the simulator / the
 * resolver left it behind for every chunk (chunkLastErr, chunkLastBits): one warp per item, a lookup and a few
 * atomics.  No earlier found cell at all: the reference's value is uninitialised memory -> defined as 0 and counted. */
__global__ void scb_k_fill(ScbResolveArgs R)
{
    const ScbSimArgs &A = R.s;
    const int N = A.N, NB = (N + 31) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned nItems = *A.fillCount;
    const unsigned nWarps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < nItems; it += nWarps) {
        const int p = (int)A.fillItems[it].x, q = (int)A.fillItems[it].y, pend = (int)A.fillItems[it].z;
        int qq = q - 1, cell = -1;
        for (; qq >= 0; --qq) {
            cell = A.chunkLast[(long long)p * A.chunksPerInd + qq];
            if (cell >= 0) break;
        }
        if (qq < 0) {
            if (lane == 0) atomicAdd(&A.diag->undefined_leading, (unsigned long long)pend);
            if (A.mode == 2)
                for (int i = lane; i < pend; i += 32) A.cellwise[(long long)p * A.C + (long long)q * 1024 + i] = 0;
            continue;
        }
        const int err = A.chunkLastErr[(long long)p * A.chunksPerInd + qq];
        if (lane == 0 && err) atomicAdd(&A.fitness[p], (unsigned long long)err * (unsigned long long)pend);
        if (A.mode == 1)
            for (int i = lane; i < N; i += 32)
                if ((A.chunkLastBits[((long long)p * A.chunksPerInd + qq) * NB + (i >> 5)] >> (i & 31)) & 1u)
                    atomicAdd(&A.nodewise[(long long)p * N + i], pend);
        if (A.mode == 2)
            for (int i = lane; i < pend; i += 32) A.cellwise[(long long)p * A.C + (long long)q * 1024 + i] = err;
    }
}

/* ------------------------------------------------------------------------------------------
 * KG: the general simulator.  scSyncBool accepts what K2 cannot hold: any [7][3] term table -- a rule over up to 21
 * distinct state bits, where K0 compiles at most three into one LOP3 -- and networks up to the reference's compile-time
 * NODE (simulator.c:4-5), whose state does not fit shared memory.  Such input takes this path (and only such input,
 * or SCB_OPT_GENERIC): the reference's update rule evaluated term by term on packed words (32 cells per operation,
 * simulator.c:10-50, :377-416 as restated in scb_compile.h), the state of a 1024-cell chunk in global memory
 * (L2-resident), the attractor test literally as getAttractCycle2 (simulator.c:306-331): pass a = 99 steps -> T = S_99,
 * pass b = the same 98 steps again, every S_t compared with T.  One CTA per (individual, chunk); the epilogue is the
 * exact resolver's (fill-forward of not-found cells).  Exact, general and slow.
 * ---------------------------------------------------------------------------------------- */
SCB_DEV uint32_t scb_generic_node(const ScbGenericArgs &G, int p, int i, const uint32_t *cur, int lane, unsigned &flags)
{
    const int N = G.N;
    if (G.ko[i] == 1) return 0u;
    if (G.ki[i] == 1) return 0xffffffffu;
    const int al = G.andLen[i];
    int32_t tan[SCB_MAX_TERMS * SCB_MAX_IN], tai[SCB_MAX_TERMS * SCB_MAX_IN];
    const int32_t *an_i, *ai_i;
    if (G.compact) {
        const long long uoff = (G.tablesPerIndividual ? (long long)p * N * SCB_MAX_IN : 0) + (long long)i * SCB_MAX_IN;
        if (scb_expand_terms(G.an + uoff, G.ai + uoff, tan, tai) != al) flags |= SCB_NF_TERMS;
        an_i = tan; ai_i = tai;
    } else {
        const long long toff = (G.tablesPerIndividual ? (long long)p * N : 0) + i;
        an_i = G.an + toff * (SCB_MAX_TERMS * SCB_MAX_IN); ai_i = G.ai + toff * (SCB_MAX_TERMS * SCB_MAX_IN);
    }
    if (al == 1 || al == 0) {
        const int src = an_i[0];
        if (src < 0 || src >= N) { flags |= SCB_NF_BADIDX; return 0u; }
        const int fl = al == 1 ? ai_i[0] : 0;
        if (fl != 0 && fl != 1) return 0xffffffffu;
        const uint32_t x = cur[(long long)src * 32 + lane];
        return fl ? ~x : x;
    }
    const int32_t *ind = G.individuals + (long long)p * G.indLen;
    const int start = G.parse[i];
    const int end = i == N - 1 ? G.indLen : G.parse[i + 1];
    int nt = end - start;
    if (nt > SCB_MAX_TERMS) { flags |= SCB_NF_TERMS; nt = SCB_MAX_TERMS; }
    if (nt < 0 || start < 0 || end > G.indLen) { flags |= SCB_NF_TERMS; nt = 0; }
    uint32_t v = 0u;
    bool any = false;
    for (int a = 0; a < nt; ++a) {
        if (!ind[start + a]) continue;
        any = true;
        uint32_t t = 0xffffffffu;
        for (int b = 0; b < SCB_MAX_IN; ++b) {
            const int src = an_i[a * 3 + b];
            if (b > 0 && src <= -1) continue;
            if (src < 0 || src >= N) { flags |= SCB_NF_BADIDX; continue; }
            const int fl = ai_i[a * 3 + b];
            if (fl != 0 && fl != 1) continue;                          /* constant TRUE literal */
            const uint32_t x = cur[(long long)src * 32 + lane];
            t &= fl ? ~x : x;
        }
        v |= t;
    }
    if (!any) { flags |= SCB_NF_EMPTY; return 0u; }                    /* Q9 */
    return v;
}

#define SCB_GENERIC_SMEM (32 * 4 * 4 + 8 * 4 + 1024 * 2 + 1024 * 2)
__global__ void __launch_bounds__(256) scb_k_generic(ScbGenericArgs G)
{
    SCB_DYN_SMEM(smraw);                                 /* SCB_GENERIC_SMEM bytes */
    uint32_t *Fm = reinterpret_cast<uint32_t *>(smraw), *vm = Fm + 32;
    uint32_t (*acc)[32] = reinterpret_cast<uint32_t (*)[32]>(vm + 32);
    int *sh = reinterpret_cast<int *>(vm + 96);
    short *srcc = reinterpret_cast<short *>(sh + 8);
    short *jst = srcc + 1024;                            /* per cell: the largest t < 99 with S_t == S_99 = start of its attractor window */
    const int N = G.N;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, S = nthr >> 5;
    uint32_t *bA = G.scratch + (size_t)blockIdx.x * 3 * (size_t)N * 32, *bT = bA + (size_t)N * 32, *bC = bT + (size_t)N * 32;
    const unsigned total = (unsigned)G.P * (unsigned)G.chunksPerInd;
    for (;;) {
        __syncthreads();
        if (tid == 0) sh[0] = (int)atomicAdd(G.cursor, 1u);
        __syncthreads();
        const unsigned item = (unsigned)sh[0];
        if (item >= total) break;
        const int p = (int)(item / (unsigned)G.chunksPerInd), q = (int)(item % (unsigned)G.chunksPerInd);
        const int w0 = q * 32;
        const bool inX = w0 + lane < G.Wpad;
        unsigned flags = 0u;
        for (int pass = 0; pass < 2; ++pass) {
            for (int row = warp; row < N; row += S) bA[(size_t)row * 32 + lane] = inX ? G.X[(long long)row * G.Wpad + w0 + lane] : 0u;
            if (tid < 32) { vm[tid] = scb_valid_mask((long long)w0 + tid, G.C); acc[0][tid] = 0u; acc[1][tid] = 0u; if (pass == 0) Fm[tid] = 0u; }
            if (tid == 0) { sh[2] = 0; sh[3] = 0; }
            __syncthreads();
            if (q == 0 && pass == 0) {
                /* input diagnostics, once per individual (same counters as K0) */
                unsigned nEmpty = 0, nBad = 0, nTerms = 0;
                for (int i = tid; i < N; i += nthr) {
                    unsigned f = 0u;
                    (void)scb_generic_node(G, p, i, bA, 0, f);
                    nEmpty += (f & SCB_NF_EMPTY) ? 1u : 0u; nBad += (f & SCB_NF_BADIDX) ? 1u : 0u; nTerms += (f & SCB_NF_TERMS) ? 1u : 0u;
                }
                if (nEmpty) atomicAdd(&G.diag->empty_rule_nodes, (unsigned long long)nEmpty);
                if (nBad) atomicAdd(&G.diag->bad_index_nodes, (unsigned long long)nBad);
                if (nTerms) atomicAdd(&G.diag->term_overflow_nodes, (unsigned long long)nTerms);
            }
            uint32_t *cur = bA, *nxt = pass == 0 ? bT : bC;
            if (pass == 1) {
                uint32_t d = 0u;
                for (int row = warp; row < N; row += S) d |= cur[(size_t)row * 32 + lane] ^ bT[(size_t)row * 32 + lane];
                if (d) atomicOr(&acc[0][lane], d);
                __syncthreads();
                if (tid < 32) {
                    const uint32_t eq = ~acc[0][tid] & vm[tid];
                    Fm[tid] |= eq; acc[0][tid] = 0u;
                    for (uint32_t e = eq; e; e &= e - 1u) jst[tid * 32 + __ffs((int)e) - 1] = 0;
                }
                __syncthreads();
            }
            const int last = pass == 0 ? SCB_LAST : SCB_LAST - 1;
            for (int t = 1; t <= last; ++t) {
                uint32_t d = 0u;
                for (int row = warp; row < N; row += S) {
                    const uint32_t v = scb_generic_node(G, p, row, cur, lane, flags);
                    nxt[(size_t)row * 32 + lane] = v;
                    if (pass == 1) d |= v ^ bT[(size_t)row * 32 + lane];
                }
                if (pass == 1 && d) atomicOr(&acc[t & 1][lane], d);
                __syncthreads();
                /* the accumulator of step t is folded while the other warps run step t + 1 into the other one */
                if (pass == 1 && tid < 32) {
                    const uint32_t eq = ~acc[t & 1][tid] & vm[tid];
                    Fm[tid] |= eq; acc[t & 1][tid] = 0u;
                    for (uint32_t e = eq; e; e &= e - 1u) jst[tid * 32 + __ffs((int)e) - 1] = (short)t;     /* ascending t: the last one stays */
                }
                uint32_t *x = cur; cur = nxt; nxt = x;
            }
            /* pass 0 ends with S_99 in bT (99 is odd): the target.  Pass 1 ends with S_98 in bA (98 is even). */
        }
        /* e = S_98 xor S_0 on valid cells -> bC */
        for (int row = warp; row < N; row += S)
            bC[(size_t)row * 32 + lane] = (bA[(size_t)row * 32 + lane] ^ (inX ? G.X[(long long)row * G.Wpad + w0 + lane] : 0u)) & vm[lane];
        if (tid < 32) Fm[tid] &= vm[tid];
        __syncthreads();
        /* source cell of every cell: itself when found, else the nearest earlier found cell of the chunk (the
         * reference's fill-forward), -1 when that lies in an earlier chunk, -2 padding */
        for (int i = tid; i < 1024; i += nthr) {
            const int w = i >> 5, b = i & 31;
            int s = -2;
            if ((vm[w] >> b) & 1u) {
                if ((Fm[w] >> b) & 1u) s = i;
                else {
                    s = -1;
                    const uint32_t m = Fm[w] & ((1u << b) - 1u);
                    if (m) s = w * 32 + (31 - __clz((int)m));
                    else for (int ww = w - 1; ww >= 0; --ww)
                        if (Fm[ww]) { s = ww * 32 + (31 - __clz((int)Fm[ww])); break; }
                    atomicAdd(reinterpret_cast<unsigned int *>(&sh[3]), 1u);
                    if (s == -1) atomicAdd(reinterpret_cast<unsigned int *>(&sh[2]), 1u);
                }
            }
            srcc[i] = (short)s;
        }
        __syncthreads();
        unsigned tot = 0;
        for (int node = warp; node < N; node += S) {
            const uint32_t e = bC[(size_t)node * 32 + lane];
            unsigned c = (unsigned)__popc(e & Fm[lane]);
            uint32_t nfb = ~Fm[lane] & vm[lane];
            while (nfb) {
                const int b = __ffs((int)nfb) - 1;
                nfb &= nfb - 1u;
                const int sc = srcc[lane * 32 + b];
                if (sc >= 0) c += (bC[(size_t)node * 32 + (sc >> 5)] >> (sc & 31)) & 1u;
            }
            tot += c;
            if (G.mode == 1) {
                c = scb_warp_sum(c);
                if (lane == 0 && c) atomicAdd(&G.nodewise[(long long)p * N + node], (int)c);
            }
        }
        tot = scb_warp_sum(tot);
        if (lane == 0 && tot) atomicAdd(&G.fitness[p], (unsigned long long)tot);
        if (G.mode == 2) {
            for (int i = tid; i < 1024; i += nthr) {
                const int sc = srcc[i];
                if (sc < 0) continue;
                int n1 = 0;
                for (int row = 0; row < N; ++row) n1 += (int)((bC[(size_t)row * 32 + (sc >> 5)] >> (sc & 31)) & 1u);
                G.cellwise[(long long)p * G.C + (long long)q * 1024 + i] = n1;
            }
        }
        if (warp == 0) {
            /* the chunk's last found cell and its error column, for the fill-forward of later chunks */
            const unsigned any = __ballot_sync(0xffffffffu, Fm[lane] != 0u);
            int lastFound = -1;
            if (any) {
                const int w = 31 - __clz((int)any);
                lastFound = w * 32 + (31 - __clz((int)Fm[w]));
                const int b = lastFound & 31, NB = (N + 31) >> 5;
                unsigned cntl = 0;
                for (int itn = 0; itn < NB; ++itn) {
                    const int row = itn * 32 + lane;
                    const unsigned bit = row < N ? (bC[(size_t)row * 32 + w] >> b) & 1u : 0u;
                    const unsigned bal = __ballot_sync(0xffffffffu, bit != 0u);
                    cntl += (unsigned)__popc(bal);
                    if (G.mode == 1 && lane == 0) G.chunkLastBits[((long long)p * G.chunksPerInd + q) * NB + itn] = bal;
                }
                if (lane == 0) G.chunkLastErr[(long long)p * G.chunksPerInd + q] = (int)cntl;
            }
            if (lane == 0) {
                G.chunkLast[(long long)p * G.chunksPerInd + q] = lastFound;
                if (sh[2] > 0) {
                    const unsigned idx = atomicAdd(G.fillCount, 1u);
                    G.fillItems[idx] = make_uint4((unsigned)p, (unsigned)q, (unsigned)sh[2], 0u);
                }
                if (sh[3] > 0) atomicAdd(&G.diag->not_found_cells, (unsigned long long)sh[3]);
                atomicAdd(&G.diag->resolver_chunks, 1ull);
                atomicAdd(&G.diag->tiles, 1ull);
                atomicAdd(&G.diag->steps_executed, (unsigned long long)(2 * SCB_LAST - 1));
            }
        }
        if (G.mode == 1 && N >= 1000) {
            /* SURVEY Q8 (simulator.c:457-468): walking a found cell's window [j, 99) in order, the first state whose
             * error count is a new minimum AND <= 0.001 * nodeNum adds its error vector to the per-node sums one extra
             * time (and ends the search: minError = 0).  Below 1000 nodes the threshold is < 1, the vector is all zero
             * and nothing happens; from 1000 nodes on it changes the sums, so a third pass replays the trajectory and
             * does exactly that, one thread per cell. */
            __syncthreads();
            for (int row = warp; row < N; row += S) bA[(size_t)row * 32 + lane] = inX ? G.X[(long long)row * G.Wpad + w0 + lane] : 0u;
            __syncthreads();
            int mn[4];
            bool live[4];
            for (int k = 0; k < 4; ++k) {
                const int cell = tid + k * nthr;
                mn[k] = 100000;
                live[k] = cell < 1024 && ((Fm[cell >> 5] >> (cell & 31)) & 1u);
            }
            uint32_t *cur = bA, *nxt = bT;
            for (int t = 0; t <= SCB_LAST - 1; ++t) {
                if (t > 0) {
                    for (int row = warp; row < N; row += S) nxt[(size_t)row * 32 + lane] = scb_generic_node(G, p, row, cur, lane, flags);
                    __syncthreads();
                    uint32_t *x = cur; cur = nxt; nxt = x;
                }
                for (int k = 0; k < 4; ++k) {
                    const int cell = tid + k * nthr;
                    if (!live[k] || t < (int)jst[cell]) continue;
                    const int w = cell >> 5, b = cell & 31;
                    int tot = 0;
                    for (int row = 0; row < N; ++row)
                        tot += (int)(((cur[(size_t)row * 32 + w] ^ G.X[(long long)row * G.Wpad + w0 + w]) >> b) & 1u);
                    if (tot < mn[k]) {
                        mn[k] = tot;
                        if ((double)tot <= 0.001 * (double)N) {
                            mn[k] = 0; live[k] = false;
                            for (int row = 0; row < N && tot > 0; ++row)
                                if (((cur[(size_t)row * 32 + w] ^ G.X[(long long)row * G.Wpad + w0 + w]) >> b) & 1u) atomicAdd(&G.nodewise[(long long)p * N + row], 1);
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Fitness exchange over peer memory: the all-gather of a GA generation as the last kernel of the
 * generation's graph.  One CTA: the shard (a few hundred 8-byte values) is stored straight into every rank's gather
 * buffer over NVLink, a system-scope fence orders the stores, then this rank's flag is raised in every rank's flag
 * array; the CTA leaves when all ranks have raised theirs here.  No host round trip, no NCCL launch: the payload is
 * 4 KB, the exchange is pure latency.  Buffers are double buffered by generation parity: a rank can run at most one
 * generation ahead of the slowest one, whose host has consumed generation g before it launches g + 1.
 * ---------------------------------------------------------------------------------------- */
#ifndef SCB_EMU
__global__ void scb_k_gather(ScbGatherArgs G)
{
    __shared__ unsigned int gen;
    if (threadIdx.x == 0) gen = *G.generation + 1u;
    __syncthreads();
    const unsigned int g = gen;
    const size_t half = (size_t)(g & 1u) * G.world * G.width;
    for (int r = 0; r < G.world; ++r) {
        unsigned long long *dst = G.peerBuf[r] + half + (size_t)G.rank * G.width;
        for (int i = threadIdx.x; i < G.width; i += blockDim.x) dst[i] = i < G.count ? G.fitness[i] : 0ull;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < G.world) {
#if defined(__CUDA_ARCH__)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(G.peerFlag[threadIdx.x] + G.rank), "r"(g) : "memory");
        const unsigned int *mine = G.peerFlag[G.rank] + threadIdx.x;
        unsigned int v = 0u;
        long long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - g) >= 0) break;
            if (++spins > 400000000ll) { atomicExch(G.timeout, 1u); break; }       /* seconds: a peer is gone */
            __nanosleep(64);
        }
#else
        G.peerFlag[threadIdx.x][G.rank] = g;
#endif
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *G.generation = g;
}
#endif

/* ------------------------------------------------------------------------------------------
 * K4: trajectories for attractor extraction (the `cluster` companion, simulator.c:487-552, and
 * singleCell.__cluster, singleCell.py:990-1059).
 *
 * scb_k_traj   one CTA per 1024-cell tile (K = 1): runs `steps` states of one compiled program and
 *              stores every state tile-locally, traj[tile][step][N][32] words (the snapshot-store
 *              path of scb_step writes them).
 * scb_k_window per tile: row-pair equality masks, the reference's scan order for the attractor
 *              window, and the window's states gathered into cell-major packed words.
 * ---------------------------------------------------------------------------------------- */
__global__ void scb_k_traj(ScbTrajArgs A)
{
    SCB_DYN_SMEM(smraw);
    const int N = A.N;
    const uint32_t ROWB = 128u, BUFB = (uint32_t)N * ROWB;
    unsigned char *sm = smraw;
    uint4 *desc = reinterpret_cast<uint4 *>(smraw + 2u * BUFB);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, S = nthr >> 5;
    const uint32_t laneOff = (uint32_t)lane * 4u;
    const int4 cnt = A.counts[0];
    const scb_saddr sa = scb_smem_addr(sm), sdesc = scb_smem_addr(desc);
    int *winfo = reinterpret_cast<int *>(desc + scb_img_entries(N, S));
    scb_layout_program(desc, winfo, A.descs, cnt, N, S, 128u, tid, nthr);
    const int e1 = cnt.x + cnt.y + cnt.z;
    int woff, nw;
    scb_img_warp(desc, warp, S, woff, nw);
    const scb_saddr wl = sdesc + 16u * (uint32_t)woff;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
        const int w0 = tile * 32;
        uint32_t *tbase = A.traj + (unsigned long long)tile * A.steps * N * 32;
        __syncthreads();
        for (int row = warp; row < N; row += S) {
            const uint32_t v = __ldg(A.X + (long long)row * A.Wpad + w0 + lane);
            *reinterpret_cast<uint32_t *>(sm + (uint32_t)row * ROWB + laneOff) = v;
            tbase[(long long)row * 32 + lane] = v;                         /* state 0 */
        }
        __syncthreads();
        uint32_t srcOff = 0u, dstOff = BUFB;
        ScbVec<1> d2, dz;
        d2.v[0] = 0u; dz.v[0] = 0u;
        for (int t = 1; t < A.steps; ++t) {
            unsigned char *zg = reinterpret_cast<unsigned char *>(tbase + (unsigned long long)t * N * 32) + laneOff;
            scb_step<1, false>(sa + srcOff + laneOff, sa + dstOff + laneOff, wl, nw, SCB_SF_ZSNAP, zg, sa, d2, dz);
            scb_const_rows<1>(sa + dstOff + laneOff, A.descs + e1, cnt.w, warp, S, zg);
            __syncthreads();
            const uint32_t x = srcOff; srcOff = dstOff; dstOff = x;
        }
        if (A.recOn) {
            __syncthreads();
            for (int node = warp; node < N; node += S) {
                if (!A.recOn[node]) continue;
                for (int t = 1; t < A.steps; ++t)
                    tbase[((unsigned long long)t * N + node) * 32 + lane] = ((A.recBits[node * 4 + (t >> 5)] >> (t & 31)) & 1u) ? 0xffffffffu : 0u;
            }
        }
    }
}

#define SCB_WIN_MAXPAIRS 128
__global__ void scb_k_window(ScbWindowArgs A)
{
    SCB_DYN_SMEM(smraw);
    uint32_t *eq = reinterpret_cast<uint32_t *>(smraw);                  /* [pairs][32] equality masks */
    const int N = A.N, T = A.steps;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, S = nthr >> 5;
    /* pairs in the reference's scan order: i = T-1 .. 0, j = i-1 .. 0 (getAttractCycle2).  For the C
     * semantics only i = T-1 can match first (a deterministic trajectory whose last state repeats no
     * earlier one has no repeated pair at all), so only those T-1 pairs are evaluated. */
    const int nPairs = A.sem == 0 ? T - 1 : T * (T - 1) / 2;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
        const uint32_t *tbase = A.traj + (unsigned long long)tile * T * N * 32;
        __syncthreads();
        for (int pr = warp; pr < nPairs; pr += S) {
            int i, j;
            if (A.sem == 0) { i = T - 1; j = T - 2 - pr; }
            else {
                int r = pr; i = T - 1;
                while (r >= i) { r -= i; --i; }
                j = i - 1 - r;
            }
            uint32_t d = 0u;
            const uint32_t *ri = tbase + (unsigned long long)i * N * 32 + lane, *rj = tbase + (unsigned long long)j * N * 32 + lane;
            for (int n = 0; n < N; ++n) d |= ri[n * 32] ^ rj[n * 32];
            eq[pr * 32 + lane] = ~d;
        }
        __syncthreads();
        for (int cl = tid; cl < 1024; cl += nthr) {
            const long long c = (long long)tile * 1024 + cl;
            if (c >= A.C) continue;
            const int w = cl >> 5, b = cl & 31;
            int r0, r1, first, count;
            int hit = -1;
            for (int pr = 0; pr < nPairs; ++pr) if ((eq[pr * 32 + w] >> b) & 1u) { hit = pr; break; }
            if (A.sem == 0) {
                if (hit < 0) { r0 = 0; r1 = 0; first = 0; count = 0; }               /* simulator.c:307 res = {0,0} */
                else { r1 = T - 1; r0 = T - 2 - hit; first = r0; count = r1 - r0; }   /* states [r0, r1) */
            } else {
                if (hit < 0) { r0 = -1; r1 = -1; first = T - 1; count = 1; }          /* vals[-1]: the last row */
                else {
                    int r = hit, i = T - 1;
                    while (r >= i) { r -= i; --i; }
                    r1 = i; r0 = i - 1 - r; first = r0; count = r1 - r0 + 1;          /* rows r0..r1 inclusive */
                }
            }
            A.res[c * 2] = r0; A.res[c * 2 + 1] = r1; A.nStates[c] = count;
            const int ns = count < A.maxStates ? count : A.maxStates;
            for (int s = 0; s < ns; ++s) {
                const uint32_t *row = tbase + (unsigned long long)(first + s) * N * 32 + w;
                uint32_t *out = A.states + ((unsigned long long)c * A.maxStates + s) * A.NW;
                for (int wq = 0; wq < A.NW; ++wq) {
                    uint32_t acc = 0u;
                    const int lim = (N - wq * 32) < 32 ? (N - wq * 32) : 32;
                    for (int q = 0; q < lim; ++q) acc |= ((row[(wq * 32 + q) * 32] >> b) & 1u) << q;
                    out[wq] = acc;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * K5: Hamming distance of every cell's observed state to every attractor + first-minimum
 * argmin (singleCell.py:1211-1238).  One thread per cell.  The cell's state is gathered from
 * the node-major packed matrix into cell-major words (bit i%32 of word i/32 = node i; a warp's
 * 32 cells share each source word, so the gather is a broadcast load), attractors are staged
 * through shared memory as packed node words [A][ceil(N/32)].  POPC-bound.
 * ---------------------------------------------------------------------------------------- */
#define SCB_HAM_MAXW 32            /* N <= 1024 */
#define SCB_HAM_SMEM_WORDS 8192    /* 32 KB of attractor words per pass */
/* ---- unique attractor states (singleCell.py:1189-1207 collects them in a Python set) ------------------------------
 * A hash table of row indices: every valid row finds the slot of its state (linear probing; a slot is claimed once and
 * then only lowered to the smallest row index with that state), a second pass flags the rows that are their state's
 * first occurrence, a one-CTA pass writes those states out in row order (= first-seen order). */
SCB_DEV unsigned scb_row_hash(const uint32_t *k, int NW)
{
    unsigned h = 2166136261u;
    for (int w = 0; w < NW; ++w) { h ^= k[w]; h *= 16777619u; h ^= h >> 15; }
    return h;
}
SCB_DEV bool scb_row_valid(const ScbDedupeArgs &A, long long r)
{
    const int c = (int)(r / A.maxStates), s = (int)(r % A.maxStates);
    const int n = A.nStates[c] < A.maxStates ? A.nStates[c] : A.maxStates;
    return s < n;
}
SCB_DEV bool scb_row_equal(const uint32_t *a, const uint32_t *b, int NW)
{
    for (int w = 0; w < NW; ++w) if (a[w] != b[w]) return false;
    return true;
}
__global__ void scb_k_dedupe_insert(ScbDedupeArgs A)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (long long)A.C * A.maxStates || !scb_row_valid(A, r)) return;
    const uint32_t *key = A.states + r * A.NW;
    unsigned slot = scb_row_hash(key, A.NW) & A.mask;
    for (;;) {
        unsigned cur = scb_ld_volatile(&A.table[slot]);
        if (cur == 0xffffffffu) {
            const unsigned prev = atomicCAS(&A.table[slot], 0xffffffffu, (unsigned)r);
            if (prev == 0xffffffffu) return;
            cur = prev;
        }
        if (scb_row_equal(A.states + (long long)cur * A.NW, key, A.NW)) { atomicMin(&A.table[slot], (unsigned)r); return; }
        slot = (slot + 1u) & A.mask;
    }
}
__global__ void scb_k_dedupe_flag(ScbDedupeArgs A)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (long long)A.C * A.maxStates || !scb_row_valid(A, r)) return;
    const uint32_t *key = A.states + r * A.NW;
    unsigned slot = scb_row_hash(key, A.NW) & A.mask;
    for (;;) {
        const unsigned cur = A.table[slot];
        if (scb_row_equal(A.states + (long long)cur * A.NW, key, A.NW)) { A.flags[r] = cur == (unsigned)r ? 1 : 0; return; }
        slot = (slot + 1u) & A.mask;
    }
}
__global__ void scb_k_dedupe_compact(ScbDedupeArgs A)
{
    SCB_DYN_SMEM(smraw);                                  /* blockDim.x + 1 unsigned */
    unsigned *pre = reinterpret_cast<unsigned *>(smraw);
    const long long rows = (long long)A.C * A.maxStates;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long seg = (rows + nthr - 1) / nthr, lo = seg * tid, hi = lo + seg < rows ? lo + seg : rows;
    unsigned n = 0;
    for (long long r = lo; r < hi; ++r) n += A.flags[r];
    pre[tid + 1] = n;
    __syncthreads();
    if (tid == 0) {
        pre[0] = 0u;
        for (int t = 0; t < nthr; ++t) pre[t + 1] += pre[t];
        *A.count = pre[nthr];
    }
    __syncthreads();
    unsigned o = pre[tid];
    for (long long r = lo; r < hi; ++r) {
        if (!A.flags[r]) continue;
        if (o < (unsigned)A.cap) for (int w = 0; w < A.NW; ++w) A.out[(long long)o * A.NW + w] = A.states[r * A.NW + w];
        ++o;
    }
}

/* attractors as the caller holds them (int32 0/1 per node) -> bit rows of NW4 * 4 words, padded with zeros: one warp
 * per output word, a ballot per 32 values; anything but 0/1 is counted in *bad */
__global__ void scb_k_pack_attr(const int32_t *raw, uint32_t *packed, int nAttr, int N, int NWP, unsigned int *bad)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nWarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long o = warp; o < (long long)nAttr * NWP; o += nWarps) {
        const long long a = o / NWP;
        const int w = (int)(o - a * NWP), i = w * 32 + lane;
        const int32_t v = i < N ? raw[a * N + i] : 0;
        const unsigned word = __ballot_sync(0xffffffffu, v == 1);
        if (v != 0 && v != 1) atomicAdd(bad, 1u);
        if (lane == 0) packed[o] = word;
    }
}

/* packed attractor rows of NW words -> rows of NWP >= NW words, zero padded */
__global__ void scb_k_restride(const uint32_t *src, uint32_t *dst, long long rows, int NW, int NWP)
{
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < rows * NWP; o += (long long)gridDim.x * blockDim.x) {
        const long long r = o / NWP;
        const int w = (int)(o - r * NWP);
        dst[o] = w < NW ? src[r * NW + w] : 0u;
    }
}

/* NW4 = 16-byte groups per state (N <= 128 * NW4): the cell's state lives in registers, the attractors of a pass are
 * read from shared memory as broadcast 16-byte loads; per 32 node bits one XOR and one POPC (the bound: POPC issues
 * at a quarter of the LOP3 rate) */
template <int NW4>
__global__ void __launch_bounds__(128) scb_k_hamming(ScbHammingArgs A)
{
    SCB_DYN_SMEM(smraw);
    constexpr int NWP = NW4 * 4;
    const uint4 *sAttr = reinterpret_cast<const uint4 *>(smraw);
    uint4 *sAttrW = reinterpret_cast<uint4 *>(smraw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long c = (long long)blockIdx.x * nthr + tid;
    uint32_t cs[NWP];
#pragma unroll
    for (int w = 0; w < NWP; ++w) {
        uint32_t acc = 0u;
        if (c < A.C && w * 32 < A.N) {
            const int lim = (A.N - w * 32) < 32 ? (A.N - w * 32) : 32;
            for (int b = 0; b < lim; ++b)
                acc |= ((A.X[(long long)(w * 32 + b) * A.Wpad + (c >> 5)] >> (c & 31)) & 1u) << b;
        }
        cs[w] = acc;
    }
    int best = -1, bestD = 0;
    const int ACH = (SCB_HAM_SMEM_WORDS / 4) / NW4;           /* attractors per shared-memory pass */
    const uint4 *gAttr = reinterpret_cast<const uint4 *>(A.attr);
    for (int a0 = 0; a0 < A.A; a0 += ACH) {
        const int na = (A.A - a0) < ACH ? (A.A - a0) : ACH;
        __syncthreads();
        for (int i = tid; i < na * NW4; i += nthr) sAttrW[i] = gAttr[(long long)a0 * NW4 + i];
        __syncthreads();
        if (c < A.C) {
            for (int a = 0; a < na; ++a) {
                /* Hamming weight of NWP words with carry-save adders (Harley-Seal): three words -> a sum word and a
                 * carry word with two LOP3 (0x96 = a^b^c, 0xE8 = majority), so that 4 words cost 2 POPC + 2 weighted adds
                 * less the carries deferred to the next group, instead of 4 POPC: POPC issues at a quarter of the LOP3 rate */
                int d = 0;
                uint32_t ones = 0u, twos = 0u;
#pragma unroll
                for (int g = 0; g < NW4; ++g) {
                    const uint4 t = sAttr[a * NW4 + g];
                    const uint32_t x0 = cs[4 * g] ^ t.x, x1 = cs[4 * g + 1] ^ t.y, x2 = cs[4 * g + 2] ^ t.z, x3 = cs[4 * g + 3] ^ t.w;
                    /* ones + x0 + x1 -> (ones', c0);  ones' + x2 + x3 -> (ones'', c1);  twos + c0 + c1 -> (twos', fours) */
                    const uint32_t s0 = ones ^ x0 ^ x1, c0 = (ones & x0) | (ones & x1) | (x0 & x1);
                    const uint32_t s1 = s0 ^ x2 ^ x3, c1 = (s0 & x2) | (s0 & x3) | (x2 & x3);
                    const uint32_t t2 = twos ^ c0 ^ c1, fours = (twos & c0) | (twos & c1) | (c0 & c1);
                    ones = s1; twos = t2;
                    d += 4 * __popc(fours);
                }
                d += __popc(ones) + 2 * __popc(twos);
                if (A.dist) A.dist[c * A.A + a0 + a] = d;
                if (best < 0 || d < bestD) { best = a0 + a; bestD = d; }     /* first minimum */
            }
        }
    }
    if (c < A.C) {
        A.decider[c] = best;
        A.deciderDistance[c] = bestD;
    }
}

#endif  /* SCB_SIM_ONLY */

#endif
