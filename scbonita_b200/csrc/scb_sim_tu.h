/*
 * scb_sim_tu.h -- the simulator kernel K2 (scb_k_sim<K, TM>) is compiled in translation units of its own, one per
 * instantiation (scb_sim_tu.cu with -DSCB_SIM_K=.. -DSCB_SIM_TM=..): its generated 256-way run interpreter makes it
 * by far the longest compile of the library, and the six instantiations build in parallel.  The host API reaches
 * them through these three functions.
 */
#ifndef SCB_SIM_TU_H
#define SCB_SIM_TU_H
#include "scb_types.h"

#ifdef SCB_EMU
#define SCB_LAUNCH(kernel, grid, block, smem, stream, ...) emu_launch(kernel, grid, block, smem, __VA_ARGS__)
#else
#define SCB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

/* each returns a cudaError_t */
#define SCB_SIM_DECL(K, TM)                                                                                      \
    int scb_sim_launch_k##K##_tm##TM(dim3 grid, dim3 block, size_t smem, cudaStream_t st, const ScbSimArgs &a); \
    int scb_sim_occupancy_k##K##_tm##TM(int threads, size_t smem, int *occ);                                     \
    int scb_sim_set_smem_k##K##_tm##TM(int bytes);
SCB_SIM_DECL(1, 0) SCB_SIM_DECL(1, 1) SCB_SIM_DECL(2, 0) SCB_SIM_DECL(2, 1) SCB_SIM_DECL(4, 0) SCB_SIM_DECL(4, 1)
#undef SCB_SIM_DECL

#endif
