/*
 * scb_sim_tu.cu -- one instantiation of the simulator kernel K2, scb_k_sim<SCB_SIM_K, SCB_SIM_TM> (see scb_sim_tu.h).
 * Built six times (K = 1, 2, 4 words per lane x snapshot in tensor memory or in global scratch).
 */
#ifndef SCB_SIM_ONLY
#define SCB_SIM_ONLY
#endif
#include "scb_sim_tu.h"
#include "scb_kernels.cuh"
#undef SCB_CAT2
#undef SCB_CAT
#undef SCB_FN

#ifndef SCB_SIM_K
#error "compile with -DSCB_SIM_K=1|2|4 -DSCB_SIM_TM=0|1"
#endif
#define SCB_CAT2(a, b, c, d) a##b##c##d
#define SCB_CAT(a, b, c, d) SCB_CAT2(a, b, c, d)
#define SCB_FN(stem) SCB_CAT(stem, SCB_SIM_K, _tm, SCB_SIM_TM)

int SCB_FN(scb_sim_launch_k)(dim3 grid, dim3 block, size_t smem, cudaStream_t st, const ScbSimArgs &a)
{
    SCB_LAUNCH((scb_k_sim<SCB_SIM_K, (SCB_SIM_TM != 0)>), grid, block, smem, st, a);
    return (int)cudaGetLastError();
}

int SCB_FN(scb_sim_occupancy_k)(int threads, size_t smem, int *occ)
{
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, (scb_k_sim<SCB_SIM_K, (SCB_SIM_TM != 0)>), threads, smem);
}

int SCB_FN(scb_sim_set_smem_k)(int bytes)
{
    return (int)cudaFuncSetAttribute((scb_k_sim<SCB_SIM_K, (SCB_SIM_TM != 0)>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
