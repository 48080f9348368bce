/*
 * scb_api.cu -- host side of the C-ABI declared in include/scbonita_b200.h.
 *
 * Owns device memory, the stream and the launch geometry; all arithmetic of the hot path
 * happens in the kernels (scb_kernels.cuh).  There is deliberately no CPU implementation
 * behind any entry point: without a usable device every call returns SCB_ERR_CUDA.
 *
 * Reference interface replaced: the ctypes boundary of ruleMaker.__evaluateByNode
 * (ruleMaker.py:1077-1221) onto simulator.c's scSyncBool / importanceScore / cluster.
 */
#include "../../include/scbonita_b200.h"
#include "scb_sim_tu.h"
#include "scb_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>


namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(SCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

/* bumped by every (re)allocation of a device buffer: a captured generation graph holds raw pointers */
std::atomic<unsigned long long> g_allocEpoch{0};

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;            /* elements */
    int ensure(size_t n, bool inGraphs = true)
    {
        if (n <= cap && p) return SCB_OK;
        if (inGraphs) ++g_allocEpoch;                /* (a buffer no captured graph refers to leaves the graphs valid) */
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = std::max<size_t>(n, 1);
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T));
        if (e != cudaSuccess) return fail(SCB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
        cap = want;
        return SCB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct SimGeom { int K = 0, warps = 0, ctas = 0, tmemCols = 0, tmemSlotsPerWarp = 0; size_t smem = 0; };

struct scb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t evRun0 = nullptr, evRun1 = nullptr, evSim0 = nullptr, evSim1 = nullptr, evK0 = nullptr, evFill0 = nullptr, evFill1 = nullptr;
    int numSM = 0;
    size_t smemLimit = 0;
    int N = 0, C = 0, W = 0, Wpad = 0;
    /* options */
    int optEarly = 1, optSnap = 1, optWarps = 0, optMaxCtas = 0, optWords = 0, optChk = 12, optSearch = 0, optMinGap = 8;
    int optTmem = 1, optFuse = 1, optTma = 1, optSteady = 1, optChkFrom = SCB_NPROG - 2, optGraph = 1;
    unsigned long long optVersion = 0;   /* bumped by every scb_ctx_set_option */
    unsigned long long snapLo = (1ull << 36), snapHi = (1ull << 2);   /* snapshots when computing S_36 and S_66 */
    /* geometry (derived) */
    int K = 0, warps = 0, ctas = 0, occ = 1;
    SimGeom main, tail;
    int mainTiles = 0, tailTiles = 0;
    size_t smemSim = 0, smemRes = 0;
    int resWarps = 8, resCtas = 0, resGroup = 1, fillWarps = 8, fillCtas = 0;
    bool geomValid = false;
    /* device data */
    DevBuf<uint32_t> dX;
    DevBuf<int32_t> dInd, dAndLen, dParse, dAn, dAi, dKo, dKi;
    DevBuf<uint4> dDescs, dProgs, dImgs;
    DevBuf<int4> dCounts;
    DevBuf<unsigned char> dProgFrom;
    DevBuf<unsigned long long> dFitness;
    DevBuf<int32_t> dNodewise, dCellwise;
    DevBuf<unsigned int> dCounters;      /* [0] K2 work head, [1] resolve count, [2] K3 work head */
    DevBuf<uint2> dItems;
    DevBuf<uint4> dFillItems;
    DevBuf<int32_t> dChunkLast, dChunkLastErr, dAttrRaw;
    DevBuf<uint32_t> dChunkLastBits;
    DevBuf<uint32_t> dResolveT;
    DevBuf<unsigned int> dReady;
    unsigned resolveTCap = 0;
    DevBuf<uint32_t> dZ;
    DevBuf<ScbDiag> dDiag;
    DevBuf<uint32_t> dAttr;
    DevBuf<int32_t> dDist, dDecider, dDeciderDist;
    DevBuf<unsigned char> dFlush;
    DevBuf<unsigned int> dCost;          /* [2][P] cost per individual in the last run: sum over its tiles, costliest tile */
    DevBuf<int32_t> dOrder;              /* [P] individuals, costliest first (scb_pop_set_order) */
    bool haveOrder = false;
    /* resident population (scb_pop_update): the next run scores only the replaced slots */
    /* the general simulator (scb_k_generic): the only path when the state does not fit shared memory, the second pass of
     * a population that holds rules over more than three distinct inputs, or on request (SCB_OPT_GENERIC) */
    bool genericOnly = false, optGeneric = false, lastGeneric = false;
    int optResGroup = 0;
    bool optAutoOrder = true;
    DevBuf<unsigned int> dPred;
    DevBuf<int32_t> dOrderAuto;
    int lastSem = 0;
    DevBuf<uint32_t> dGScratch;
    bool delta = false;
    int nEval = 0;
    DevBuf<int32_t> dSlots, dStage;
    /* the generation as a CUDA graph: clears + K0 + K2 + fill (+ peer gather), replayed while nothing it refers to moves */
#ifndef SCB_EMU
    cudaGraphExec_t graphExec = nullptr;
#endif
    unsigned long long graphKey = 0;
    /* fitness exchange over peer memory (scb_gather_*) */
    int gWorld = 0, gRank = 0, gWidth = 0;
    bool gConnected = false;
    unsigned long long *gBuf = nullptr;  /* own gather buffer [2][world * width] */
    unsigned int *gFlags = nullptr;      /* own flag array [world], then generation counter, then timeout flag */
    unsigned long long *gPeerBuf[SCB_MAX_PEERS] = {};
    unsigned int *gPeerFlag[SCB_MAX_PEERS] = {};
    DevBuf<uint32_t> dTraj, dStates, dRecBits, dDedupeTable, dDedupeOut;
    DevBuf<unsigned char> dDedupeFlags;
    DevBuf<unsigned char> dRecOn;
    DevBuf<int32_t> dRes, dNStates;
    /* current population */
    int P = 0, indLen = 0, tpi = 0;
    int koStride = 0;                    /* 0: shared knock-out array, N: one per program */
    int compact = 0;                     /* dAn / dAi hold [P or 1][N][3] upstream triples instead of term tables */
    int lastMode = -1;
    bool ran = false;
    ScbDiag hDiag = {};
    std::recursive_mutex mu;             /* one lock per public call; composite calls hold it across upload + run + fetch */
};

namespace {

/* launch geometry of the simulator kernel for tiles of K words per lane and S warps per CTA: shared memory of the
 * launch (the in-kernel resolver shares the CTA), CTAs per SM, tensor-memory columns */
int sim_geometry(scb_ctx *c, int K, int S, SimGeom *g, int *occOut)
{
    const int N = c->N;
    g->K = K;
    S = std::max(S, K);                 /* the mask arrays need 32*K threads */
    S = std::min(S, SCB_SIM_MAX_WARPS); /* __launch_bounds__ of scb_k_sim */
    g->warps = S;
    g->smem = scb_smem_bytes(N, K, 2, S);
    /* a warp's piece of the table-sorted program is cut by rows moved, not by nodes: at most 4N/S + 4 rows, i.e.
     * 2 ceil(N/S) + 2 one-input nodes (scb_layout_program): slots of the snapshot / the resolver's target state */
    const int npw = 2 * ((N + S - 1) / S) + 2;
    const int need = K * ((S + 3) / 4) * npw;
    int cols = 32;
    while (cols < need) cols <<= 1;
    const bool tmemFits = c->optTmem && c->optSnap && cols <= 512;
    /* inside the simulator launch the resolver keeps its target state in tensor memory and works on whole tiles with
     * the simulator's own two-buffer layout; without tensor memory it needs three buffers of one or two chunks */
    size_t smemLaunch = g->smem;
    if (c->optFuse) {
        if (tmemFits) smemLaunch = std::max(smemLaunch, scb_resolve_smem_bytes(N, K, true, S));
        else smemLaunch = std::max(smemLaunch, scb_resolve_smem_bytes(N, (K >= 2 && scb_resolve_smem_bytes(N, 2) <= c->smemLimit) ? 2 : 1));
    }
    if (smemLaunch > c->smemLimit) { *occOut = 0; return SCB_OK; }
    int occ = 0;
    if (K == 4) CK((cudaError_t)scb_sim_occupancy_k4_tm0(S * 32, smemLaunch, &occ));
    else if (K == 2) CK((cudaError_t)scb_sim_occupancy_k2_tm0(S * 32, smemLaunch, &occ));
    else CK((cudaError_t)scb_sim_occupancy_k1_tm0(S * 32, smemLaunch, &occ));
    if (occ > 4) occ = 4;
    g->ctas = c->numSM * std::max(occ, 1);
    if (c->optMaxCtas > 0) g->ctas = std::min(g->ctas, c->optMaxCtas);
    /* snapshot in tensor memory when the columns of all co-resident CTAs fit into the SM's 512 */
    g->tmemCols = (tmemFits && cols * std::max(occ, 1) <= 512) ? cols : 0;
    g->tmemSlotsPerWarp = npw;
    *occOut = occ;
    return SCB_OK;
}

int pick_geometry(scb_ctx *c)
{
    const int N = c->N;
    if (c->optWords != 0 && c->optWords != 1 && c->optWords != 2 && c->optWords != 4) return fail(SCB_ERR_INVALID, "SCB_OPT_WORDS must be 0, 1, 2 or 4");
    if (scb_resolve_smem_bytes(N, 1) > c->smemLimit || scb_compile_smem_bytes(N) > c->smemLimit) {
        /* the state (or K0's staging) does not fit shared memory: populations are scored by the general simulator,
         * everything else (attractor analysis, importance scores) is refused by its own entry point */
        c->genericOnly = true; c->geomValid = false;
        c->fillWarps = 8; c->fillCtas = 8;
        return SCB_OK;
    }
#ifndef SCB_EMU
    CK((cudaError_t)scb_sim_set_smem_k1_tm0((int)c->smemLimit));
    CK((cudaError_t)scb_sim_set_smem_k2_tm0((int)c->smemLimit));
    CK((cudaError_t)scb_sim_set_smem_k4_tm0((int)c->smemLimit));
    CK((cudaError_t)scb_sim_set_smem_k1_tm1((int)c->smemLimit));
    CK((cudaError_t)scb_sim_set_smem_k2_tm1((int)c->smemLimit));
    CK((cudaError_t)scb_sim_set_smem_k4_tm1((int)c->smemLimit));
    CK(cudaFuncSetAttribute(scb_k_resolve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smemLimit));
    CK(cudaFuncSetAttribute(scb_k_resolve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smemLimit));
    CK(cudaFuncSetAttribute(scb_k_compile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smemLimit));
    /* several rule-compiler CTAs per SM: without the hint the driver sizes the shared-memory carve-out for two */
    CK(cudaFuncSetAttribute(scb_k_compile, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(scb_k_traj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smemLimit));
#endif
    /* Candidates: tile width K (words per lane) x warps per CTA.  The widest tile that fits wins: a narrower tile
     * doubles the instructions and the per-warp chain of shared-memory round trips per cell (measured on C3: two CTAs
     * of K = 2 tiles per SM 3.87 ms against one CTA of K = 4 tiles 3.03 ms).  At that width, a second CTA on the SM
     * fills the phases in which the first one has nothing in flight (barrier, table dispatch): measured at 90 nodes,
     * 2 CTAs x 8 warps 1.15 ms against 1 CTA x 16 warps 1.53 ms.  Then the most resident warps. */
    SimGeom best; int bestOcc = 0, bestScore = -1;
    int rc;
    const int kList[3] = {4, 2, 1}, sList[4] = {16, 8, 5, 4};
    int widest = 0;
    for (int ki = 0; ki < 3; ++ki) {
        const int K = kList[ki];
        if (c->optWords && K != c->optWords) continue;
        if (widest && K < widest) break;
        for (int si = 0; si < 4; ++si) {
            int S = c->optWarps ? c->optWarps : sList[si];
            if (!c->optWarps) S = std::min(S, std::max(1, (N + 3) / 4));
            SimGeom g; int occ = 0;
            if ((rc = sim_geometry(c, K, S, &g, &occ))) return rc;
            if (occ < 1) continue;
            if (!widest) widest = K;
            const int resident = std::min(occ * g.warps, SCB_SIM_MAX_WARPS);
            const int score = (std::min(occ, 2) >= 2 ? 1000 : 0) + resident * 10 + (K == 4 ? 2 : K == 2 ? 1 : 0);
            if (score > bestScore) { bestScore = score; best = g; bestOcc = occ; }
            if (c->optWarps) break;
        }
    }
    if (bestScore < 0) {
        /* the state does not fit shared memory: populations are scored by the general simulator, everything else
         * (attractor analysis, importance scores) is refused by its own entry point */
        c->genericOnly = true; c->geomValid = false;
        c->fillWarps = 8; c->fillCtas = 8;
        return SCB_OK;
    }
    c->genericOnly = false;
    c->main = best; c->occ = bestOcc;
    const int K = best.K;
    /* One launch covers all tiles; the last, partly filled tile is made cheap inside the kernel by
     * idling the lanes that hold no valid word (a separate narrow-tile launch was measured slower:
     * its few long-running tiles leave most SMs idle).  The tail geometry stays available. */
    const int WT = 32 * K;
    c->mainTiles = (c->W + WT - 1) / WT;
    c->tailTiles = 0;
    c->K = K; c->warps = c->main.warps; c->ctas = c->main.ctas; c->smemSim = c->main.smem;
    /* resolver items: whole tiles with the target state in tensor memory inside the simulator launch; else two
     * adjacent chunks (one for K = 1 or when three 64-word buffers do not fit) with three shared-memory buffers */
    c->resGroup = (K >= 2 && scb_resolve_smem_bytes(N, 2) <= c->smemLimit) ? 2 : 1;
    c->smemRes = scb_resolve_smem_bytes(N, c->resGroup);
    if (c->optFuse && c->main.tmemCols > 0 && c->optResGroup == 0) {
        c->resGroup = K;
        c->smemRes = scb_resolve_smem_bytes(N, K, true, c->main.warps);
    } else if (c->optResGroup && c->optResGroup <= K && scb_resolve_smem_bytes(N, c->optResGroup) <= c->smemLimit) {
        c->resGroup = c->optResGroup;
        c->smemRes = scb_resolve_smem_bytes(N, c->resGroup);
    }
    int occR = 1;
    c->resWarps = std::min(c->resGroup == 2 ? 24 : 16, std::max(c->resGroup, (N + 3) / 4));
    if (!(c->optFuse && c->main.tmemCols > 0)) {
        if (c->resGroup == 2) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occR, scb_k_resolve<2>, c->resWarps * 32, c->smemRes));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occR, scb_k_resolve<1>, c->resWarps * 32, c->smemRes));
    }
    c->resCtas = c->numSM * std::max(1, occR);
    if (c->optMaxCtas > 0) c->resCtas = std::min(c->resCtas, c->optMaxCtas);
    /* fill-forward: one warp per item, a lookup each */
    c->fillWarps = 8;
    c->fillCtas = 8;
    if (c->optMaxCtas > 0) c->fillCtas = std::min(c->fillCtas, c->optMaxCtas);
    size_t zwords = (size_t)c->main.ctas * N * 32 * K;
    if ((rc = c->dZ.ensure(zwords))) return rc;
    c->geomValid = true;
    return SCB_OK;
}

/* K0 on the uploaded population; `knock` is the array the rules read as knock-outs (the importance score
 * passes its knock-in list through the same slot, simulator.c:674) */
int launch_compile(scb_ctx *c, int sem, const int32_t *knock, bool wantImages = false, bool population = false)
{
    const int N = c->N, P = c->P;
    ScbCompileArgs ca;
    ca.P = P; ca.N = N; ca.indLen = c->indLen; ca.tablesPerIndividual = c->tpi; ca.Npad = N; ca.sem = sem;
    ca.individuals = c->dInd.p; ca.andLen = c->dAndLen.p; ca.parse = c->dParse.p;
    ca.an = c->dAn.p; ca.ai = c->dAi.p; ca.ko = knock; ca.ki = c->dKi.p; ca.koStride = c->koStride; ca.compact = c->compact;
    ca.descs = c->dDescs.p; ca.counts = c->dCounts.p; ca.diag = c->dDiag.p;
    ca.progs = c->dProgs.p; ca.progFrom = c->dProgFrom.p;
    ca.imgs = nullptr; ca.imgStride = 0; ca.imgWarps = 0; ca.imgRowBytes = 0;
    ca.slots = nullptr; ca.fitness = nullptr; ca.cost = nullptr; ca.pred = nullptr;
    int grid = P;
    if (population) {
        /* the population runs clear their own result sums and cost counters; a resident population compiles only the
         * replaced slots */
        ca.fitness = c->dFitness.p; ca.cost = c->dCost.p; ca.pred = c->dPred.p;
        if (c->delta) { ca.slots = c->dSlots.p; grid = c->nEval; }
    }
    if (wantImages) {
        /* the program images are cut for the simulator launch that follows: its warps, its row width */
        ca.imgStride = (int)scb_img_entries(N, c->main.warps);
        ca.imgWarps = c->main.warps; ca.imgRowBytes = 128u * (unsigned)c->main.K;
        int rc = c->dImgs.ensure((size_t)P * (SCB_NPROG + 1) * ca.imgStride);
        if (rc) return rc;
        ca.imgs = c->dImgs.p;
    }
    const int cthr = std::min(256, ((N + 31) / 32) * 32);
    SCB_LAUNCH(scb_k_compile, dim3((unsigned)grid), dim3((unsigned)cthr), scb_compile_smem_bytes(N), c->stream, ca);
    CK(cudaGetLastError());
    return SCB_OK;
}

int check_ctx(scb_ctx *c)
{
    if (!c) return fail(SCB_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return SCB_OK;
}

size_t result_bytes(const scb_ctx *c, int mode)
{
    if (mode == SCB_MODE_SUM) return (size_t)c->P * sizeof(long long);
    if (mode == SCB_MODE_NODEWISE) return (size_t)c->P * c->N * sizeof(int32_t);
    return (size_t)c->P * c->C * sizeof(int32_t);
}

}  // namespace

extern "C" {

const char *scb_last_error(void) { return g_err; }
int scb_version(void) { return SCB_VERSION; }

int scb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int scb_ctx_create(scb_ctx **out, int device, int nodeNum, int lenSamples, const void *binMat,
                   int elemBytes, int64_t rowStride, const int32_t *nodePositions)
{
    if (!out) return fail(SCB_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (nodeNum <= 0 || lenSamples <= 0) return fail(SCB_ERR_INVALID, "nodeNum and lenSamples must be positive");
    if (nodeNum > 65535) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d > 65535", nodeNum);
    if (!binMat || !nodePositions) return fail(SCB_ERR_INVALID, "null binMat / nodePositions");
    if (elemBytes != 4 && elemBytes != 1) return fail(SCB_ERR_INVALID, "elemBytes must be 4 (int32) or 1 (uint8)");
    if (rowStride < lenSamples) return fail(SCB_ERR_INVALID, "rowStride %lld < lenSamples %d", (long long)rowStride, lenSamples);
    for (int i = 0; i < nodeNum; ++i)
        if (nodePositions[i] < 0) return fail(SCB_ERR_INVALID, "nodePositions[%d] is negative", i);
    {   /* two nodes on one gene row alias one simData column in the reference (the later node overwrites the earlier
         * one in the attractor search and in the error); the packed kernels keep nodes independent, so refuse */
        std::vector<int32_t> sorted(nodePositions, nodePositions + nodeNum);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 1; i < nodeNum; ++i)
            if (sorted[i] == sorted[i - 1]) return fail(SCB_ERR_UNSUPPORTED, "nodePositions holds gene row %d twice", sorted[i]);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(SCB_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return fail(SCB_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    scb_ctx *c = new (std::nothrow) scb_ctx();
    if (!c) return fail(SCB_ERR_NOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return fail(SCB_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    c->numSM = prop.multiProcessorCount;
    c->smemLimit = prop.sharedMemPerBlockOptin;
    c->N = nodeNum; c->C = lenSamples;
    c->W = (lenSamples + 31) / 32;
    c->Wpad = ((c->W + 127) / 128) * 128;
    auto bail = [&](int rc) { scb_ctx_destroy(c); return rc; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "cudaStreamCreate failed"));
    for (auto &e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaEventCreate(&c->evRun0) != cudaSuccess || cudaEventCreate(&c->evRun1) != cudaSuccess ||
        cudaEventCreate(&c->evSim0) != cudaSuccess || cudaEventCreate(&c->evSim1) != cudaSuccess ||
        cudaEventCreate(&c->evK0) != cudaSuccess || cudaEventCreate(&c->evFill0) != cudaSuccess || cudaEventCreate(&c->evFill1) != cudaSuccess)
        return bail(fail(SCB_ERR_CUDA, "cudaEventCreate failed"));
    int rc;
    if ((rc = c->dX.ensure((size_t)c->N * c->Wpad))) return bail(rc);
    if ((rc = c->dDiag.ensure(1))) return bail(rc);
    if ((rc = c->dCounters.ensure(16))) return bail(rc);

    /* stage the N gene rows the network uses (row i = gene nodePositions[i]) and pack them */
    const size_t rowBytes = (size_t)lenSamples * elemBytes;
    DevBuf<unsigned char> stage;
    if ((rc = stage.ensure((size_t)c->N * rowBytes))) return bail(rc);
    bool contiguous = true;
    for (int i = 1; i < nodeNum; ++i) contiguous = contiguous && nodePositions[i] == nodePositions[0] + i;
    const unsigned char *src = static_cast<const unsigned char *>(binMat);
    cudaError_t ce = cudaSuccess;
    if (contiguous) {
        ce = cudaMemcpy2DAsync(stage.p, rowBytes, src + (size_t)nodePositions[0] * rowStride * elemBytes,
                               (size_t)rowStride * elemBytes, rowBytes, (size_t)nodeNum, cudaMemcpyHostToDevice, c->stream);
    } else {
        for (int i = 0; i < nodeNum && ce == cudaSuccess; ++i)
            ce = cudaMemcpyAsync(stage.p + (size_t)i * rowBytes, src + (size_t)nodePositions[i] * rowStride * elemBytes,
                                 rowBytes, cudaMemcpyHostToDevice, c->stream);
    }
    if (ce != cudaSuccess) { stage.release(); return bail(fail(SCB_ERR_CUDA, "H2D copy of binMat rows failed: %s", cudaGetErrorString(ce))); }
    cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), c->stream);
    ScbPackArgs pa;
    pa.src = stage.p; pa.elemBytes = elemBytes; pa.rowStride = lenSamples;
    pa.N = c->N; pa.C = c->C; pa.Wpad = c->Wpad; pa.X = c->dX.p; pa.diag = c->dDiag.p;
    const int wpb = 4;
    dim3 grid((unsigned)((c->Wpad / 32 + wpb - 1) / wpb), (unsigned)c->N);
    SCB_LAUNCH(scb_k_pack, grid, dim3(wpb * 32), 0, c->stream, pa);
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&c->hDiag, c->dDiag.p, sizeof(ScbDiag), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    stage.release();
    if (ce != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "pack kernel failed: %s", cudaGetErrorString(ce)));
    if (c->hDiag.nonbinary_cells)
        return bail(fail(SCB_ERR_INVALID, "binMat holds %llu values other than 0/1 in the network's rows", c->hDiag.nonbinary_cells));
    if ((rc = pick_geometry(c))) return bail(rc);
    *out = c;
    return SCB_OK;
}

/* Context from a CSR matrix: only the index/value slices of the N gene rows the network uses
 * are copied to the device; no dense detour. */
int scb_ctx_create_csr(scb_ctx **out, int device, int nodeNum, int lenSamples, const int32_t *indptr,
                       const int32_t *indices, const double *data, int nRows, const int32_t *nodePositions)
{
    if (!out) return fail(SCB_ERR_INVALID, "null output pointer");
    *out = nullptr;
    if (nodeNum <= 0 || lenSamples <= 0) return fail(SCB_ERR_INVALID, "nodeNum and lenSamples must be positive");
    if (nodeNum > 65535) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d > 65535", nodeNum);
    if (!indptr || !indices || !nodePositions) return fail(SCB_ERR_INVALID, "null CSR / nodePositions pointer");
    long long nnz = 0;
    {
        std::vector<int32_t> sorted(nodePositions, nodePositions + nodeNum);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 1; i < nodeNum; ++i)
            if (sorted[i] == sorted[i - 1]) return fail(SCB_ERR_UNSUPPORTED, "nodePositions holds gene row %d twice", sorted[i]);
    }
    for (int i = 0; i < nodeNum; ++i) {
        const int r = nodePositions[i];
        if (r < 0 || r >= nRows) return fail(SCB_ERR_INVALID, "nodePositions[%d] = %d outside the %d CSR rows", i, r, nRows);
        if (indptr[r + 1] < indptr[r]) return fail(SCB_ERR_INVALID, "CSR indptr is not monotone at row %d", r);
        nnz += indptr[r + 1] - indptr[r];
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(SCB_ERR_CUDA, "no CUDA device available");
    if (device < 0 || device >= ndev) return fail(SCB_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    scb_ctx *c = new (std::nothrow) scb_ctx();
    if (!c) return fail(SCB_ERR_NOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return fail(SCB_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    c->numSM = prop.multiProcessorCount;
    c->smemLimit = prop.sharedMemPerBlockOptin;
    c->N = nodeNum; c->C = lenSamples;
    c->W = (lenSamples + 31) / 32;
    c->Wpad = ((c->W + 127) / 128) * 128;
    auto bail = [&](int rc) { scb_ctx_destroy(c); return rc; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "cudaStreamCreate failed"));
    for (auto &e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaEventCreate(&c->evRun0) != cudaSuccess || cudaEventCreate(&c->evRun1) != cudaSuccess ||
        cudaEventCreate(&c->evSim0) != cudaSuccess || cudaEventCreate(&c->evSim1) != cudaSuccess ||
        cudaEventCreate(&c->evK0) != cudaSuccess || cudaEventCreate(&c->evFill0) != cudaSuccess || cudaEventCreate(&c->evFill1) != cudaSuccess)
        return bail(fail(SCB_ERR_CUDA, "cudaEventCreate failed"));
    int rc;
    if ((rc = c->dX.ensure((size_t)c->N * c->Wpad))) return bail(rc);
    if ((rc = c->dDiag.ensure(1))) return bail(rc);
    if ((rc = c->dCounters.ensure(16))) return bail(rc);
    /* gather the selected rows' slices on the host (pointer arithmetic only), one H2D each */
    std::vector<int32_t> rowOf((size_t)std::max<long long>(nnz, 1)), col((size_t)std::max<long long>(nnz, 1));
    std::vector<double> val(data ? (size_t)std::max<long long>(nnz, 1) : 0);
    long long k = 0;
    for (int i = 0; i < nodeNum; ++i) {
        const int r = nodePositions[i];
        for (int e = indptr[r]; e < indptr[r + 1]; ++e, ++k) {
            rowOf[k] = i; col[k] = indices[e];
            if (data) val[k] = data[e];
        }
    }
    DevBuf<int32_t> dRow, dCol;
    DevBuf<double> dVal;
    if ((rc = dRow.ensure(rowOf.size())) || (rc = dCol.ensure(col.size())) || (data && (rc = dVal.ensure(val.size())))) {
        dRow.release(); dCol.release(); dVal.release();
        return bail(rc);
    }
    cudaError_t ce = cudaMemsetAsync(c->dX.p, 0, (size_t)c->N * c->Wpad * 4, c->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), c->stream);
    if (ce == cudaSuccess && nnz) ce = cudaMemcpyAsync(dRow.p, rowOf.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && nnz) ce = cudaMemcpyAsync(dCol.p, col.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && nnz && data) ce = cudaMemcpyAsync(dVal.p, val.data(), (size_t)nnz * 8, cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && nnz) {
        ScbPackCsrArgs pa;
        pa.rowOfNnz = dRow.p; pa.col = dCol.p; pa.val = data ? dVal.p : nullptr; pa.nnz = nnz;
        pa.C = c->C; pa.Wpad = c->Wpad; pa.X = c->dX.p; pa.diag = c->dDiag.p;
        SCB_LAUNCH(scb_k_pack_csr, dim3((unsigned)((nnz + 255) / 256)), dim3(256), 0, c->stream, pa);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&c->hDiag, c->dDiag.p, sizeof(ScbDiag), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    dRow.release(); dCol.release(); dVal.release();
    if (ce != cudaSuccess) return bail(fail(SCB_ERR_CUDA, "CSR pack failed: %s", cudaGetErrorString(ce)));
    if (c->hDiag.nonbinary_cells)
        return bail(fail(SCB_ERR_INVALID, "the CSR matrix holds %llu values other than 0/1 in the network's rows", c->hDiag.nonbinary_cells));
    if ((rc = pick_geometry(c))) return bail(rc);
    *out = c;
    return SCB_OK;
}

#ifndef SCB_EMU
static void gather_release(scb_ctx *c);
#endif

int scb_ctx_destroy(scb_ctx *c)
{
    if (!c) return SCB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
#ifndef SCB_EMU
    gather_release(c);
    if (c->graphExec) cudaGraphExecDestroy(c->graphExec);
#endif
    c->dCost.release(); c->dOrder.release(); c->dSlots.release(); c->dStage.release(); c->dGScratch.release(); c->dPred.release(); c->dOrderAuto.release();
    c->dX.release(); c->dInd.release(); c->dAndLen.release(); c->dParse.release(); c->dAn.release();
    c->dAi.release(); c->dKo.release(); c->dKi.release(); c->dDescs.release(); c->dCounts.release(); c->dProgs.release(); c->dImgs.release(); c->dProgFrom.release();
    c->dFitness.release(); c->dNodewise.release(); c->dCellwise.release(); c->dCounters.release();
    c->dItems.release(); c->dReady.release(); c->dFillItems.release(); c->dChunkLast.release(); c->dChunkLastErr.release(); c->dChunkLastBits.release(); c->dResolveT.release(); c->dZ.release(); c->dDiag.release(); c->dAttr.release(); c->dAttrRaw.release(); c->dDist.release();
    c->dDecider.release(); c->dDeciderDist.release(); c->dFlush.release();
    c->dDedupeTable.release(); c->dDedupeOut.release(); c->dDedupeFlags.release(); c->dTraj.release(); c->dStates.release(); c->dRes.release(); c->dNStates.release(); c->dRecBits.release(); c->dRecOn.release();
    if (c->evSim0) cudaEventDestroy(c->evSim0);
    if (c->evK0) cudaEventDestroy(c->evK0);
    if (c->evFill0) cudaEventDestroy(c->evFill0);
    if (c->evFill1) cudaEventDestroy(c->evFill1);
    if (c->evSim1) cudaEventDestroy(c->evSim1);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->evRun0) cudaEventDestroy(c->evRun0);
    if (c->evRun1) cudaEventDestroy(c->evRun1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SCB_OK;
}

int scb_ctx_set_option(scb_ctx *c, int key, int64_t value)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    ++c->optVersion;
    switch (key) {
        case SCB_OPT_GRAPH: c->optGraph = value != 0; return SCB_OK;
        case SCB_OPT_GENERIC: c->optGeneric = value != 0; return SCB_OK;
        case SCB_OPT_AUTO_ORDER: c->optAutoOrder = value != 0; return SCB_OK;
        case SCB_OPT_RESOLVE_GROUP:
            if (value != 0 && value != 1 && value != 2) return fail(SCB_ERR_INVALID, "resolver group must be 0, 1 or 2");
            c->optResGroup = (int)value; break;
        case SCB_OPT_EARLY_EXIT: c->optEarly = value != 0; return SCB_OK;
        case SCB_OPT_SNAPSHOTS: c->optSnap = value != 0; break;
        case SCB_OPT_CHECK_EVERY:
            if (value < 1 || value > 99) return fail(SCB_ERR_INVALID, "check cadence must be in [1,99]");
            c->optChk = (int)value; return SCB_OK;
        case SCB_OPT_PERIOD_SEARCH:
            if (value < 0 || value > 99) return fail(SCB_ERR_INVALID, "period search must be in [0,99]");
            c->optSearch = (int)value; return SCB_OK;
        case SCB_OPT_SNAP_MIN_GAP:
            if (value < 1 || value > 99) return fail(SCB_ERR_INVALID, "snapshot gap must be in [1,99]");
            c->optMinGap = (int)value; return SCB_OK;
        case SCB_OPT_TMA: c->optTma = value != 0; return SCB_OK;
        case SCB_OPT_STEADY: c->optSteady = value != 0; return SCB_OK;
        case SCB_OPT_CHECK_FROM:
            if (value < -1 || value >= SCB_NPROG) return fail(SCB_ERR_INVALID, "check-from program must be in [-1,%d]", SCB_NPROG - 1);
            c->optChkFrom = (int)value; return SCB_OK;
        case SCB_OPT_TMEM: c->optTmem = value != 0; break;
        case SCB_OPT_FUSE: c->optFuse = value != 0; break;
        case SCB_OPT_SNAP_LO: c->snapLo = (unsigned long long)value; return SCB_OK;
        case SCB_OPT_SNAP_HI: c->snapHi = (unsigned long long)value; return SCB_OK;
        case SCB_OPT_WARPS:
            if (value < 0 || value > SCB_SIM_MAX_WARPS) return fail(SCB_ERR_INVALID, "warps must be in [0,%d]", SCB_SIM_MAX_WARPS);
            c->optWarps = (int)value; break;
        case SCB_OPT_MAX_CTAS:
            if (value < 0) return fail(SCB_ERR_INVALID, "max CTAs must be >= 0");
            c->optMaxCtas = (int)value; break;
        case SCB_OPT_WORDS: {
            const int old = c->optWords;
            c->optWords = (int)value;
            rc = pick_geometry(c);
            if (rc) { c->optWords = old; pick_geometry(c); }
            return rc;
        }
        default: return fail(SCB_ERR_INVALID, "unknown option %d", key);
    }
    return pick_geometry(c);
}

int scb_ctx_info(scb_ctx *c, int *nodeNum, int *lenSamples, int *wordsPerLane, int *warps, int *ctas, int *smemBytes)
{
    if (!c) return fail(SCB_ERR_INVALID, "null context");
    if (nodeNum) *nodeNum = c->N;
    if (lenSamples) *lenSamples = c->C;
    if (wordsPerLane) *wordsPerLane = c->K;
    if (warps) *warps = c->warps;
    if (ctas) *ctas = c->ctas;
    if (smemBytes) *smemBytes = (int)c->smemSim;
    return SCB_OK;
}

/* scb_pop_upload with `indRows` bit-strings (P, or 1 when the programs do not read them) and `koRows` knock-out
 * arrays (1 shared, or P: one per program, the batched importance score) */
static int pop_upload_impl(scb_ctx *c, int P, int indRows, const int32_t *individuals, int indLen, const int32_t *andLenList,
                           const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                           int tablesPerIndividual, int koRows, const int32_t *knockouts, const int32_t *knockins,
                           int compact = 0)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (P <= 0 || P > 65535 * 32) return fail(SCB_ERR_INVALID, "population size %d out of range", P);
    if (indLen < 0) return fail(SCB_ERR_INVALID, "indLen must be >= 0");
    if (!andLenList || !individualParse || !andNodes || !andNodeInvertList || !knockouts || !knockins || (!individuals && indLen > 0))
        return fail(SCB_ERR_INVALID, "null table pointer");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    const int N = c->N;
    const size_t tbl = compact ? (size_t)N * SCB_MAX_IN : (size_t)N * SCB_MAX_TERMS * SCB_MAX_IN;
    const size_t ntab = tablesPerIndividual ? (size_t)P : 1;
    if ((rc = c->dInd.ensure((size_t)P * std::max(indLen, 1)))) return rc;
    if ((rc = c->dAndLen.ensure(N)) || (rc = c->dParse.ensure(N)) || (rc = c->dKo.ensure((size_t)koRows * N)) || (rc = c->dKi.ensure(N))) return rc;
    if ((rc = c->dAn.ensure(ntab * tbl)) || (rc = c->dAi.ensure(ntab * tbl))) return rc;
    if (!c->genericOnly) {
        if ((rc = c->dDescs.ensure((size_t)P * N)) || (rc = c->dCounts.ensure(P))) return rc;
        if ((rc = c->dProgs.ensure((size_t)P * SCB_NPROG * (N + 1))) || (rc = c->dProgFrom.ensure((size_t)P * 8))) return rc;
    }
    if ((rc = c->dFitness.ensure(P))) return rc;
    const size_t chunks = (size_t)c->Wpad / 32;
    if ((rc = c->dReady.ensure((size_t)2 * P * chunks))) return rc;      /* ready flags of the resolver items, then the handed-back tiles */
    if ((rc = c->dItems.ensure((size_t)P * chunks)) || (rc = c->dFillItems.ensure((size_t)P * chunks)) ||
        (rc = c->dChunkLast.ensure((size_t)P * chunks)) || (rc = c->dChunkLastErr.ensure((size_t)P * chunks))) return rc;
    if (!c->geomValid && !c->genericOnly && (rc = pick_geometry(c))) return rc;
    if (!c->genericOnly) {   /* S_99 hand-over slots for the resolver: at most 64 MiB */
        const size_t per = (size_t)N * 32 * c->resGroup;
        const size_t want = std::min<size_t>((size_t)P * chunks, std::max<size_t>(64, ((size_t)64 << 20) / (per * 4)));
        if ((rc = c->dResolveT.ensure(want * per))) return rc;
        c->resolveTCap = (unsigned)(c->dResolveT.cap / per);
    }
    cudaStream_t st = c->stream;
    if (indLen > 0) CK(cudaMemcpyAsync(c->dInd.p, individuals, (size_t)indRows * indLen * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dAndLen.p, andLenList, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dParse.p, individualParse, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dKo.p, knockouts, (size_t)koRows * N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dKi.p, knockins, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dAn.p, andNodes, ntab * tbl * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->dAi.p, andNodeInvertList, ntab * tbl * 4, cudaMemcpyHostToDevice, st));
    c->koStride = koRows > 1 ? N : 0;
    if (P != c->P) c->haveOrder = false;            /* an order is a permutation of the population it was measured on */
    c->P = P; c->indLen = indLen; c->tpi = tablesPerIndividual ? 1 : 0;
    c->compact = compact ? 1 : 0;
    c->ran = false; c->delta = false; c->nEval = 0;
    return SCB_OK;
}

int scb_pop_upload(scb_ctx *c, int P, const int32_t *individuals, int indLen, const int32_t *andLenList,
                   const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                   int tablesPerIndividual, const int32_t *knockouts, const int32_t *knockins)
{
    return pop_upload_impl(c, P, P, individuals, indLen, andLenList, individualParse, andNodes, andNodeInvertList,
                           tablesPerIndividual, 1, knockouts, knockins);
}

/* the population as the GA needs to keep it -- bit-strings plus each node's upstream triple
 * (what ruleMaker.__update_upstream changes, ruleMaker.py:300-335) instead of a deep copy of the whole model per
 * individual (ruleMaker.py:650).  K0 rebuilds the [7][3] term tables on the device. */
int scb_pop_upload_compact(scb_ctx *c, int P, const int32_t *individuals, int indLen, const int32_t *andLenList,
                           const int32_t *individualParse, const int32_t *upstream, const int32_t *upstreamInvert,
                           int perIndividual, const int32_t *knockouts, const int32_t *knockins)
{
    return pop_upload_impl(c, P, P, individuals, indLen, andLenList, individualParse, upstream, upstreamInvert,
                           perIndividual, 1, knockouts, knockins, 1);
}

/* SURVEY 8f N4, resident population: `n` individuals of the uploaded population are replaced in place and the next
 * scb_pop_run(SCB_MODE_SUM) compiles and simulates exactly those; every other slot keeps the fitness of the run that
 * last scored it.  The rows go through a staging buffer (one copy per array) and a scatter kernel on the context's
 * stream. */
int scb_pop_update(scb_ctx *c, int n, const int32_t *slots, const int32_t *individuals, const int32_t *andNodes,
                   const int32_t *andNodeInvertList)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (c->P <= 0) return fail(SCB_ERR_INVALID, "no population uploaded");
    if (c->koStride) return fail(SCB_ERR_UNSUPPORTED, "batched importance programs cannot be updated in place");
    if (n < 0 || n > c->P) return fail(SCB_ERR_INVALID, "%d replaced individuals for a population of %d", n, c->P);
    if (n > 0 && (!slots || (!individuals && c->indLen > 0))) return fail(SCB_ERR_INVALID, "null argument");
    if (n > 0 && c->tpi && (!andNodes || !andNodeInvertList)) return fail(SCB_ERR_INVALID, "the population has per-individual tables: the replaced individuals need theirs");
    if (c->delta && !c->ran && c->nEval > 0) return fail(SCB_ERR_INVALID, "the previous update has not been run yet");
    std::vector<char> seen((size_t)c->P, 0);
    for (int i = 0; i < n; ++i) {
        if (slots[i] < 0 || slots[i] >= c->P || seen[slots[i]]) return fail(SCB_ERR_INVALID, "slot list is not a set of distinct indices below %d", c->P);
        seen[slots[i]] = 1;
    }
    const int N = c->N;
    const size_t tbl = c->compact ? (size_t)N * SCB_MAX_IN : (size_t)N * SCB_MAX_TERMS * SCB_MAX_IN;
    const size_t rowInd = (size_t)std::max(c->indLen, 0);
    cudaStream_t st = c->stream;
    if (n > 0) {
        const size_t need = (size_t)n * (rowInd + (c->tpi ? 2 * tbl : 0));
        if ((rc = c->dSlots.ensure((size_t)c->P)) || (rc = c->dStage.ensure(std::max<size_t>(need, 1)))) return rc;
        /* the previous update's scatter kernels have read the staging buffer: same stream, in order */
        CK(cudaMemcpyAsync(c->dSlots.p, slots, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        int32_t *sInd = c->dStage.p, *sAn = sInd + (size_t)n * rowInd, *sAi = sAn + (size_t)n * tbl;
        const unsigned grid = (unsigned)std::min(n, 1024);
        if (rowInd) {
            CK(cudaMemcpyAsync(sInd, individuals, (size_t)n * rowInd * 4, cudaMemcpyHostToDevice, st));
            SCB_LAUNCH(scb_k_scatter_rows, dim3(grid), dim3(128), 0, st, c->dInd.p, (const int32_t *)sInd, (const int32_t *)c->dSlots.p, (int)rowInd, n);
        }
        if (c->tpi) {
            CK(cudaMemcpyAsync(sAn, andNodes, (size_t)n * tbl * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(sAi, andNodeInvertList, (size_t)n * tbl * 4, cudaMemcpyHostToDevice, st));
            SCB_LAUNCH(scb_k_scatter_rows, dim3(grid), dim3(256), 0, st, c->dAn.p, (const int32_t *)sAn, (const int32_t *)c->dSlots.p, (int)tbl, n);
            SCB_LAUNCH(scb_k_scatter_rows, dim3(grid), dim3(256), 0, st, c->dAi.p, (const int32_t *)sAi, (const int32_t *)c->dSlots.p, (int)tbl, n);
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));                /* the caller's arrays (pageable or not) may be reused on return */
    }
    if (c->ran || c->delta) {
        /* every other slot holds a valid fitness: only the replaced ones are scored.  (Before the first run of an
         * upload the whole population is still to be scored, replaced rows included.) */
        c->delta = true; c->nEval = n;
        c->ran = false;
    }
    return SCB_OK;
}

static int pop_run_impl(scb_ctx *c, int mode, int sem);

int scb_pop_run(scb_ctx *c, int mode)
{
    if (mode < 0 || mode > 2) return fail(SCB_ERR_INVALID, "unknown mode %d", mode);
    return pop_run_impl(c, mode, SCB_SEM_C);
}

/* K0 + K2 + K3 on the uploaded population; mode may also be the internal SCB_MODE_IMPORTANCE (with
 * sem = SCB_SEM_IMPORTANCE), whose per-node counts land in the nodewise buffer */
static int pop_enqueue(scb_ctx *c, int mode, int sem, bool capturing);

/* the uploaded population through the general simulator (scb_k_generic) + the fill-forward kernel */
static int generic_run(scb_ctx *c, int mode)
{
    int rc;
    const int N = c->N, P = c->P;
    const int cpi = (c->C + 1023) / 1024;
    cudaStream_t st = c->stream;
    if (c->koStride) return fail(SCB_ERR_UNSUPPORTED, "batched importance programs have no general path");
    if (mode == SCB_MODE_NODEWISE && ((rc = c->dNodewise.ensure((size_t)P * N)) || (rc = c->dChunkLastBits.ensure((size_t)P * cpi * ((N + 31) / 32))))) return rc;
    if (mode == SCB_MODE_CELLWISE && (rc = c->dCellwise.ensure((size_t)P * c->C))) return rc;
    const long long items = (long long)P * cpi;
    const size_t perCta = (size_t)3 * N * 32;                              /* words */
    long long ctas = std::min<long long>(items, (long long)c->numSM * 4);
    ctas = std::min<long long>(ctas, std::max<long long>(1, (long long)(((size_t)1 << 28) / perCta)));   /* at most 1 GiB of state */
    if (c->optMaxCtas > 0) ctas = std::min<long long>(ctas, c->optMaxCtas);
    if ((rc = c->dGScratch.ensure((size_t)ctas * perCta))) return rc;
    CK(cudaEventRecord(c->evRun0, st));
    CK(cudaMemsetAsync(c->dFitness.p, 0, (size_t)P * 8, st));
    if (mode == SCB_MODE_NODEWISE) CK(cudaMemsetAsync(c->dNodewise.p, 0, (size_t)P * N * 4, st));
    CK(cudaMemsetAsync(c->dCounters.p, 0, 16 * sizeof(unsigned int), st));
    CK(cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), st));
    CK(cudaEventRecord(c->evK0, st));
    CK(cudaEventRecord(c->evSim0, st));
    ScbGenericArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.X = c->dX.p; ga.N = N; ga.C = c->C; ga.Wpad = c->Wpad; ga.P = P; ga.indLen = c->indLen;
    ga.tablesPerIndividual = c->tpi; ga.compact = c->compact; ga.mode = mode;
    ga.individuals = c->dInd.p; ga.andLen = c->dAndLen.p; ga.parse = c->dParse.p; ga.an = c->dAn.p; ga.ai = c->dAi.p;
    ga.ko = c->dKo.p; ga.ki = c->dKi.p;
    ga.scratch = c->dGScratch.p;
    ga.fitness = c->dFitness.p; ga.nodewise = c->dNodewise.p; ga.cellwise = c->dCellwise.p;
    ga.chunkLast = c->dChunkLast.p; ga.chunkLastErr = c->dChunkLastErr.p; ga.chunkLastBits = c->dChunkLastBits.p;
    ga.chunksPerInd = cpi;
    ga.fillCount = c->dCounters.p + 3; ga.fillItems = c->dFillItems.p;
    ga.cursor = c->dCounters.p + 0;
    ga.diag = c->dDiag.p;
    SCB_LAUNCH(scb_k_generic, dim3((unsigned)ctas), dim3(256), SCB_GENERIC_SMEM, st, ga);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->evSim1, st));
    ScbResolveArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.s.N = N; ra.s.C = c->C; ra.s.Wpad = c->Wpad; ra.s.P = P; ra.s.mode = mode;
    ra.s.fitness = c->dFitness.p; ra.s.nodewise = c->dNodewise.p; ra.s.cellwise = c->dCellwise.p;
    ra.s.chunkLast = c->dChunkLast.p; ra.s.chunkLastErr = c->dChunkLastErr.p; ra.s.chunkLastBits = c->dChunkLastBits.p;
    ra.s.chunksPerInd = cpi; ra.s.fillCount = c->dCounters.p + 3; ra.s.fillItems = c->dFillItems.p; ra.s.diag = c->dDiag.p;
    ra.itemCursor = c->dCounters.p + 4;
    CK(cudaEventRecord(c->evFill0, st));
    SCB_LAUNCH(scb_k_fill, dim3((unsigned)c->fillCtas), dim3((unsigned)c->fillWarps * 32), 0, st, ra);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->evFill1, st));
    CK(cudaEventRecord(c->evRun1, st));
    c->lastMode = mode; c->ran = true; c->delta = false; c->lastGeneric = true;
    return SCB_OK;
}

static int pop_run_impl(scb_ctx *c, int mode, int sem)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (c->P <= 0) return fail(SCB_ERR_INVALID, "no population uploaded");
    c->lastSem = sem;
    if (c->genericOnly || (c->optGeneric && sem == SCB_SEM_C && mode <= 2)) {
        if (sem != SCB_SEM_C || mode > 2) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d does not fit the %zu-byte shared-memory state buffers: only population scoring is available", c->N, c->smemLimit);
        return generic_run(c, mode);
    }
    c->lastGeneric = false;
    if (!c->geomValid && (rc = pick_geometry(c))) return rc;
    const int N = c->N, P = c->P;
    cudaStream_t st = c->stream;
    if (c->delta && mode != SCB_MODE_SUM) return fail(SCB_ERR_UNSUPPORTED, "after scb_pop_update only the fitness sums (SCB_MODE_SUM) are kept up to date");
    if (c->delta && c->nEval == 0) {                       /* nothing was replaced: every fitness is current */
        CK(cudaEventRecord(c->evRun0, st)); CK(cudaEventRecord(c->evRun1, st));
        CK(cudaEventRecord(c->evSim0, st)); CK(cudaEventRecord(c->evSim1, st));
        CK(cudaEventRecord(c->evK0, st)); CK(cudaEventRecord(c->evFill0, st)); CK(cudaEventRecord(c->evFill1, st));
        c->lastMode = mode; c->ran = true; c->delta = false;
        return SCB_OK;
    }
    const bool wantNodewise = mode == SCB_MODE_NODEWISE || mode == SCB_MODE_IMPORTANCE;
    /* every allocation happens here, before anything is enqueued (or captured) */
    if (wantNodewise && (rc = c->dNodewise.ensure((size_t)P * N))) return rc;
    if (mode == SCB_MODE_NODEWISE && (rc = c->dChunkLastBits.ensure((size_t)P * ((c->C + 1023) / 1024) * ((N + 31) / 32)))) return rc;
    if (mode == SCB_MODE_CELLWISE && (rc = c->dCellwise.ensure((size_t)P * c->C))) return rc;
    if ((rc = c->dImgs.ensure((size_t)P * (SCB_NPROG + 1) * scb_img_entries(N, c->main.warps)))) return rc;
    if ((rc = c->dCost.ensure((size_t)2 * P)) || (rc = c->dPred.ensure(P)) || (rc = c->dOrderAuto.ensure(P))) return rc;
#ifndef SCB_EMU
    if (c->optGraph) {
        /* One launch per generation: the clears, K0, K2 (+ resolver), the fill-forward kernel and the peer exchange are
         * captured once and replayed for as long as nothing the graph refers to has moved (buffers, population shape,
         * options, peers).  The generation's inputs change in place (uploads into the same buffers). */
        unsigned long long key = 1469598103934665603ull;
        auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
        mix(g_allocEpoch.load()); mix(c->optVersion); mix((unsigned long long)P); mix((unsigned long long)mode); mix((unsigned long long)sem);
        mix((unsigned long long)c->indLen); mix((unsigned long long)c->tpi); mix((unsigned long long)c->compact); mix((unsigned long long)c->koStride);
        mix(c->haveOrder ? 1ull : 0ull); mix(c->delta ? 1ull + (unsigned long long)c->nEval : 0ull); mix(c->gConnected ? (unsigned long long)(1 + c->gRank + 64 * c->gWorld + 4096ull * c->gWidth) : 0ull);
        if (!c->graphExec || key != c->graphKey) {
            if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = pop_enqueue(c, mode, sem, true);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(SCB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ci = cudaGraphInstantiate(&c->graphExec, graph, 0);
            cudaGraphDestroy(graph);
            if (ci != cudaSuccess) { c->graphExec = nullptr; return fail(SCB_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ci)); }
            c->graphKey = key;
        }
        CK(cudaEventRecord(c->evRun0, st));
        CK(cudaGraphLaunch(c->graphExec, st));
        CK(cudaEventRecord(c->evRun1, st));
        CK(cudaEventRecord(c->evSim0, st));          /* the per-kernel split needs SCB_OPT_GRAPH = 0 */
        CK(cudaEventRecord(c->evSim1, st));
        CK(cudaEventRecord(c->evK0, st)); CK(cudaEventRecord(c->evFill0, st)); CK(cudaEventRecord(c->evFill1, st));
        c->lastMode = mode; c->ran = true; c->delta = false;
        return SCB_OK;
    }
#endif
    CK(cudaEventRecord(c->evRun0, st));
    if ((rc = pop_enqueue(c, mode, sem, false))) return rc;
    CK(cudaEventRecord(c->evRun1, st));
    c->lastMode = mode; c->ran = true; c->delta = false;
    return SCB_OK;
}

/* the generation on the context's stream: clears, K0, K2 (+ in-kernel resolver), resolver / fill-forward kernels and,
 * when peers are connected, the fitness exchange */
static int pop_enqueue(scb_ctx *c, int mode, int sem, bool capturing)
{
    int rc;
    const int N = c->N, P = c->P;
    cudaStream_t st = c->stream;
    const bool wantNodewise = mode == SCB_MODE_NODEWISE || mode == SCB_MODE_IMPORTANCE;
    /* (K0 clears the fitness sums and cost counters of the individuals it compiles) */
    if (wantNodewise) CK(cudaMemsetAsync(c->dNodewise.p, 0, (size_t)P * N * 4, st));
    CK(cudaMemsetAsync(c->dCounters.p, 0, 16 * sizeof(unsigned int), st));
    CK(cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), st));

    if (!capturing) CK(cudaEventRecord(c->evK0, st));
    if ((rc = launch_compile(c, sem, c->dKo.p, true, true))) return rc;
    /* a population nobody gave an order for is handed out by the cost K0 predicts (individual-major, costliest first) */
    const bool autoOrder = c->optAutoOrder && !c->haveOrder && !c->delta && c->optSteady && P <= SCB_ORDER_MAX_P && P > 1;
    if (autoOrder) {
        SCB_LAUNCH(scb_k_order, dim3(1), dim3((unsigned)std::min(1024, ((P + 31) / 32) * 32)), (size_t)P * 4, st, (const unsigned int *)c->dPred.p, c->dOrderAuto.p, P);
        CK(cudaGetLastError());
    }

    ScbSimArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.X = c->dX.p; sa.N = N; sa.C = c->C; sa.Wpad = c->Wpad;
    sa.P = P; sa.Npad = N; sa.mode = mode;
    sa.descs = c->dDescs.p; sa.counts = c->dCounts.p;
    sa.progs = c->dProgs.p; sa.progFrom = c->dProgFrom.p; sa.useSteady = c->optSteady;
    sa.imgs = c->dImgs.p; sa.imgStride = (int)scb_img_entries(N, c->main.warps);
    sa.cost = c->dCost.p; sa.order = c->delta ? c->dSlots.p : (c->haveOrder ? c->dOrder.p : (autoOrder ? c->dOrderAuto.p : nullptr));
    sa.nEval = c->delta ? c->nEval : P;
    sa.chkFromProg = c->optChkFrom;
    sa.firstSnap = 1000;
    for (int t = SCB_LAST - 1; t >= 1; --t)
        if (t < 64 ? ((c->snapLo >> t) & 1ull) : ((c->snapHi >> (t - 64)) & 1ull)) sa.firstSnap = t;
    sa.fitness = c->dFitness.p; sa.nodewise = c->dNodewise.p; sa.cellwise = c->dCellwise.p;
    sa.resolveCount = c->dCounters.p + 1;
    sa.resolveItems = c->dItems.p; sa.resolveCap = (unsigned)std::min<size_t>(std::min(c->dItems.cap, c->dReady.cap / 2), 0x7fffffffu);
    sa.resolveT = c->dResolveT.p; sa.resolveGroup = c->resGroup;
    sa.resolveTCap = (unsigned)(c->dResolveT.cap / ((size_t)N * 32 * c->resGroup));
    sa.chunkLast = c->dChunkLast.p; sa.chunksPerInd = (c->C + 1023) / 1024;
    sa.chunkLastErr = c->dChunkLastErr.p; sa.chunkLastBits = c->dChunkLastBits.p;
    sa.fillCount = c->dCounters.p + 3; sa.fillItems = c->dFillItems.p;
    sa.zscratch = c->dZ.p;
    sa.flags = (c->optEarly ? SCB_KF_EARLY : 0u) | (c->optSnap ? SCB_KF_SNAP : 0u);
    sa.periodSearch = c->optSnap ? c->optSearch : 0;
    sa.snapMinGap = c->optMinGap;
    {   /* the static schedule of the flagged steps as 128-bit masks over the steps 1 .. 99 */
        unsigned long long chk[2] = {0ull, 0ull}, snp[2] = {0ull, 0ull}, zc[2] = {0ull, 0ull};
        auto setb = [](unsigned long long *m, int t) { m[t >> 6] |= 1ull << (t & 63); };
        auto getb = [](const unsigned long long *m, int t) { return (m[t >> 6] >> (t & 63)) & 1ull; };
        const unsigned long long given[2] = {c->snapLo, c->snapHi};
        if (c->optEarly) for (int t = 3; t < SCB_LAST; ++t) if (t % c->optChk == 0) setb(chk, t);
        setb(chk, SCB_LAST);                                       /* the last state is always compared: E2 or the resolver */
        int last = -1;
        for (int t = 1; t <= SCB_LAST; ++t) {
            if (c->optSnap && last >= 0 && (t - last >= c->optMinGap || t == SCB_LAST)) setb(zc, t);
            if (c->optSnap && t < SCB_LAST && getb(given, t)) { setb(snp, t); last = t; }
        }
        sa.chkLo = chk[0]; sa.chkHi = chk[1]; sa.snapLo = snp[0]; sa.snapHi = snp[1]; sa.zcmpLo = zc[0]; sa.zcmpHi = zc[1];
    }
    sa.diag = c->dDiag.p;
    /* the simulator resolves queued groups in its own tail when a single launch covers all tiles */
    const bool fuse = c->optFuse && c->tailTiles == 0;
    sa.fuseResolve = fuse ? 1 : 0;
    sa.useTma = c->optTma;
    sa.tilesDone = c->dCounters.p + 6; sa.readyFlags = c->dReady.p; sa.resolveCursor = c->dCounters.p + 2;
    sa.retCount = c->dCounters.p + 8; sa.retCursor = c->dCounters.p + 9; sa.retItems = c->dReady.p + sa.resolveCap;
    if (fuse) CK(cudaMemsetAsync(c->dReady.p, 0, (size_t)2 * sa.resolveCap * sizeof(unsigned int), st));
    if (!capturing) CK(cudaEventRecord(c->evSim0, st));
    for (int part = 0; part < 2; ++part) {
        const SimGeom &g = part == 0 ? c->main : c->tail;
        const int tiles = part == 0 ? c->mainTiles : c->tailTiles;
        if (tiles <= 0) continue;
        sa.tiles = tiles;
        sa.wordBase = part == 0 ? 0 : c->mainTiles * 32 * c->K;
        sa.workCounter = c->dCounters.p + (part == 0 ? 0 : 5);
        sa.zstride = (unsigned long long)N * 32 * g.K;
        sa.tmemCols = g.tmemCols; sa.tmemSlotsPerWarp = g.tmemSlotsPerWarp;
        const long long items = (long long)sa.nEval * tiles;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(items, g.ctas));
        const dim3 block((unsigned)g.warps * 32);
        const size_t smemLaunch = fuse ? std::max(g.smem, c->smemRes) : g.smem;
        if (smemLaunch > c->smemLimit) return fail(SCB_ERR_UNSUPPORTED, "simulator launch needs %zu bytes of shared memory", smemLaunch);
        sa.smemBytes = (unsigned)smemLaunch;
        int le;
        if (sa.tmemCols > 0) {
            if (g.K == 4) le = scb_sim_launch_k4_tm1(dim3(grid), block, smemLaunch, st, sa);
            else if (g.K == 2) le = scb_sim_launch_k2_tm1(dim3(grid), block, smemLaunch, st, sa);
            else le = scb_sim_launch_k1_tm1(dim3(grid), block, smemLaunch, st, sa);
        } else {
            if (g.K == 4) le = scb_sim_launch_k4_tm0(dim3(grid), block, smemLaunch, st, sa);
            else if (g.K == 2) le = scb_sim_launch_k2_tm0(dim3(grid), block, smemLaunch, st, sa);
            else le = scb_sim_launch_k1_tm0(dim3(grid), block, smemLaunch, st, sa);
        }
        CK((cudaError_t)le);
    }
    if (!capturing) CK(cudaEventRecord(c->evSim1, st));

    ScbResolveArgs ra;
    ra.s = sa; ra.itemCursor = c->dCounters.p + 2;
    if (!fuse) {
        if (c->resGroup == 2) SCB_LAUNCH(scb_k_resolve<2>, dim3((unsigned)c->resCtas), dim3((unsigned)c->resWarps * 32), c->smemRes, st, ra);
        else SCB_LAUNCH(scb_k_resolve<1>, dim3((unsigned)c->resCtas), dim3((unsigned)c->resWarps * 32), c->smemRes, st, ra);
    }
    CK(cudaGetLastError());
    ra.itemCursor = c->dCounters.p + 4;
    if (!capturing) CK(cudaEventRecord(c->evFill0, st));
    SCB_LAUNCH(scb_k_fill, dim3((unsigned)c->fillCtas), dim3((unsigned)c->fillWarps * 32), 0, st, ra);
    CK(cudaGetLastError());
    if (!capturing) CK(cudaEventRecord(c->evFill1, st));
    if (c->gConnected && mode == SCB_MODE_SUM) {
        /* the generation's all-gather: this rank's shard into every rank's buffer over peer memory */
        ScbGatherArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.fitness = c->dFitness.p; ga.count = std::min(P, c->gWidth); ga.width = c->gWidth; ga.rank = c->gRank; ga.world = c->gWorld;
        for (int r = 0; r < c->gWorld; ++r) { ga.peerBuf[r] = c->gPeerBuf[r]; ga.peerFlag[r] = c->gPeerFlag[r]; }
        ga.generation = c->gFlags + SCB_MAX_PEERS; ga.timeout = c->gFlags + SCB_MAX_PEERS + 1;
#ifndef SCB_EMU
        SCB_LAUNCH(scb_k_gather, dim3(1), dim3(256), 0, st, ga);
        CK(cudaGetLastError());
#endif
    }
    return SCB_OK;
}

/* ---- cost-ordered work list ------------------------------------------------------------------------------------ */
int scb_pop_costs(scb_ctx *c, uint32_t *costs, uint32_t *tileMax)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->ran || !costs) return fail(SCB_ERR_INVALID, "no finished run / null output");
    if (c->lastGeneric) return fail(SCB_ERR_UNSUPPORTED, "the general simulator keeps no per-individual costs");
    CK(cudaMemcpyAsync(costs, c->dCost.p, (size_t)c->P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (tileMax) CK(cudaMemcpyAsync(tileMax, c->dCost.p + c->P, (size_t)c->P * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SCB_OK;
}

int scb_pop_set_order(scb_ctx *c, const int32_t *order, int P)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!order) { c->haveOrder = false; return SCB_OK; }
    if (P <= 0) return fail(SCB_ERR_INVALID, "population size must be positive");
    std::vector<char> seen((size_t)P, 0);
    for (int i = 0; i < P; ++i) {
        if (order[i] < 0 || order[i] >= P || seen[order[i]]) return fail(SCB_ERR_INVALID, "order is not a permutation of 0..%d", P - 1);
        seen[order[i]] = 1;
    }
    if ((rc = c->dOrder.ensure(P))) return rc;
    CK(cudaMemcpyAsync(c->dOrder.p, order, (size_t)P * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));            /* `order` is the caller's temporary */
    c->haveOrder = true;
    return SCB_OK;
}

/* ---- fitness exchange over peer memory (one process per GPU of a box) ------------------------------------------ */
#ifndef SCB_EMU
static void gather_release(scb_ctx *c)
{
    for (int r = 0; r < c->gWorld; ++r) {
        if (r == c->gRank) continue;
        if (c->gPeerBuf[r]) cudaIpcCloseMemHandle(c->gPeerBuf[r]);
        if (c->gPeerFlag[r]) cudaIpcCloseMemHandle(c->gPeerFlag[r]);
    }
    if (c->gBuf) cudaFree(c->gBuf);
    if (c->gFlags) cudaFree(c->gFlags);
    c->gBuf = nullptr; c->gFlags = nullptr; c->gConnected = false; c->gWorld = 0;
    for (int r = 0; r < SCB_MAX_PEERS; ++r) { c->gPeerBuf[r] = nullptr; c->gPeerFlag[r] = nullptr; }
    ++g_allocEpoch;
}
#endif

int scb_gather_init(scb_ctx *c, int world, int rank, int width, unsigned char *handles)
{
#ifdef SCB_EMU
    (void)c; (void)world; (void)rank; (void)width; (void)handles;
    return fail(SCB_ERR_UNSUPPORTED, "peer-memory exchange needs CUDA devices");
#else
    int rc = check_ctx(c);
    if (rc) return rc;
    if (world < 1 || world > SCB_MAX_PEERS || rank < 0 || rank >= world || width < 1 || !handles)
        return fail(SCB_ERR_INVALID, "bad gather geometry (world %d, rank %d, width %d)", world, rank, width);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    CK(cudaStreamSynchronize(c->stream));
    gather_release(c);
    c->gWorld = world; c->gRank = rank; c->gWidth = width;
    const size_t bufBytes = (size_t)2 * world * width * 8, flagBytes = (size_t)(SCB_MAX_PEERS + 2) * 4;
    CK(cudaMalloc(reinterpret_cast<void **>(&c->gBuf), bufBytes));
    CK(cudaMalloc(reinterpret_cast<void **>(&c->gFlags), flagBytes));
    CK(cudaMemset(c->gBuf, 0, bufBytes));
    CK(cudaMemset(c->gFlags, 0, flagBytes));
    cudaIpcMemHandle_t hb, hf;
    CK(cudaIpcGetMemHandle(&hb, c->gBuf));
    CK(cudaIpcGetMemHandle(&hf, c->gFlags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handles, &hb, 64);
    memcpy(handles + 64, &hf, 64);
    return SCB_OK;
#endif
}

int scb_gather_connect(scb_ctx *c, const unsigned char *allHandles)
{
#ifdef SCB_EMU
    (void)c; (void)allHandles;
    return fail(SCB_ERR_UNSUPPORTED, "peer-memory exchange needs CUDA devices");
#else
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->gBuf || !allHandles) return fail(SCB_ERR_INVALID, "scb_gather_init first");
    for (int r = 0; r < c->gWorld; ++r) {
        if (r == c->gRank) { c->gPeerBuf[r] = c->gBuf; c->gPeerFlag[r] = c->gFlags; continue; }
        cudaIpcMemHandle_t hb, hf;
        memcpy(&hb, allHandles + (size_t)r * 128, 64);
        memcpy(&hf, allHandles + (size_t)r * 128 + 64, 64);
        void *pb = nullptr, *pf = nullptr;
        CK(cudaIpcOpenMemHandle(&pb, hb, cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
        c->gPeerBuf[r] = static_cast<unsigned long long *>(pb);
        c->gPeerFlag[r] = static_cast<unsigned int *>(pf);
    }
    c->gConnected = true;
    return SCB_OK;
#endif
}

int scb_gather_result(scb_ctx *c, int64_t *out)
{
#ifdef SCB_EMU
    (void)c; (void)out;
    return fail(SCB_ERR_UNSUPPORTED, "peer-memory exchange needs CUDA devices");
#else
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->gConnected || !out) return fail(SCB_ERR_INVALID, "no connected gather / null output");
    const size_t n = (size_t)c->gWorld * c->gWidth;
    std::vector<unsigned long long> both(2 * n);
    unsigned int tail[2] = {0u, 0u};
    CK(cudaMemcpyAsync(both.data(), c->gBuf, 2 * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tail, c->gFlags + SCB_MAX_PEERS, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&c->hDiag, c->dDiag.p, sizeof(ScbDiag), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (tail[1]) return fail(SCB_ERR_CUDA, "a peer did not take part in the fitness exchange (timeout)");
    /* this rank's own shard went out as K0 + K2 computed it: garbage must not pass for fitness (the other ranks' input
     * problems are reported by their own scb_gather_result) */
    if (c->hDiag.bad_index_nodes || c->hDiag.term_overflow_nodes)
        return fail(SCB_ERR_INVALID, "the population of this rank holds malformed rules (input index outside [0,%d), more than 7 terms): use scb_pop_fetch for the details", c->N);
    if (c->hDiag.unsupported_nodes)
        return fail(SCB_ERR_UNSUPPORTED, "%llu rules of this rank's population use more than 3 distinct inputs: such populations are scored by the general simulator at scb_pop_fetch and cannot take part in the in-graph exchange", c->hDiag.unsupported_nodes);
    memcpy(out, both.data() + (size_t)(tail[0] & 1u) * n, n * 8);
    return SCB_OK;
#endif
}

int scb_gather_close(scb_ctx *c)
{
#ifndef SCB_EMU
    if (!c) return SCB_OK;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    gather_release(c);
#else
    (void)c;
#endif
    return SCB_OK;
}

int scb_pop_fetch(scb_ctx *c, int mode, void *out, scb_stats *stats)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->ran) return fail(SCB_ERR_INVALID, "scb_pop_run has not been called for this population");
    if (mode != c->lastMode && mode != SCB_MODE_SUM) return fail(SCB_ERR_INVALID, "mode %d was not computed by the last run (mode %d)", mode, c->lastMode);
    cudaStream_t st = c->stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (out) {
            const void *src = mode == SCB_MODE_SUM ? (const void *)c->dFitness.p
                              : mode == SCB_MODE_NODEWISE ? (const void *)c->dNodewise.p : (const void *)c->dCellwise.p;
            CK(cudaMemcpyAsync(out, src, result_bytes(c, mode), cudaMemcpyDeviceToHost, st));
        }
        CK(cudaMemcpyAsync(&c->hDiag, c->dDiag.p, sizeof(ScbDiag), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        /* K0 met rules over more than three distinct inputs: legal input of the reference, which the compiled path
         * cannot express -- the population is scored again by the general simulator (exact, slow) */
        if (attempt == 0 && c->hDiag.unsupported_nodes && !c->lastGeneric && c->lastSem == SCB_SEM_C && c->lastMode <= 2 && !c->koStride) {
            if ((rc = generic_run(c, c->lastMode))) return rc;
            continue;
        }
        break;
    }
    const ScbDiag &d = c->hDiag;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->not_found_cells = (int64_t)d.not_found_cells;
        stats->undefined_leading = (int64_t)d.undefined_leading;
        stats->empty_rule_nodes = (int64_t)d.empty_rule_nodes;
        stats->unsupported_nodes = (int64_t)d.unsupported_nodes;
        stats->bad_index_nodes = (int64_t)d.bad_index_nodes;
        stats->term_overflow_nodes = (int64_t)d.term_overflow_nodes;
        stats->resolver_chunks = (int64_t)d.resolver_chunks;
        stats->steps_executed = (int64_t)d.steps_executed;
        stats->tiles = (int64_t)d.tiles;
        stats->state_rows_moved = (int64_t)d.rows_moved;
        stats->cyc_prologue = (int64_t)d.cyc_prologue;
        stats->cyc_steps = (int64_t)d.cyc_steps;
        stats->cyc_epilogue = (int64_t)d.cyc_epilogue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->evRun0, c->evRun1) == cudaSuccess) stats->kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evSim0, c->evSim1) == cudaSuccess) stats->sim_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evRun0, c->evK0) == cudaSuccess) stats->clear_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evK0, c->evSim0) == cudaSuccess) stats->compile_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evSim1, c->evFill0) == cudaSuccess) stats->resolve_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evFill0, c->evFill1) == cudaSuccess) stats->fill_ms = ms;
        if (cudaEventElapsedTime(&ms, c->evFill1, c->evRun1) == cudaSuccess) stats->gather_ms = ms;
    }
    if (d.bad_index_nodes) return fail(SCB_ERR_INVALID, "%llu (individual,node) rules reference an input outside [0,%d)", d.bad_index_nodes, c->N);
    if (d.term_overflow_nodes) return fail(SCB_ERR_INVALID, "%llu (individual,node) rules span more than 7 AND-terms, leave the bit-string, or (compact upload) disagree with andLenList", d.term_overflow_nodes);
    if (d.unsupported_nodes) return fail(SCB_ERR_UNSUPPORTED, "%llu (individual,node) rules use more than 3 distinct inputs", d.unsupported_nodes);
    return SCB_OK;
}

/* test / diagnostics hook: the compiled programs of individual p of the last run (uint32[N][4] each: three input
 * nodes, node | table << 24; counts = nodes of arity 3, 2, 1, 0): the full program and the reduced program `k`
 * (0 .. SCB_NPROG-1, the last one is the fixed point) with the first step it may compute (0: not emitted) */
int scb_debug_programs(scb_ctx *c, int p, int k, uint32_t *descs, int32_t *counts, uint32_t *descsK, int32_t *countsK, int32_t *fromK)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->ran || p < 0 || p >= c->P || k < 0 || k >= SCB_NPROG) return fail(SCB_ERR_INVALID, "no such program in the last run");
    if (c->genericOnly || c->lastGeneric) return fail(SCB_ERR_UNSUPPORTED, "the general simulator compiles no programs");
    const size_t n = (size_t)c->N;
    CK(cudaStreamSynchronize(c->stream));
    const uint4 *pk = c->dProgs.p + ((size_t)p * SCB_NPROG + k) * (n + 1);
    unsigned char from[8];
    if (descs) CK(cudaMemcpy(descs, c->dDescs.p + p * n, n * 16, cudaMemcpyDeviceToHost));
    if (counts) CK(cudaMemcpy(counts, c->dCounts.p + p, 16, cudaMemcpyDeviceToHost));
    if (descsK) CK(cudaMemcpy(descsK, pk, n * 16, cudaMemcpyDeviceToHost));
    if (countsK) CK(cudaMemcpy(countsK, pk + n, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(from, c->dProgFrom.p + (size_t)p * 8, 8, cudaMemcpyDeviceToHost));
    if (fromK) *fromK = from[k];
    return SCB_OK;
}

#ifdef SCB_TRACE
/* measurement build only: the step trace of the last fetched run (see ScbDiag::trace) */
extern "C" int scb_debug_trace(scb_ctx *c, unsigned int *trace, unsigned int *flags)
{
    if (!c || !trace || !flags) return fail(SCB_ERR_INVALID, "null argument");
    memcpy(trace, c->hDiag.trace, sizeof(c->hDiag.trace));
    memcpy(flags, c->hDiag.traceFlags, sizeof(c->hDiag.traceFlags));
    memcpy(flags + 100, c->hDiag.exitHist, sizeof(c->hDiag.exitHist));
    return SCB_OK;
}
/* timeline of the last launch: up to 16384 x {kind | CTA << 8, start ns, end ns, extra}; returns the number of events */
extern "C" int scb_debug_events(scb_ctx *c, unsigned int *events)
{
    if (!c || !events) return -1;
    memcpy(events, c->hDiag.evLog, sizeof(c->hDiag.evLog));
    return (int)std::min(c->hDiag.evCount, 16384u);
}
#endif

int scb_pop_result_dev(scb_ctx *c, int mode, void **devPtr, int64_t *bytes)
{
    if (!c || !devPtr) return fail(SCB_ERR_INVALID, "null argument");
    if (mode < 0 || mode > 2) return fail(SCB_ERR_INVALID, "unknown mode %d", mode);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    if (!c->ran) return fail(SCB_ERR_INVALID, "scb_pop_run has not been called for this population");
    if (mode != c->lastMode && mode != SCB_MODE_SUM) return fail(SCB_ERR_INVALID, "mode %d was not computed by the last run (mode %d)", mode, c->lastMode);
    *devPtr = mode == SCB_MODE_SUM ? (void *)c->dFitness.p : mode == SCB_MODE_NODEWISE ? (void *)c->dNodewise.p : (void *)c->dCellwise.p;
    if (bytes) *bytes = (int64_t)result_bytes(c, mode);
    return SCB_OK;
}

int scb_eval_population(scb_ctx *c, int P, const int32_t *individuals, int indLen, const int32_t *andLenList,
                        const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                        int tablesPerIndividual, const int32_t *knockouts, const int32_t *knockins, int mode,
                        void *out, scb_stats *stats)
{
    if (!c) return fail(SCB_ERR_INVALID, "null context");
    /* upload + run + fetch under ONE lock: concurrent callers of a shared context (the reference drives local
     * search from a ThreadPool, singleCell.py:490-492) must not interleave their populations */
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    int rc = scb_pop_upload(c, P, individuals, indLen, andLenList, individualParse, andNodes, andNodeInvertList,
                            tablesPerIndividual, knockouts, knockins);
    if (rc) return rc;
    if ((rc = scb_pop_run(c, mode))) return rc;
    return scb_pop_fetch(c, mode, out, stats);
}

int scb_ctx_stream(scb_ctx *c, void **s)
{
    if (!c || !s) return fail(SCB_ERR_INVALID, "null argument");
    *s = (void *)c->stream;
    return SCB_OK;
}

int scb_ctx_sync(scb_ctx *c)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return SCB_OK;
}

int scb_ctx_event_record(scb_ctx *c, int slot)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (slot < 0 || slot >= 8) return fail(SCB_ERR_INVALID, "event slot out of range");
    CK(cudaEventRecord(c->ev[slot], c->stream));
    return SCB_OK;
}

int scb_ctx_event_elapsed_ms(scb_ctx *c, int s0, int s1, float *ms)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (s0 < 0 || s0 >= 8 || s1 < 0 || s1 >= 8 || !ms) return fail(SCB_ERR_INVALID, "bad event slot");
    CK(cudaEventSynchronize(c->ev[s1]));
    CK(cudaEventElapsedTime(ms, c->ev[s0], c->ev[s1]));
    return SCB_OK;
}

int scb_ctx_flush_l2(scb_ctx *c)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    const size_t bytes = (size_t)256 << 20;              /* > 126 MB of L2 */
    if ((rc = c->dFlush.ensure(bytes, false))) return rc;
    CK(cudaMemsetAsync(c->dFlush.p, 0x5a, bytes, c->stream));
    return SCB_OK;
}

int scb_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return fail(SCB_ERR_INVALID, "bad argument");
    CK(cudaMallocHost(ptr, (size_t)bytes));
    return SCB_OK;
}

int scb_host_free(void *ptr)
{
    if (ptr) CK(cudaFreeHost(ptr));
    return SCB_OK;
}

int scb_eval_local_search(scb_ctx *c, const int32_t *individual, int indLen, const int32_t *andLenList,
                          const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                          const int32_t *knockouts, const int32_t *knockins, int node, int32_t *nodeErrors,
                          int maxVariants, int *nVariants)
{
    if (!c) return fail(SCB_ERR_INVALID, "null context");
    if (!individual || !individualParse || !nodeErrors) return fail(SCB_ERR_INVALID, "null argument");
    const int N = c->N;
    if (node < 0 || node >= N) return fail(SCB_ERR_INVALID, "node %d out of range", node);
    /* ruleMaker.py:1237-1242: the node's bits are [parse[node], parse[node+1]) (indLen for the last) */
    const int start = individualParse[node];
    const int end = (node == N - 1) ? indLen : individualParse[node + 1];
    const int k = end - start;
    if (k < 1 || k > SCB_MAX_TERMS || start < 0 || end > indLen) return fail(SCB_ERR_INVALID, "node %d has %d rule bits", node, k);
    const int nv = (1 << k) - 1;
    if (nVariants) *nVariants = nv;
    if (maxVariants < nv) return fail(SCB_ERR_INVALID, "nodeErrors holds %d entries, %d needed", maxVariants, nv);
    std::vector<int32_t> pop((size_t)nv * indLen);
    for (int v = 1; v <= nv; ++v) {
        int32_t *row = pop.data() + (size_t)(v - 1) * indLen;
        memcpy(row, individual, (size_t)indLen * 4);
        /* ruleMaker.py:796-802 (__bitList reverses bin(v)): bit j of the node = (v >> j) & 1 */
        for (int j = 0; j < k; ++j) row[start + j] = (v >> j) & 1;
    }
    std::vector<int32_t> nodewise((size_t)nv * N);
    int rc = scb_eval_population(c, nv, pop.data(), indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                                 knockouts, knockins, SCB_MODE_NODEWISE, nodewise.data(), nullptr);
    if (rc) return rc;
    for (int v = 0; v < nv; ++v) nodeErrors[v] = nodewise[(size_t)v * N + node];
    return SCB_OK;
}

/* All nodes at once (the driver loop singleCell.py:478-510 calls __checkNodePossibilities node by node, or through a
 * ThreadPool): the variants of every node with 1 <= k <= 7 rule bits are scored in ONE launch.
 * offsets[i] .. offsets[i+1] index nodeErrors for node i (offsets has N + 1 entries; nodes with k == 0 get none). */
int scb_eval_local_search_all(scb_ctx *c, const int32_t *individual, int indLen, const int32_t *andLenList,
                              const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                              const int32_t *knockouts, const int32_t *knockins, int32_t *nodeErrors, int maxVariants,
                              int32_t *offsets)
{
    if (!c) return fail(SCB_ERR_INVALID, "null context");
    if (!individual || !individualParse || !nodeErrors || !offsets) return fail(SCB_ERR_INVALID, "null argument");
    const int N = c->N;
    int total = 0;
    for (int node = 0; node < N; ++node) {
        const int start = individualParse[node];
        const int end = (node == N - 1) ? indLen : individualParse[node + 1];
        const int k = end - start;
        if (k < 0 || k > SCB_MAX_TERMS || start < 0 || end > indLen) return fail(SCB_ERR_INVALID, "node %d has %d rule bits", node, k);
        offsets[node] = total;
        total += k > 0 ? (1 << k) - 1 : 0;
    }
    offsets[N] = total;
    if (maxVariants < total) return fail(SCB_ERR_INVALID, "nodeErrors holds %d entries, %d needed", maxVariants, total);
    if (total == 0) return SCB_OK;
    std::vector<int32_t> pop((size_t)total * indLen);
    for (int node = 0; node < N; ++node) {
        const int start = individualParse[node];
        const int nv = offsets[node + 1] - offsets[node];
        const int k = (node == N - 1 ? indLen : individualParse[node + 1]) - start;
        for (int v = 1; v <= nv; ++v) {
            int32_t *row = pop.data() + (size_t)(offsets[node] + v - 1) * indLen;
            memcpy(row, individual, (size_t)indLen * 4);
            for (int j = 0; j < k; ++j) row[start + j] = (v >> j) & 1;            /* __bitList, ruleMaker.py:796-802 */
        }
    }
    std::vector<int32_t> nodewise((size_t)total * N);
    int rc = scb_eval_population(c, total, pop.data(), indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                                 knockouts, knockins, SCB_MODE_NODEWISE, nodewise.data(), nullptr);
    if (rc) return rc;
    for (int node = 0; node < N; ++node)
        for (int v = offsets[node]; v < offsets[node + 1]; ++v) nodeErrors[v] = nodewise[(size_t)v * N + node];
    return SCB_OK;
}

/* distance + decider for nAttr attractors packed in c->dAttr ([nAttr][NW4 * 4]) */
static int hamming_run(scb_ctx *c, int nAttr, int NW4, int32_t *dist, int32_t *decider, int32_t *deciderDistance, unsigned int *badValues)
{
    int rc;
    const int N = c->N, C = c->C;
    if ((rc = c->dDecider.ensure(C)) || (rc = c->dDeciderDist.ensure(C))) return rc;
    if (dist && (rc = c->dDist.ensure((size_t)C * nAttr))) return rc;
    cudaStream_t st = c->stream;
    ScbHammingArgs ha;
    ha.X = c->dX.p; ha.N = N; ha.C = C; ha.Wpad = c->Wpad; ha.A = nAttr; ha.attr = c->dAttr.p;
    ha.dist = dist ? c->dDist.p : nullptr; ha.decider = c->dDecider.p; ha.deciderDistance = c->dDeciderDist.p;
    const int nthr = 128;
    const dim3 hgrid((unsigned)((C + nthr - 1) / nthr));
    switch (NW4) {
        case 1: SCB_LAUNCH(scb_k_hamming<1>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
        case 2: SCB_LAUNCH(scb_k_hamming<2>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
        case 3: SCB_LAUNCH(scb_k_hamming<3>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
        case 4: SCB_LAUNCH(scb_k_hamming<4>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
        case 6: SCB_LAUNCH(scb_k_hamming<6>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
        default: SCB_LAUNCH(scb_k_hamming<8>, hgrid, dim3(nthr), SCB_HAM_SMEM_WORDS * 4, st, ha); break;
    }
    CK(cudaGetLastError());
    if (badValues) CK(cudaMemcpyAsync(badValues, c->dCounters.p + 10, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    if (dist) CK(cudaMemcpyAsync(dist, c->dDist.p, (size_t)C * nAttr * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(decider, c->dDecider.p, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(deciderDistance, c->dDeciderDist.p, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SCB_OK;
}

static int hamming_width(int NW)
{
    const int NW4 = (NW + 3) / 4;
    return NW4 <= 4 ? NW4 : (NW4 <= 6 ? 6 : 8);           /* the kernel's instantiations; rows are zero padded to that width */
}

int scb_hamming_argmin(scb_ctx *c, const int32_t *attractors, int nAttr, int32_t *dist, int32_t *decider,
                       int32_t *deciderDistance)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!attractors || nAttr <= 0 || !decider || !deciderDistance) return fail(SCB_ERR_INVALID, "bad argument");
    const int N = c->N, NW = (N + 31) / 32;
    if (NW > SCB_HAM_MAXW) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d > %d for the Hamming kernel", N, SCB_HAM_MAXW * 32);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    /* the attractors go up as the caller holds them and are packed on the device (a host loop over 32 x the packed
     * size was most of this call's time) */
    const int NW4 = hamming_width(NW), NWP = NW4 * 4;
    if ((rc = c->dAttrRaw.ensure((size_t)nAttr * N)) || (rc = c->dAttr.ensure((size_t)nAttr * NWP))) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(c->dAttrRaw.p, attractors, (size_t)nAttr * N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->dCounters.p + 10, 0, sizeof(unsigned int), st));
    {
        const long long warps = (long long)nAttr * NWP;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((warps + 7) / 8, (long long)c->numSM * 16));
        SCB_LAUNCH(scb_k_pack_attr, dim3(grid), dim3(256), 0, st, (const int32_t *)c->dAttrRaw.p, c->dAttr.p, nAttr, N, NWP, c->dCounters.p + 10);
        CK(cudaGetLastError());
    }
    unsigned int badValues = 0;
    if ((rc = hamming_run(c, nAttr, NW4, dist, decider, deciderDistance, &badValues))) return rc;
    if (badValues) return fail(SCB_ERR_INVALID, "%u attractor entries are not 0/1", badValues);
    return SCB_OK;
}

int scb_hamming_argmin_packed(scb_ctx *c, const uint32_t *attractors, int nAttr, int32_t *dist, int32_t *decider,
                              int32_t *deciderDistance)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!attractors || nAttr <= 0 || !decider || !deciderDistance) return fail(SCB_ERR_INVALID, "bad argument");
    const int N = c->N, NW = (N + 31) / 32;
    if (NW > SCB_HAM_MAXW) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d > %d for the Hamming kernel", N, SCB_HAM_MAXW * 32);
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    const int NW4 = hamming_width(NW), NWP = NW4 * 4;
    if ((rc = c->dAttr.ensure((size_t)nAttr * NWP))) return rc;
    cudaStream_t st = c->stream;
    if (NWP == NW) CK(cudaMemcpyAsync(c->dAttr.p, attractors, (size_t)nAttr * NW * 4, cudaMemcpyHostToDevice, st));
    else {
        /* one contiguous copy, then re-strided on the device (a pitched copy from pageable memory goes row by row) */
        if ((rc = c->dAttrRaw.ensure((size_t)nAttr * NW))) return rc;
        CK(cudaMemcpyAsync(c->dAttrRaw.p, attractors, (size_t)nAttr * NW * 4, cudaMemcpyHostToDevice, st));
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(((long long)nAttr * NWP + 255) / 256, (long long)c->numSM * 8));
        SCB_LAUNCH(scb_k_restride, dim3(grid), dim3(256), 0, st, (const uint32_t *)c->dAttrRaw.p, c->dAttr.p, (long long)nAttr, NW, NWP);
        CK(cudaGetLastError());
    }
    return hamming_run(c, nAttr, NW4, dist, decider, deciderDistance, nullptr);
}

}  /* extern "C" */

#include "scb_legacy.inl"
#include "scb_microbench.inl"
