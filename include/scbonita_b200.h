/*
 * scbonita_b200.h -- C-ABI of the B200-native replacement for scBONITA's rule-inference
 * fitness hot path (the compiled C simulator the reference loads with
 * ctypes.cdll.LoadLibrary("./simulator.so"), singleCell.py:453, ruleMaker.py:979).
 *
 * Two groups of entry points:
 *
 *  1. the reference's own exported symbols, same names and argument lists, so the unchanged
 *     Python (ruleMaker.__evaluateByNode, ruleMaker.py:1077-1221) keeps working when this
 *     library is installed as ./simulator.so:
 *         scSyncBool        replaces simulator.c:350-485
 *         importanceScore   replaces simulator.c:555-755
 *         cluster           replaces simulator.c:487-552
 *  2. a context / batch API with runtime sizes (no compile-time NODE/CELL, simulator.c:6-8)
 *     that packs the binarized matrix once and scores a whole GA population, or all local-search
 *     variants, in one launch.  This is what the reference-side binding in INTEGRATION.md
 *     calls from the two list comprehensions at ruleMaker.py:982-985 / 1030-1033 and from
 *     __checkNodePossibilities (ruleMaker.py:1235-1266).
 *
 * Everything is plain C: pointers, sizes, int status codes.  All int32 tables are in the
 * reference's layouts (what __evaluateByNode marshals, ruleMaker.py:1089-1129):
 *     individual        int32[indLen]            rule bit-string
 *     andLenList        int32[N]                 number of AND-terms per node
 *     individualParse   int32[N] (or N+1)        start of each node's bits in `individual`
 *     andNodes          int32[N][7][3]           inputs per AND-term, padded with -1 / [0,0,0]
 *     andNodeInvertList int32[N][7][3]           inversion flag per input (int comparison, Q3)
 *     knockouts/knockins int32[N]                ==1 knocks the node out / in
 *     binMat            int32[rows][rowStride]   binarized expression, gene-major
 *     nodePositions     int32[N]                 gene row of node i
 *
 * The library needs a CUDA device (sm_100a build); there is no CPU fallback: every entry
 * point fails with SCB_ERR_CUDA when no device is usable.
 */
#ifndef SCBONITA_B200_H
#define SCBONITA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_VERSION 100

/* status codes (0 = success); scb_last_error() has the text for the calling thread */
#define SCB_OK               0
#define SCB_ERR_INVALID     -1   /* bad argument / contract violation (sizes, simSteps != 100, ...)  */
#define SCB_ERR_CUDA        -2   /* CUDA runtime error or no device                                   */
#define SCB_ERR_UNSUPPORTED -3   /* a rule needs > 3 distinct inputs, or N exceeds on-chip capacity   */
#define SCB_ERR_NOMEM       -4

/* output shapes of a population evaluation */
#define SCB_MODE_SUM       0   /* int64[P]      sum over cells of the cellwise error = Python's sum(errors),
                                  ruleMaker.py:1220-1221 (GA fitness)                                 */
#define SCB_MODE_NODEWISE  1   /* int32[P][N]   errors[i] per node (localSearch != 0, simulator.c:477-479) */
#define SCB_MODE_CELLWISE  2   /* int32[P][C]   errors[c] per cell (localSearch == 0, simulator.c:481)     */

/* scb_ctx_set_option keys */
#define SCB_OPT_EARLY_EXIT   1  /* 1 (default): stop a cell tile once its future is proven periodic  */
#define SCB_OPT_SNAPSHOTS    2  /* 1 (default): resolve cycles with in-flight snapshot compares       */
#define SCB_OPT_WARPS        3  /* warps per CTA of the simulator kernel (0 = auto)                   */
#define SCB_OPT_MAX_CTAS     4  /* cap on persistent CTAs (0 = SM count x occupancy)                  */
#define SCB_OPT_WORDS        5  /* 32-bit words per lane (0 = auto: 4, 2 or 1 by node count)          */
#define SCB_OPT_CHECK_EVERY  6  /* period-<=2 check cadence in steps (default 12)                     */
#define SCB_OPT_PERIOD_SEARCH 7 /* extra snapshot-compare steps spent looking for a tile-wide period  */
#define SCB_OPT_SNAP_LO      8  /* bit t (t < 64): take a snapshot when computing S_t                 */
#define SCB_OPT_SNAP_HI      9  /* bit t-64 for steps 64..98 (default: snapshots at 36 and 66)        */
#define SCB_OPT_SNAP_MIN_GAP 10 /* first snapshot compare this many steps after a snapshot (default 8) */
#define SCB_OPT_TMEM         11 /* 1 (default): keep the snapshot in tensor memory, 0: in global scratch    */
#define SCB_OPT_TMA          13 /* 1 (default): load the S_0 tile with TMA bulk copies (cp.async.bulk + mbarrier)   */
#define SCB_OPT_STEADY       14 /* 1 (default): steps after the constants have spread run the constant-propagated program */
#define SCB_OPT_CHECK_FROM   15 /* k (default 3): no period-<=2 check before reduced program k may run; -1: check from step 3 */
#define SCB_OPT_GRAPH        16 /* 1 (default): a generation is one CUDA graph launch; 0: plain launches (per-kernel timing in scb_stats) */
#define SCB_OPT_GENERIC      17 /* 1: score populations with the general simulator (any rule width, any node count; slow) even where the compiled path applies; default 0 */
#define SCB_OPT_RESOLVE_GROUP 18 /* chunks per resolver item inside the simulator launch: 0 (default) = a whole tile with the target state in tensor memory; 1 or 2 = narrower items with three shared-memory buffers (shorter critical path, more SM time) */
#define SCB_OPT_AUTO_ORDER   19 /* 1 (default): without an scb_pop_set_order, tiles are handed out by the cost K0 predicts for each individual (rows of its fixed-point program); 0: index order, tile-major */
#define SCB_OPT_FUSE         12 /* 1 (default): CTAs that run out of tiles resolve queued groups in the same launch */

typedef struct scb_ctx scb_ctx;

/* counters of one evaluation (all optional diagnostics; results do not depend on them) */
typedef struct scb_stats {
    int64_t not_found_cells;      /* cells whose state 99 repeats no earlier state        */
    int64_t undefined_leading;    /* not-found cells before the first found one: the reference reads  */
                                  /* uninitialised memory there; defined as 0 here, as in the oracle  */
    int64_t empty_rule_nodes;     /* (individual,node) pairs with no selected AND-term (Q9): value 0  */
    int64_t unsupported_nodes;    /* rules with > 3 distinct inputs -> SCB_ERR_UNSUPPORTED            */
    int64_t bad_index_nodes;      /* input index outside [0,N) -> SCB_ERR_INVALID                     */
    int64_t term_overflow_nodes;  /* more than 7 AND-terms for a node (Q10) -> SCB_ERR_INVALID        */
    int64_t resolver_chunks;      /* 1024-cell chunks that needed the exact second-pass resolver      */
    int64_t steps_executed;       /* simulated (tile, step) pairs, for executed-work accounting       */
    int64_t tiles;                /* tiles simulated (individuals x cell tiles)                       */
    float   kernel_ms;            /* device time of the last evaluation (CUDA events on the stream)   */
    float   sim_ms;               /* of which the simulator kernel K2 alone                           */
    int64_t cyc_prologue;         /* SM clocks summed over the K2 CTAs: work fetch + program + S_0 tile load */
    int64_t cyc_steps;            /*   ... the synchronous steps                                      */
    int64_t cyc_epilogue;         /*   ... chunk flags, error reduction, outputs                      */
    int64_t state_rows_moved;     /* state rows (128 x wordsPerLane bytes each) the simulator's steps loaded + stored */
    /* the other parts of kernel_ms (SCB_OPT_GRAPH = 0 only; all 0 when the generation runs as one graph):            */
    float   clear_ms;             /*   result / counter clears                                        */
    float   compile_ms;           /*   K0, the rule compiler                                          */
    float   resolve_ms;           /*   the resolver kernel (0 when it runs inside K2)                 */
    float   fill_ms;              /*   the fill-forward kernel                                        */
    float   gather_ms;            /*   the peer-memory fitness exchange                               */
    float   reserved_;
} scb_stats;

const char *scb_last_error(void);
int scb_version(void);
int scb_device_count(void);

/* ---- context: the packed cell matrix of one (dataset, network) pair on one GPU -------------
 * Packs rows nodePositions[i] of binMat (elemBytes 4: int32 as the reference passes it,
 * ruleMaker.py:1109-1111; elemBytes 1: uint8) into node-major 32-bit words on the device
 * (kernel K1).  binMat values must be 0/1.  rowStride is in elements (reference: CELL,
 * simulator.c:8).  The caller's buffers are not retained. */
int scb_ctx_create(scb_ctx **ctx, int device, int nodeNum, int lenSamples,
                   const void *binMat, int elemBytes, int64_t rowStride,
                   const int32_t *nodePositions);
/* The same from a CSR matrix (scipy.sparse.csr_matrix: indptr[nRows+1], indices, data), which is how the
 * reference stores binMat (singleCell.py:25-90): only the slices of the N gene rows are copied, there is no
 * dense detour (the reference densifies 20000 x 15000 on every evaluation, ruleMaker.py:1109-1111).
 * data may be NULL (stored element = 1); values must be 0 or 1; columns >= lenSamples are ignored. */
int scb_ctx_create_csr(scb_ctx **ctx, int device, int nodeNum, int lenSamples,
                       const int32_t *indptr, const int32_t *indices, const double *data,
                       int nRows, const int32_t *nodePositions);
int scb_ctx_destroy(scb_ctx *ctx);
int scb_ctx_set_option(scb_ctx *ctx, int key, int64_t value);
int scb_ctx_info(scb_ctx *ctx, int *nodeNum, int *lenSamples, int *wordsPerLane, int *warps,
                 int *ctas, int *smemBytes);

/* ---- whole-population evaluation, host buffers in and out (replaces P calls of scSyncBool,
 * ruleMaker.py:982-985).  individuals: int32[P][indLen]; andNodes / andNodeInvertList are
 * [P][N][7][3] when tablesPerIndividual != 0 (each individual owns a model copy,
 * ruleMaker.py:650) else one shared [N][7][3].  simSteps is fixed at 100.
 * out: see SCB_MODE_*.  stats may be NULL.  Synchronous. */
int scb_eval_population(scb_ctx *ctx, int P, const int32_t *individuals, int indLen,
                        const int32_t *andLenList, const int32_t *individualParse,
                        const int32_t *andNodes, const int32_t *andNodeInvertList,
                        int tablesPerIndividual, const int32_t *knockouts,
                        const int32_t *knockins, int mode, void *out, scb_stats *stats);

/* ---- the same, split so that inputs can stay resident in HBM between launches ------------
 * scb_pop_upload copies the reference-format tables to the device (async on the context's
 * stream); scb_pop_run launches rule compilation + simulation + resolver (async);
 * scb_pop_fetch waits and copies the result of `mode` to the host.  scb_pop_result_dev
 * exposes the device buffer (int64[P] / int32[P][N] / int32[P][C]) for a collective.
 * The host buffers given to scb_pop_upload must stay valid until scb_pop_fetch or scb_ctx_sync returns
 * (with pinned memory, scb_host_alloc, the copy is truly asynchronous). */
int scb_pop_upload(scb_ctx *ctx, int P, const int32_t *individuals, int indLen,
                   const int32_t *andLenList, const int32_t *individualParse,
                   const int32_t *andNodes, const int32_t *andNodeInvertList,
                   int tablesPerIndividual, const int32_t *knockouts, const int32_t *knockins);
/* The same upload in the compact form a GA driver can keep per individual instead of a deep copy of the model
 * (ruleMaker.py:650): upstream / upstreamInvert are int32[P][N][3] (perIndividual != 0) or one shared [N][3] --
 * each node's at most three upstream nodes as a prefix, -1 after the last, and their inversion flags.  The
 * [7][3] AND-term tables are rebuilt on the device in the reference's order [abc, ab, ac, a, bc, b, c]
 * (ruleMaker.py:188-200, padded like __updateCpointers, :337-360); andLenList[i] must equal 2^k - 1 for a node
 * with k upstream nodes, else scb_pop_fetch fails with SCB_ERR_INVALID.  24 instead of 168 bytes per node. */
int scb_pop_upload_compact(scb_ctx *ctx, int P, const int32_t *individuals, int indLen,
                           const int32_t *andLenList, const int32_t *individualParse,
                           const int32_t *upstream, const int32_t *upstreamInvert,
                           int perIndividual, const int32_t *knockouts, const int32_t *knockins);
/* Resident population.  The GA replaces part of its population every generation
 * (ruleMaker.py:982-1006: only offspring whose fitness is invalid are re-scored) while this library keeps the
 * population in device memory: scb_pop_update overwrites `n` individuals in place -- slots[i] (distinct, < P) gets
 * bit-string individuals[i] and, when the population was uploaded with per-individual tables, andNodes[i] /
 * andNodeInvertList[i] in the format of that upload ([N][7][3], or [N][3] upstream triples after
 * scb_pop_upload_compact; ignored for shared tables) -- and the next scb_pop_run(SCB_MODE_SUM) compiles and simulates
 * exactly those n.  scb_pop_fetch still returns all P sums: the other slots keep the value of the run that last
 * scored them.  Before the first run of an upload the whole population is scored whatever was replaced.  The arrays
 * may be reused when the call returns. */
int scb_pop_update(scb_ctx *ctx, int n, const int32_t *slots, const int32_t *individuals,
                   const int32_t *andNodes, const int32_t *andNodeInvertList);
int scb_pop_run(scb_ctx *ctx, int mode);
/* Work-list order.  scb_pop_costs returns, per individual of the last finished run, its cost in state-row units
 * (rows its tiles moved plus a fixed share per step; `costs` = the sum over its tiles -- what a shard of individuals
 * will take, `tileMax`, optional = its costliest tile -- what decides how late it may start); scb_pop_set_order
 * hands the next runs a permutation of the individuals.  With the individuals holding the costliest tiles first the
 * last tiles of a launch are short ones (most individuals of a (mu + lambda) generation were scored before, and
 * offspring behave much like their parents: the last cost predicts the next).  order = NULL: back to index order. */
int scb_pop_costs(scb_ctx *ctx, uint32_t *costs, uint32_t *tileMax);
int scb_pop_set_order(scb_ctx *ctx, const int32_t *order, int P);

/* ---- the generation's fitness all-gather over peer memory (one process per GPU of a box) -----------------------
 * Replaces the NCCL all-gather of the sharded population with one small kernel at the end of the
 * generation's graph: every rank stores its shard of `width` int64 values into every rank's gather buffer over
 * NVLink and raises a flag there.  scb_gather_init allocates this rank's buffers and returns two CUDA IPC handles
 * (128 bytes); the caller exchanges them between the ranks (any transport; bench.py uses torch.distributed) and
 * passes all of them, in rank order, to scb_gather_connect.  From then on every scb_pop_run(ctx, SCB_MODE_SUM) ends
 * with the exchange; scb_gather_result waits and copies the gathered vector int64[world * width] (rank r's shard at
 * r * width).  All ranks must run the same number of generations.  The exchange carries what the compiled path
 * computed: scb_gather_result fails on a rank whose population needs the general simulator (rules over more than three
 * inputs) or holds malformed rules; such populations go through scb_pop_fetch + any host-side collective. */
int scb_gather_init(scb_ctx *ctx, int world, int rank, int width, unsigned char *handles128);
int scb_gather_connect(scb_ctx *ctx, const unsigned char *allHandles);
int scb_gather_result(scb_ctx *ctx, int64_t *gathered);
int scb_gather_close(scb_ctx *ctx);
int scb_pop_fetch(scb_ctx *ctx, int mode, void *out, scb_stats *stats);
int scb_pop_result_dev(scb_ctx *ctx, int mode, void **devPtr, int64_t *bytes);

/* stream / timing plumbing for callers that time or chain work on the context's stream.
 * scb_ctx_flush_l2 overwrites a buffer larger than L2 on the stream (benchmark hygiene);
 * scb_host_alloc / scb_host_free hand out pinned host memory for asynchronous uploads. */
int scb_ctx_flush_l2(scb_ctx *ctx);
int scb_host_alloc(void **ptr, int64_t bytes);
int scb_host_free(void *ptr);
int scb_ctx_stream(scb_ctx *ctx, void **cudaStream);
int scb_ctx_sync(scb_ctx *ctx);
int scb_ctx_event_record(scb_ctx *ctx, int slot);            /* slot 0..7                     */
int scb_ctx_event_elapsed_ms(scb_ctx *ctx, int slot0, int slot1, float *ms);

/* ---- local search (ruleMaker.__checkNodePossibilities, ruleMaker.py:1235-1266): all
 * 2^k - 1 rule bit-strings of `node` on top of the base individual, one launch;
 * nodeErrors[v-1] = errors[node] of variant v, whose bits (least significant first, as
 * __bitList produces them, ruleMaker.py:796-802) replace the node's k bits exactly like
 * ruleMaker.py:1247-1253. */
int scb_eval_local_search(scb_ctx *ctx, const int32_t *individual, int indLen,
                          const int32_t *andLenList, const int32_t *individualParse,
                          const int32_t *andNodes, const int32_t *andNodeInvertList,
                          const int32_t *knockouts, const int32_t *knockins, int node,
                          int32_t *nodeErrors, int maxVariants, int *nVariants);

/* The whole local search of one individual in one launch: the 2^k - 1 variants of EVERY node (the loop over
 * nodes at singleCell.py:478-510).  nodeErrors[offsets[i] + v - 1] = errors[i] of variant v of node i;
 * offsets has nodeNum + 1 entries (prefix sums of 2^k_i - 1, 0 for nodes without rule bits). */
int scb_eval_local_search_all(scb_ctx *ctx, const int32_t *individual, int indLen,
                              const int32_t *andLenList, const int32_t *individualParse,
                              const int32_t *andNodes, const int32_t *andNodeInvertList,
                              const int32_t *knockouts, const int32_t *knockins,
                              int32_t *nodeErrors, int maxVariants, int32_t *offsets);

/* ---- attractor companion (singleCell.assignAttractors, singleCell.py:1211-1238):
 * dist[c][a] = sum_i |attractors[a][i] - cell c's observed state|, decider = first minimum
 * (pandas idxmin), deciderDistance = the minimum.  dist may be NULL. */
int scb_hamming_argmin(scb_ctx *ctx, const int32_t *attractors, int nAttr, int32_t *dist,
                       int32_t *decider, int32_t *deciderDistance);

/* ---- trajectories for attractor extraction.
 * semantics 0: C `cluster` (simulator.c:487-552), simSteps 100; 1: Python `__cluster`
 * (singleCell.py:990-1059), simSteps 10.  For every cell: res[c][2] = attractor window as the
 * respective reference returns it, and the window's states bit-packed node-major:
 * states[c][s][ceil(N/32)] for s < maxStates (rows res[0]..res[1] inclusive for Python
 * semantics, [res[0], res[1]) for C).  nStates[c] = number of rows written. */
int scb_attractors(scb_ctx *ctx, const int32_t *individual, int indLen,
                   const int32_t *andLenList, const int32_t *individualParse,
                   const int32_t *andNodes, const int32_t *andNodeInvertList,
                   const int32_t *knockouts, const int32_t *knockins, int semantics,
                   int maxStates, int32_t *res, int32_t *nStates, uint32_t *states);

/* The distinct attractor states of all cells, in first-seen order (cell-major, then state order), deduplicated on
 * the device: what assignAttractors collects in a Python set (singleCell.py:1189-1207).  unique[min(*nUnique, cap)]
 * [ceil(N/32)] packed like `states` above; *nUnique is the true count (call again with a larger cap if it exceeds cap). */
int scb_attractor_states(scb_ctx *ctx, const int32_t *individual, int indLen,
                         const int32_t *andLenList, const int32_t *individualParse,
                         const int32_t *andNodes, const int32_t *andNodeInvertList,
                         const int32_t *knockouts, const int32_t *knockins, int semantics,
                         int maxStates, uint32_t *unique, int cap, int *nUnique);
/* scb_hamming_argmin on attractors already packed that way ([nAttr][ceil(N/32)], bits beyond N zero) */
int scb_hamming_argmin_packed(scb_ctx *ctx, const uint32_t *attractors, int nAttr, int32_t *dist,
                              int32_t *decider, int32_t *deciderDistance);

/* ---- importance score of one knock-out / knock-in pair (ruleMaker.__calcImportance calls the C function
 * importanceScore 3 x N times per network, ruleMaker.py:1268-1412): same arithmetic as simulator.c:555-755
 * on the context's packed matrix.  Fails with SCB_ERR_INVALID (score = NaN) when a cell has no attractor
 * window, where the reference divides by zero. */
int scb_importance_score(scb_ctx *ctx, const int32_t *individual, int indLen,
                         const int32_t *andLenList, const int32_t *individualParse,
                         const int32_t *andNodes, const int32_t *andNodeInvertList,
                         const int32_t *knockouts, const int32_t *knockins, double *score);

/* The same for nItems knock-out / knock-in pairs in ONE simulator launch: knockouts and knockins are
 * int32[nItems][N], scores double[nItems].  This is the loop over nodes of ruleMaker.__calcImportance
 * (ruleMaker.py:1285-1300: one importanceScore call per node with KOlist = KIlist = [node]); every score is
 * bit-identical to the single call.  Items whose cells lack an attractor window get NaN and the call returns
 * SCB_ERR_INVALID (the other scores are valid). */
int scb_importance_scores(scb_ctx *ctx, const int32_t *individual, int indLen,
                          const int32_t *andLenList, const int32_t *individualParse,
                          const int32_t *andNodes, const int32_t *andNodeInvertList,
                          int nItems, const int32_t *knockouts, const int32_t *knockins, double *scores);

/* ---- diagnostics / tests: the compiled programs of individual p of the last scb_pop_run.  descs / descsK are
 * uint32[N][4] (three input nodes, node | truth table << 24, sorted by arity), counts / countsK the number of nodes of
 * arity 3, 2, 1, 0.  The full program, and reduced program k (0 .. 4: constants propagated to depth 2, 5, 8, 12 and to
 * the fixed point) with fromK = the first step it may compute (0: not emitted).  Any pointer may be NULL. */
int scb_debug_programs(scb_ctx *ctx, int p, int k, uint32_t *descs, int32_t *counts,
                       uint32_t *descsK, int32_t *countsK, int32_t *fromK);

/* ---- measured on-box peaks for the roofline report (bench.py): kind 0 = 32-bit LOP3 ops/s, 1 = POPC ops/s,
 * 2 = bytes/s of conflict-free 128-bit shared-memory loads, all SMs busy.  Not part of the reference's interface. */
int scb_microbench(scb_ctx *ctx, int kind, double *perSecond);

/* ---- the reference's exported symbols (drop-in ./simulator.so) --------------------------
 * Argument lists are exactly simulator.c:350, :555 and :487.  ctypes passes every argument as
 * c_void_p (ruleMaker.py:1117-1138), so integers arrive in 64-bit registers: they are
 * declared intptr_t here and truncated to int.  binMat's row stride and simData's row stride
 * are the compile-time CELL / NODE of the reference (15000 / 20000) unless changed with
 * scb_set_legacy_geometry or the environment variables SCB_LEGACY_CELL / SCB_LEGACY_NODE.
 * These never throw or abort: on error they print to stderr and leave `errors` untouched. */
void scb_set_legacy_geometry(int nodeStride, int cellStride);
void scSyncBool(void *simData, const int32_t *individual, intptr_t indLen, intptr_t nodeNum,
                const int32_t *andLenList, const int32_t *individualParse,
                const int32_t *andNodes, const int32_t *andNodeInvertList, intptr_t simSteps,
                const int32_t *knockouts, const int32_t *knockins, intptr_t lenSamples,
                const int32_t *binMat, const int32_t *nodePositions, int32_t *errors,
                intptr_t localSearch, intptr_t importanceScores);
void importanceScore(void *simData, const int32_t *individual, intptr_t indLen, intptr_t nodeNum,
                     const int32_t *andLenList, const int32_t *individualParse,
                     const int32_t *andNodes, const int32_t *andNodeInvertList,
                     intptr_t simSteps, const int32_t *knockouts, const int32_t *knockins,
                     intptr_t lenSamples, const int32_t *binMat, const int32_t *nodePositions,
                     double *importanceScores);
void cluster(void *simData, int32_t *resSubmit, intptr_t sampleIndex, const int32_t *individual,
             intptr_t indLen, intptr_t nodeNum, const int32_t *andLenList,
             const int32_t *individualParse, const int32_t *andNodes,
             const int32_t *andNodeInvertList, intptr_t simSteps, const int32_t *knockouts,
             const int32_t *knockins, const int32_t *binMat, const int32_t *nodePositions);

#ifdef __cplusplus
}
#endif
#endif
