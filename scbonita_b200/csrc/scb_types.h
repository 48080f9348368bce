/*
 * scb_types.h -- structures shared by the host API and the kernels.
 *
 * Compiled two ways: by nvcc for sm_100a (the product), and by g++ with -DSCB_EMU against the
 * host-thread CUDA emulator in tests/emu/ (test infrastructure only: it lets the CPU test
 * suite run the *same kernel source* when no GPU is present; it is never built into or
 * loaded by the product library).
 */
#ifndef SCB_TYPES_H
#define SCB_TYPES_H
#include <stdint.h>

#ifdef SCB_EMU
#include "emu_cuda.h"
#else
#include <cuda_runtime.h>
#endif

#if defined(__CUDACC__) || defined(SCB_EMU)
#define SCB_HD __host__ __device__
#define SCB_DEV __device__ __forceinline__
#else
#define SCB_HD
#define SCB_DEV inline
#endif

#define SCB_MAX_TERMS 7   /* andNodes[NODE][7][3], simulator.c:350 */
#define SCB_MAX_IN 3
#define SCB_STEPS 100     /* simulator.c:6 (#define STEP); states S_0 .. S_99 */
#define SCB_LAST (SCB_STEPS - 1)

/* per-node compile flags */
#define SCB_NF_UNSUPPORTED 1u
#define SCB_NF_EMPTY       2u
#define SCB_NF_BADIDX      4u
#define SCB_NF_TERMS       8u

/* step flags */
#define SCB_SF_CMP2   1u   /* compare S_t with S_{t-2} (the row being overwritten)        */
#define SCB_SF_ZCMP   2u   /* compare S_t with the snapshot in global scratch             */
#define SCB_SF_ZSNAP  4u   /* write S_t to the snapshot                                    */
#define SCB_SF_TCMP   8u   /* compare S_t with the target state held in shared memory     */
#define SCB_SF_CONST 16u   /* also write the constant rows (steps 1 and 2 only)            */

/* internal output shape of K2 / K3 next to the public SCB_MODE_* (0..2): the counts importanceScore needs,
 * nodewise[p][i] = number of fixed-point cells (S_98 == S_99) with node i set, fitness[p] = cells without an
 * attractor window (simulator.c:641-648) */
#define SCB_MODE_IMPORTANCE 3

/* constant propagation depth cap of the steady program (K0): must stay below the first snapshot step */
#ifndef SCB_STEADY_MAX_DEPTH
#define SCB_STEADY_MAX_DEPTH 24
#endif
/* reduced programs per individual: at these propagation depths, plus the fixed point (or the cap) */
#define SCB_NPROG 5
#define SCB_STAGE_DEPTHS {2, 5, 8, 12}

/* kernel behaviour flags (ScbSimArgs.flags) */
#define SCB_KF_EARLY  1u
#define SCB_KF_SNAP   2u

struct ScbDiag {                      /* device-side counters, zeroed per evaluation */
    unsigned long long unsupported_nodes;
    unsigned long long empty_rule_nodes;
    unsigned long long bad_index_nodes;
    unsigned long long term_overflow_nodes;
    unsigned long long nonbinary_cells;
    unsigned long long not_found_cells;
    unsigned long long undefined_leading;
    unsigned long long resolver_chunks;
    unsigned long long steps_executed;
    unsigned long long tiles;
    unsigned long long rows_moved;      /* state rows K2 loaded + stored in its steps (x 128*K bytes = shared-memory traffic) */
    unsigned long long cyc_prologue, cyc_steps, cyc_epilogue;   /* SM clocks of thread 0 of every K2 CTA, per phase */
#ifdef SCB_TRACE
    /* measurement build only (-DSCB_TRACE): SM clock of every warp of CTA 0 at the start and the end of every step
     * of its first tile, [step][warp][2]; step 0 holds the step's flags in [0][0][0] .. */
    unsigned int trace[100 * 24 * 4];
    unsigned int traceFlags[100];
    unsigned int evCount;
    unsigned int evLog[16384 * 4];   /* {kind (0 tile, 1 resolver item) | CTA << 8, globaltimer ns at start, at end, steps | individual << 8} */
    unsigned int exitHist[104];   /* tiles by the step their loop ended at; [100] E2 exits, [101] E3, [102] all proven, [103] queued for K3 */
#endif
};

/* One compiled node: d.x, d.y, d.z = input node indices, d.w = node index | lut8 << 24.
 * Sorted per individual by arity: [arity 3 | arity 2 | arity 1 | constants], counts in an int4
 * (x = n3, y = n2, z = n1, w = n0).  lut8 follows lop3's convention: bit ((a<<2)|(b<<1)|c). */

/* KG, the general simulator (scb_k_generic): arguments */
struct ScbGenericArgs {
    const uint32_t *X;            /* packed S_0, [N][Wpad] */
    int N, C, Wpad, P, indLen, tablesPerIndividual, compact, mode;
    const int32_t *individuals, *andLen, *parse, *an, *ai, *ko, *ki;    /* as uploaded (ScbCompileArgs) */
    uint32_t *scratch;            /* [CTAs][3][N][32] state buffers in global memory */
    unsigned long long *fitness;
    int32_t *nodewise, *cellwise;
    int32_t *chunkLast, *chunkLastErr;
    uint32_t *chunkLastBits;
    int chunksPerInd;
    unsigned int *fillCount;
    uint4 *fillItems;
    unsigned int *cursor;         /* work counter over the P x chunksPerInd items */
    ScbDiag *diag;
};

struct ScbCompileArgs {
    int P, N, indLen, tablesPerIndividual, Npad;
    int sem;                      /* SCB_SEM_* (scb_compile.h) */
    const int32_t *individuals;   /* [P][indLen] */
    const int32_t *andLen;        /* [N] */
    const int32_t *parse;         /* [N] */
    const int32_t *an, *ai;       /* [P or 1][N][7][3]; compact != 0: [P or 1][N][3] upstream nodes / inversion flags */
    int compact;                  /* != 0: the term tables are expanded on the fly from the upstream triples */
    const int32_t *ko, *ki;       /* [N] */
    int koStride;                 /* 0: one shared `ko`; N: program p reads ko + p * N (batched importance scores) */
    uint4 *descs;                 /* [P][Npad] */
    int4 *counts;                 /* [P] */
    uint4 *progs;                 /* [P][SCB_NPROG][N + 1] reduced programs (constants propagated to a fixed depth; the
                                     last one to the fixed point); entry N of each holds its arity counts */
    unsigned char *progFrom;      /* [P][8] first step each reduced program may compute (0: not emitted) */
    uint4 *imgs;                  /* [P][SCB_NPROG + 1][imgStride] program images for the simulator (scb_layout_program):
                                     image 0 = the full program, 1 + k = reduced program k; null: not wanted */
    int imgStride;                /* entries per image */
    int imgWarps;                 /* warps of the simulator CTA the images are cut for */
    unsigned int imgRowBytes;     /* bytes of a state row in the simulator's shared memory (128 * K) */
    ScbDiag *diag;
    const int32_t *slots;         /* null: CTA b compiles individual b; else individual slots[b] (resident population,
                                     only the replaced individuals are compiled) */
    unsigned long long *fitness;  /* [P] result sums, cleared here for the compiled individuals (null: not wanted) */
    unsigned int *cost;           /* [2][P] cost counters of the simulator, cleared likewise */
    unsigned int *pred;           /* [P] predicted cost of the individual = state rows per step of its fixed-point program
                                     (Spearman 0.7-0.8 with the measured cost on C3); orders a population never scored before */
};

struct ScbPackArgs {
    const void *src;              /* [N][rowStride] elements, row i = node i */
    int elemBytes;                /* 4 or 1 */
    long long rowStride;          /* elements */
    int N, C, Wpad;
    uint32_t *X;                  /* [N][Wpad] */
    ScbDiag *diag;
};

struct ScbSimArgs {
    const uint32_t *X;            /* packed S_0: [N][Wpad] node-major, 32 cells per word */
    int N, C, Wpad, tiles, P, Npad, mode;
    int nEval;                    /* individuals this launch simulates: P, or the replaced ones of a resident population (order[] lists them) */
    int wordBase;                 /* first word of this launch's tile range (tail tiles use a narrower K) */
    const uint4 *descs;
    const int4 *counts;
    const uint4 *progs;           /* reduced programs of K0, [P][SCB_NPROG][N + 1] */
    const uint4 *imgs;            /* program images of K0 for this launch's geometry, [P][SCB_NPROG + 1][imgStride] */
    int imgStride;
    const unsigned char *progFrom;/* [P][8] */
    int useSteady;
    int chkFromProg;              /* >= 0: no period-<=2 check before reduced program `chkFromProg` may run; -1: no such delay */
    unsigned long long chkLo, chkHi;    /* bit t set: S_t is compared with S_{t-2} (every chkEvery-th step, and step 99) */
    unsigned long long zcmpLo, zcmpHi;  /* bit t set: S_t is compared with the latest snapshot (from snapMinGap steps after it on) */
    int firstSnap;                /* first step that takes a snapshot (1000: none): the steady program must be in use by then */
    unsigned long long *fitness;  /* [P] */
    int32_t *nodewise;            /* [P][N] or null */
    int32_t *cellwise;            /* [P][C] or null */
    unsigned int *workCounter;    /* persistent-CTA work queue head */
    unsigned int *resolveCount;   /* number of chunks queued for the exact resolver */
    uint2 *resolveItems;          /* (individual, chunk) */
    unsigned int resolveCap;
    int resolveGroup;             /* adjacent chunks per resolver item (1 or 2; <= K of the launch) */
    uint32_t *resolveT;           /* S_99 of queued groups: [resolveTCap][N][32*resolveGroup] words (saves K3 a pass) */
    unsigned int resolveTCap;
    int32_t *chunkLast;           /* [P][chunksPerInd]: last found valid cell of each chunk, -1 none */
    int32_t *chunkLastErr;        /* [P][chunksPerInd]: number of nodes that cell gets wrong (its column of e = S_98 xor S_0) */
    uint32_t *chunkLastBits;      /* [P][chunksPerInd][ceil(N/32)]: that column, bit = node; written in the per-node mode only */
    int chunksPerInd;
    unsigned int *fillCount;      /* chunks whose leading not-found cells copy an earlier chunk's cell */
    uint4 *fillItems;             /* (individual, chunk, pending cells, first pending-free cell) */
    uint32_t *zscratch;           /* snapshot scratch, zstride words per CTA */
    unsigned long long zstride;
    unsigned int flags;           /* SCB_KF_* */
    unsigned long long snapLo, snapHi;  /* bit t set: take a snapshot when computing S_t (0 when snapshots are off) */
    int periodSearch;             /* snapshot-compare steps kept after all cells resolved */
    int snapMinGap;               /* first snapshot compare this many steps after the snapshot */
    int fuseResolve;              /* != 0: CTAs that run out of tiles resolve queued groups in the same launch */
    unsigned int *tilesDone;      /* tiles completed (all their resolver items are published) */
    unsigned int *readyFlags;     /* [resolveCap] item i is complete in memory */
    unsigned int *resolveCursor;  /* ticket counter of the in-kernel resolver */
    unsigned int *retCount;       /* tiles handed back: a CTA that takes a resolver item gives up the tile ticket it holds */
    unsigned int *retCursor;      /*   ... taken again */
    unsigned int *retItems;       /* [resolveCap] handed-back tile item + 1, 0 = not written yet */
    unsigned int smemBytes;       /* dynamic shared memory of the launch (the item slot sits in its last 64 bytes) */
    unsigned int *cost;           /* [P] or null: state rows each individual's tiles moved (the cost predictor of the next generation) */
    const int32_t *order;         /* [P] or null: individuals in the order their tiles are handed out (costliest first) */
    int useTma;                   /* != 0: the S_0 tile is loaded with cp.async.bulk (TMA) row copies */
    int tmemCols;                 /* > 0: the snapshot lives in tensor memory, this many columns per CTA */
    int tmemSlotsPerWarp;         /* node slots reserved per warp for the resolver's target state (program order) */
    ScbDiag *diag;
};

/* K4: trajectories of one individual for every cell, stored tile-local: traj[tile][step][N][32] words */
struct ScbTrajArgs {
    const uint32_t *X;            /* [N][Wpad] */
    int N, C, Wpad, tiles, steps;
    const uint4 *descs;           /* one compiled program */
    const int4 *counts;
    uint32_t *traj;
    const unsigned char *recOn;   /* [N] or null: != 0 -> the node's RECORDED rows of steps >= 1 are overridden (the dynamics are not) */
    const uint32_t *recBits;      /* [N][4]: bit t = recorded value at step t of an overridden node (simulator.c:518) */
};

struct ScbWindowArgs {
    const uint32_t *traj;
    int N, C, tiles, steps, sem, maxStates, NW;
    int32_t *res;                 /* [C][2] */
    int32_t *nStates;             /* [C] */
    uint32_t *states;             /* [C][maxStates][NW] */
};

/* fitness exchange between the GPUs of a box over peer memory (NVLink): every rank writes its shard into every
 * rank's gather buffer and then raises its flag there; see scb_k_gather */
#define SCB_MAX_PEERS 16
struct ScbGatherArgs {
    const unsigned long long *fitness;   /* this rank's shard, [count] */
    int count, width, rank, world;       /* every rank owns `width` slots of the gathered vector */
    unsigned long long *peerBuf[SCB_MAX_PEERS];   /* gather buffers [2][world * width] of every rank (own one included) */
    unsigned int *peerFlag[SCB_MAX_PEERS];        /* flag arrays [world] of every rank */
    unsigned int *generation;            /* this rank's generation counter (device) */
    unsigned int *timeout;               /* set when a peer did not show up */
};

/* unique attractor states on the device (scb_k_dedupe_*): rows = [C][maxStates][NW] packed states, valid where
 * s < min(nStates[c], maxStates) */
struct ScbDedupeArgs {
    const uint32_t *states;
    const int32_t *nStates;
    int C, maxStates, NW;
    unsigned int *table;          /* open addressing, 0xffffffff = empty, else the lowest row index seen with that state */
    unsigned int mask;            /* table size - 1 (a power of two, >= 2 x rows) */
    unsigned char *flags;         /* [rows] 1 = first occurrence of its state */
    unsigned int *count;          /* number of distinct states */
    uint32_t *out;                /* [cap][NW] distinct states in first-seen order */
    int cap;
};

struct ScbHammingArgs {
    const uint32_t *X;            /* [N][Wpad] */
    int N, C, Wpad, A;
    const uint32_t *attr;         /* [A][NWP]: bit i%32 of word i/32 = node i of the attractor; NWP = ceil(N/128) * 4 words, zero padded */
    int32_t *dist;                /* [C][A] or null */
    int32_t *decider;             /* [C] */
    int32_t *deciderDistance;     /* [C] */
};

#endif
