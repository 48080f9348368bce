/*
 * scb_oracle.h -- CPU restatement of scBONITA's rule-inference fitness hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (scbonita_b200/) never links or imports this.
 *
 * Parity pin: the reference ships no bit-exact known-answer vectors for the simulator
 * (SURVEY.md section 8c), so these restatements are pinned by differential tests against the
 * UNMODIFIED reference C compiled into oracle/_ref/ (tests/test_oracle_vs_ref.py) and by the
 * golden vectors generated from that build (tests/golden/).  The Hamming/argmin stage is
 * additionally pinned by the reference's own packaged attractorDistance.csv.
 *
 * All citations are file:line under /root/reference/src/scBONITA/.
 */
#ifndef SCB_ORACLE_H
#define SCB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCBO_STEP 100      /* simulator.c:6  (#define STEP) */
#define SCBO_MAX_TERMS 7   /* andNodes[NODE][7][3], simulator.c:350 */
#define SCBO_MAX_IN 3

/* diagnostics filled by the restatements (all optional: pass NULL) */
typedef struct {
    int64_t not_found_cells;     /* cells whose row 99 matched no earlier row (Q7)            */
    int64_t undefined_leading;   /* not-found cells before the first found one: the reference */
                                 /* reads uninitialised stack there; we define the value as 0 */
    int64_t empty_rule_evals;    /* updateBool with no selected term (Q9): defined as 0 here  */
} scbo_diag;

/* simulator.c:350-485 (scSyncBool), literal restatement with runtime strides.
 *   binMat         int32, row stride binMatStride (reference: compile-time CELL)
 *   errors         localSearch==0: errors[c]  = ... for c < lenSamples (assignment, :481)
 *                  localSearch!=0: errors[i] += ... for i < nodeNum    (accumulate, :477-479)
 *   simData        optional caller scratch [100][simStride] (reference: simData[STEP][NODE]);
 *                  NULL -> an internal zeroed scratch with stride max(nodePositions)+1.
 *   found          optional int8[lenSamples]: 1 if getAttract located a repeat for the cell.
 * returns 0, or -1 on a contract violation (term count > 7, negative sizes). */
int scbo_sync_bool(const int32_t *individual, int indLen, int nodeNum,
                   const int32_t *andLenList, const int32_t *individualParse,
                   const int32_t *andNodes, const int32_t *andNodeInvertList,
                   int simSteps, const int32_t *knockouts, const int32_t *knockins,
                   int lenSamples, const int32_t *binMat, int64_t binMatStride,
                   const int32_t *nodePositions, int32_t *errors, int localSearch,
                   int32_t *simData, int64_t simStride, int8_t *found, scbo_diag *diag);

/* simulator.c:487-552 (cluster) for one cell: trajectory in node-index order
 * traj[step*nodeNum + i], res = {first, second} from getAttractCycle2 (:306-331). */
int scbo_cluster(int sampleIndex, const int32_t *individual, int indLen, int nodeNum,
                 const int32_t *andLenList, const int32_t *individualParse,
                 const int32_t *andNodes, const int32_t *andNodeInvertList,
                 int simSteps, const int32_t *knockouts, const int32_t *knockins,
                 const int32_t *binMat, int64_t binMatStride, const int32_t *nodePositions,
                 int32_t *traj, int32_t res[2], scbo_diag *diag);

/* singleCell.py:990-1059 (__cluster) + :961-982 (__getAttractCycle2), Python semantics,
 * for all cells: vals[c][step][i] (int32, what the Python code leaves in `vals`),
 * res[c][2] = [j, i] or [-1,-1].  simSteps must be 10 (STEP hard-coded at :963). */
int scbo_py_cluster(const int32_t *individual, int indLen, int nodeNum,
                    const int32_t *andLenList, const int32_t *individualParse,
                    const int32_t *andNodes, const int32_t *andNodeInvertList,
                    int simSteps, const int32_t *knockouts, const int32_t *knockins,
                    int lenSamples, const int32_t *binMat, int64_t binMatStride,
                    const int32_t *nodePositions, int32_t *vals, int32_t *res, scbo_diag *diag);

/* singleCell.py:1211-1238: D[c][a] = sum_i |attr[a][i] - binMat[pos[i]][c]|, first-minimum
 * column (pandas idxmin) and the minimum. dist may be NULL. */
int scbo_hamming_argmin(const int32_t *attractors, int nAttr, int nodeNum, int lenSamples,
                        const int32_t *binMat, int64_t binMatStride,
                        const int32_t *nodePositions, int32_t *dist, int32_t *decider,
                        int32_t *deciderDistance);

/* simulator.c:555-755 (importanceScore), literal incl. its quirks.
 * returns -2 instead of dividing by zero when a cell's attractor is not found. */
/* `cluster` restated literally on the caller's scratch, incl. the column slip of its knock-out branch
 * (simulator.c:518); see scb_oracle.c */
int scbo_cluster_literal(int sampleIndex, const int32_t *individual, int indLen, int nodeNum,
                         const int32_t *andLenList, const int32_t *individualParse,
                         const int32_t *andNodes, const int32_t *andNodeInvertList,
                         int simSteps, const int32_t *knockouts, const int32_t *knockins,
                         const int32_t *binMat, int64_t binMatStride, const int32_t *nodePositions,
                         int32_t *simData, int64_t simStride, int32_t res[2], scbo_diag *diag);

int scbo_importance_score(const int32_t *individual, int indLen, int nodeNum,
                          const int32_t *andLenList, const int32_t *individualParse,
                          const int32_t *andNodes, const int32_t *andNodeInvertList,
                          int simSteps, const int32_t *knockouts, const int32_t *knockins,
                          int lenSamples, const int32_t *binMat, int64_t binMatStride,
                          const int32_t *nodePositions, double *score);

/* ---- bit-packed restatement (scb_oracle_packed.c): same semantics, 64 cells per word,
 * OpenMP over individuals; fast enough for the BASELINE-size configs. -------------------
 *   andNodes/andInv: [P][N][7][3] when tablesPerIndividual != 0, else [N][7][3] shared
 *   individuals:     [P][indLen]
 *   fitness:         int64[P]      (sum of the cellwise vector, = what Python's sum(errors) returns)
 *   nodewise:        optional int32[P][N]
 *   cellwise:        optional int32[P][C]
 *   notFound:        optional int64[P] count of not-found cells
 *   muLambda:        optional int32[P][C][2]: transient and period of each cell's trajectory
 *                    (-1,-1 when no repeat within 100 states) -- analysis aid, not reference output */
int scbo_packed_population(int P, const int32_t *individuals, int indLen, int nodeNum,
                           const int32_t *andLenList, const int32_t *individualParse,
                           const int32_t *andNodes, const int32_t *andNodeInvertList,
                           int tablesPerIndividual,
                           const int32_t *knockouts, const int32_t *knockins,
                           int lenSamples, const int32_t *binMat, int64_t binMatStride,
                           const int32_t *nodePositions,
                           int64_t *fitness, int32_t *nodewise, int32_t *cellwise,
                           int64_t *notFound, int32_t *muLambda, int nThreads);

#ifdef __cplusplus
}
#endif
#endif
