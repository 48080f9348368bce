/*
 * scb_oracle_packed.c -- bit-packed (64 cells per word) CPU restatement of scSyncBool for
 * whole populations.  TEST INFRASTRUCTURE ONLY (see scb_oracle.h).
 *
 * Same semantics as scbo_sync_bool for simSteps == 100, evaluated
 * term by term straight from the reference-format tables (no truth-table compilation, so it
 * is algorithmically independent of the CUDA path's rule compiler).  It is itself checked
 * against scbo_sync_bool and against oracle/_ref in tests/test_oracle_vs_ref.py, and exists
 * so that the BASELINE-size configurations can be compared in full.
 *
 * Reference lines restated: simulator.c:10-50 (rule), :369-422 (trajectory), :306-331 and
 * :426-482 (attractor window and error, reduced to: found <=> state 99 repeats an earlier
 * state; error taken from state 98; not-found cells carry the previous cell's error).
 */
#include "scb_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TBL(t, i, a, b) ((t)[((int64_t)(i) * SCBO_MAX_TERMS + (a)) * SCBO_MAX_IN + (b)])
typedef uint64_t w64;

static inline w64 lit(w64 x, int32_t flag)
{   /* int comparison old != flag on 0/1 values (Q3) */
    return flag == 0 ? x : (flag == 1 ? ~x : ~(w64)0);
}

static void step_words(int n, int indLen, const int32_t *ind, const int32_t *andLen,
                       const int32_t *parse, const int32_t *an, const int32_t *ai,
                       const int32_t *ko, const int32_t *ki, const w64 *old, w64 *nw)
{
    for (int i = 0; i < n; ++i) {
        w64 v;
        if (ko[i] == 1) v = 0;
        else if (ki[i] == 1) v = ~(w64)0;
        else if (andLen[i] == 1) v = lit(old[TBL(an, i, 0, 0)], TBL(ai, i, 0, 0));
        else if (andLen[i] == 0) v = old[TBL(an, i, 0, 0)];
        else {
            int start = parse[i], end = (i == n - 1) ? indLen : parse[i + 1];
            v = 0;
            for (int a = 0; a < end - start; ++a) {
                if (!ind[start + a]) continue;
                w64 t = lit(old[TBL(an, i, a, 0)], TBL(ai, i, a, 0));
                for (int b = 1; b < SCBO_MAX_IN; ++b)
                    if (TBL(an, i, a, b) > -1) t &= lit(old[TBL(an, i, a, b)], TBL(ai, i, a, b));
                v |= t;
            }
        }
        nw[i] = v;
    }
}

int scbo_packed_population(int P, const int32_t *individuals, int indLen, int nodeNum,
                           const int32_t *andLenList, const int32_t *individualParse,
                           const int32_t *andNodes, const int32_t *andNodeInvertList,
                           int tablesPerIndividual,
                           const int32_t *knockouts, const int32_t *knockins,
                           int lenSamples, const int32_t *binMat, int64_t binMatStride,
                           const int32_t *nodePositions,
                           int64_t *fitness, int32_t *nodewise, int32_t *cellwise,
                           int64_t *notFound, int32_t *muLambda, int nThreads)
{
    const int n = nodeNum, C = lenSamples;
    if (n <= 0 || C < 0 || P < 0) return -1;
    for (int i = 0; i < n; ++i) {
        int end = (i == n - 1) ? indLen : individualParse[i + 1];
        if (andLenList[i] >= 2 && end - individualParse[i] > SCBO_MAX_TERMS) return -1;
    }
    const int W = (C + 63) / 64;
#ifdef _OPENMP
    if (nThreads > 0) omp_set_num_threads(nThreads);
#endif
    /* pack S0: node-major words */
    w64 *s0 = (w64 *)calloc((size_t)n * (W ? W : 1), sizeof(w64));
    for (int i = 0; i < n; ++i) {
        const int32_t *row = binMat + (int64_t)nodePositions[i] * binMatStride;
        for (int c = 0; c < C; ++c)
            if (row[c]) s0[(size_t)i * W + (c >> 6)] |= (w64)1 << (c & 63);
    }
    w64 *E = (w64 *)malloc(sizeof(w64) * (size_t)n * (W ? W : 1));   /* error bits [n][W]   */
    w64 *F = (w64 *)malloc(sizeof(w64) * (size_t)(W ? W : 1));       /* found mask [W]      */
    int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(C ? C : 1));
    const int64_t tstride = (int64_t)SCBO_MAX_TERMS * SCBO_MAX_IN * n;

    for (int p = 0; p < P; ++p) {
        const int32_t *ind = individuals + (int64_t)p * indLen;
        const int32_t *an = andNodes + (tablesPerIndividual ? p * tstride : 0);
        const int32_t *ai = andNodeInvertList + (tablesPerIndividual ? p * tstride : 0);
#pragma omp parallel
        {
            w64 *traj = (w64 *)malloc(sizeof(w64) * (size_t)SCBO_STEP * n);
#pragma omp for schedule(dynamic, 4)
            for (int w = 0; w < W; ++w) {
                for (int i = 0; i < n; ++i) traj[i] = s0[(size_t)i * W + w];
                for (int s = 1; s < SCBO_STEP; ++s)
                    step_words(n, indLen, ind, andLenList, individualParse, an, ai,
                               knockouts, knockins, traj + (size_t)(s - 1) * n, traj + (size_t)s * n);
                const w64 *last = traj + (size_t)(SCBO_STEP - 1) * n;
                w64 valid = (w == W - 1 && (C & 63)) ? (((w64)1 << (C & 63)) - 1) : ~(w64)0;
                w64 found = 0;
                w64 eqAt[SCBO_STEP];
                for (int j = 0; j < SCBO_STEP - 1; ++j) {
                    const w64 *sj = traj + (size_t)j * n;
                    w64 diff = 0;
                    for (int i = 0; i < n; ++i) diff |= sj[i] ^ last[i];
                    eqAt[j] = ~diff;
                    found |= ~diff;
                }
                F[w] = found & valid;
                const w64 *s98 = traj + (size_t)(SCBO_STEP - 2) * n;
                for (int i = 0; i < n; ++i) E[(size_t)i * W + w] = (s98[i] ^ traj[i]) & valid;
                int lim = (w == W - 1 && (C & 63)) ? (C & 63) : 64;
                for (int b = 0; b < lim; ++b) {
                    int k = 0;
                    for (int i = 0; i < n; ++i) k += (int)((E[(size_t)i * W + w] >> b) & 1);
                    cnt[(size_t)w * 64 + b] = k;
                }
                if (muLambda) {
                    for (int b = 0; b < lim; ++b) {
                        int32_t *ml = muLambda + 2 * ((int64_t)p * C + (int64_t)w * 64 + b);
                        ml[0] = -1; ml[1] = -1;
                        if (!((found >> b) & 1)) continue;
                        int j = SCBO_STEP - 2;
                        while (!((eqAt[j] >> b) & 1)) --j;
                        int lam = SCBO_STEP - 1 - j, mu = 0;
                        for (;; ++mu) {          /* first t with S_t == S_{t+lam} for this cell */
                            const w64 *a = traj + (size_t)mu * n, *c2 = traj + (size_t)(mu + lam) * n;
                            w64 d = 0;
                            for (int i = 0; i < n; ++i) d |= a[i] ^ c2[i];
                            if (!((d >> b) & 1)) break;
                        }
                        ml[0] = mu; ml[1] = lam;
                    }
                }
            }
            free(traj);
        }
        /* reductions + the sequential fill-forward of not-found cells (Q7) */
        int64_t fit = 0, nf = 0;
        int32_t *nwz = nodewise ? nodewise + (int64_t)p * n : NULL;
        if (nwz) for (int i = 0; i < n; ++i) {
            int64_t s = 0;
            for (int w = 0; w < W; ++w) s += __builtin_popcountll(E[(size_t)i * W + w] & F[w]);
            nwz[i] = (int32_t)s;
        }
        int prev = -1;                               /* last found cell so far */
        for (int c = 0; c < C; ++c) {
            int isFound = (int)((F[c >> 6] >> (c & 63)) & 1);
            int val;
            if (isFound) { prev = c; val = cnt[c]; }
            else {
                ++nf;
                val = prev >= 0 ? cnt[prev] : 0;     /* leading not-found: defined as 0 */
                if (nwz && prev >= 0)
                    for (int i = 0; i < n; ++i)
                        nwz[i] += (int32_t)((E[(size_t)i * W + (prev >> 6)] >> (prev & 63)) & 1);
            }
            fit += val;
            if (cellwise) cellwise[(int64_t)p * C + c] = val;
        }
        if (fitness) fitness[p] = fit;
        if (notFound) notFound[p] = nf;
    }
    free(s0); free(E); free(F); free(cnt);
    return 0;
}
