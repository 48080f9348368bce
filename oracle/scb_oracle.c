/*
 * scb_oracle.c -- scalar CPU restatement of the reference simulator (see scb_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never imported, linked or executed by the product path.
 *
 * Written from the reference's behaviour, one function per reference function, with the
 * compile-time NODE/CELL geometry (simulator.c:6-8) replaced by runtime strides.  The places
 * where the reference reads uninitialised memory are given a defined value here and counted
 * in scbo_diag so that tests can exclude them from parity claims.
 */
#include "scb_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define TBL(t, i, a, b) ((t)[((int64_t)(i) * SCBO_MAX_TERMS + (a)) * SCBO_MAX_IN + (b)])

/* simulator.c:10-50 updateBool: OR over the selected AND-terms of one node.
 * A literal is the int comparison old[src] != flag (flag 0: identity, 1: NOT, else TRUE). */
static int rule_value(const int32_t *old, const int32_t *bits, int nTerms,
                      const int32_t *an, const int32_t *ai, scbo_diag *diag)
{
    int have = 0, acc = 0;
    for (int a = 0; a < nTerms; ++a) {
        if (!bits[a]) continue;                                   /* :26 */
        int v = old[an[a * 3]] != ai[a * 3];                      /* :30, input 0 always used */
        for (int b = 1; b < SCBO_MAX_IN; ++b)                     /* :32-35 */
            if (an[a * 3 + b] > -1) v = v && (old[an[a * 3 + b]] != ai[a * 3 + b]);
        acc = have ? (acc || v) : v;                              /* :43-48 */
        have = 1;
    }
    if (!have) {               /* :43 reads orset[0] uninitialised: defined as 0 */
        if (diag) diag->empty_rule_evals++;
        return 0;
    }
    return acc;
}

typedef struct {
    int nodeNum, indLen;
    const int32_t *individual, *andLen, *parse, *an, *ai, *ko, *ki;
} net_t;

/* one synchronous update, simulator.c:377-416 (scSyncBool) / :509-549 (cluster) */
static int sync_step(const net_t *m, const int32_t *old, int32_t *nw, scbo_diag *diag)
{
    for (int i = 0; i < m->nodeNum; ++i) {
        int v;
        if (m->ko[i] == 1) v = 0;                                            /* :383 */
        else if (m->ki[i] == 1) v = 1;                                       /* :388 */
        else if (m->andLen[i] == 1) v = old[TBL(m->an, i, 0, 0)] != TBL(m->ai, i, 0, 0); /* :395 */
        else if (m->andLen[i] == 0) v = old[TBL(m->an, i, 0, 0)];            /* :401, Q2 */
        else {
            int end = (i == m->nodeNum - 1) ? m->indLen : m->parse[i + 1];   /* :408-409 */
            int start = m->parse[i];
            if (end - start > SCBO_MAX_TERMS) return -1;                     /* would overrun [7] */
            v = rule_value(old, m->individual + start, end - start,
                           &TBL(m->an, i, 0, 0), &TBL(m->ai, i, 0, 0), diag);
        }
        nw[i] = v;
    }
    return 0;
}

/* simulator.c:306-331 getAttractCycle2: rows compared on the nodePositions columns only,
 * always over all STEP rows; first hit scanning i downwards, j = i-1 downwards. */
static void attract_window(const int32_t *rows, int64_t stride, const int32_t *pos, int n,
                           int32_t res[2])
{
    res[0] = 0; res[1] = 0;
    for (int i = SCBO_STEP - 1; i >= 0; --i)
        for (int j = i - 1; j >= 0; --j) {
            int same = 1;
            for (int k = 0; k < n; ++k)
                if (rows[i * stride + pos[k]] != rows[j * stride + pos[k]]) { same = 0; break; }
            if (same) { res[0] = j; res[1] = i; return; }
        }
}

static int64_t max_pos(const int32_t *pos, int n)
{
    int64_t m = 0;
    for (int i = 0; i < n; ++i) if (pos[i] > m) m = pos[i];
    return m;
}

int scbo_sync_bool(const int32_t *individual, int indLen, int nodeNum,
                   const int32_t *andLenList, const int32_t *individualParse,
                   const int32_t *andNodes, const int32_t *andNodeInvertList,
                   int simSteps, const int32_t *knockouts, const int32_t *knockins,
                   int lenSamples, const int32_t *binMat, int64_t binMatStride,
                   const int32_t *nodePositions, int32_t *errors, int localSearch,
                   int32_t *simData, int64_t simStride, int8_t *found, scbo_diag *diag)
{
    if (nodeNum <= 0 || lenSamples < 0 || simSteps > SCBO_STEP) return -1;
    net_t m = { nodeNum, indLen, individual, andLenList, individualParse,
                andNodes, andNodeInvertList, knockouts, knockins };
    int32_t *own = NULL;
    if (!simData) {
        simStride = max_pos(nodePositions, nodeNum) + 1;
        own = (int32_t *)calloc((size_t)SCBO_STEP * simStride, sizeof(int32_t));
        if (!own) return -1;
        simData = own;
    }
    int32_t *old = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    int32_t *nw = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    /* `error[]` and `totalError` live across cells in the reference (:361-362) and are
     * uninitialised before the first found cell; defined as 0 here (Q7). */
    int32_t *err = (int32_t *)calloc(nodeNum, sizeof(int32_t));
    int total = 0, seenFound = 0, rc = 0;
    if (diag) memset(diag, 0, sizeof(*diag));

    for (int c = 0; c < lenSamples && rc == 0; ++c) {
        for (int i = 0; i < nodeNum; ++i) {                                   /* :370-374 */
            nw[i] = binMat[(int64_t)nodePositions[i] * binMatStride + c];
            simData[nodePositions[i]] = nw[i];
        }
        for (int s = 1; s < simSteps; ++s) {                                  /* :375 */
            memcpy(old, nw, sizeof(int32_t) * nodeNum);
            if (sync_step(&m, old, nw, diag)) { rc = -1; break; }
            for (int i = 0; i < nodeNum; ++i) simData[s * simStride + nodePositions[i]] = nw[i];
        }
        if (rc) break;
        int32_t win[2];
        attract_window(simData, simStride, nodePositions, nodeNum, win);      /* :426 */
        int isFound = win[1] > win[0];
        if (found) found[c] = (int8_t)isFound;
        if (!isFound && diag) {
            diag->not_found_cells++;
            if (!seenFound) diag->undefined_leading++;
        }
        seenFound |= isFound;

        int minErr = 100000;                                                  /* :428 */
        for (int t = win[0]; t < win[1]; ++t) {                               /* :431 */
            total = 0;
            for (int i = 0; i < nodeNum; ++i) {                               /* :446-452 */
                int64_t p = nodePositions[i];
                err[i] = abs(simData[t * simStride + p] - binMat[p * binMatStride + c]);
                total += err[i];
            }
            if (total < minErr) {                                             /* :457 */
                minErr = total;
                if (total <= 0.001 * nodeNum && minErr < 10000) {             /* :459 (Q8) */
                    minErr = 0;
                    if (localSearch) for (int i = 0; i < nodeNum; ++i) errors[i] += err[i];
                    else errors[c] = total;
                }
            }
        }
        if (localSearch) for (int i = 0; i < nodeNum; ++i) errors[i] += err[i]; /* :476-479 */
        else errors[c] = total;                                                 /* :481 */
    }
    free(old); free(nw); free(err); free(own);
    return rc;
}

int scbo_cluster(int sampleIndex, const int32_t *individual, int indLen, int nodeNum,
                 const int32_t *andLenList, const int32_t *individualParse,
                 const int32_t *andNodes, const int32_t *andNodeInvertList,
                 int simSteps, const int32_t *knockouts, const int32_t *knockins,
                 const int32_t *binMat, int64_t binMatStride, const int32_t *nodePositions,
                 int32_t *traj, int32_t res[2], scbo_diag *diag)
{
    if (nodeNum <= 0 || simSteps > SCBO_STEP) return -1;
    net_t m = { nodeNum, indLen, individual, andLenList, individualParse,
                andNodes, andNodeInvertList, knockouts, knockins };
    if (diag) memset(diag, 0, sizeof(*diag));
    /* trajectory kept in node-index order; identity "positions" for the window search.
     * (The reference's KO branch writes simData[step][i] instead of [nodePos], :518 -- a
     * scratch-layout slip that does not change res when nodePositions is a permutation and
     * the scratch starts zeroed only if no node is knocked out; callers here pass no KO
     * or compare res on KO-free cases.) */
    int32_t *ident = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    for (int i = 0; i < nodeNum; ++i) ident[i] = i;
    memset(traj, 0, sizeof(int32_t) * (size_t)SCBO_STEP * nodeNum);
    for (int i = 0; i < nodeNum; ++i)
        traj[i] = binMat[(int64_t)nodePositions[i] * binMatStride + sampleIndex];
    int rc = 0;
    for (int s = 1; s < simSteps; ++s)
        if (sync_step(&m, traj + (int64_t)(s - 1) * nodeNum, traj + (int64_t)s * nodeNum, diag)) {
            rc = -1; break;
        }
    if (!rc) attract_window(traj, nodeNum, ident, nodeNum, res);
    free(ident);
    return rc;
}

/* simulator.c:487-552 `cluster`, LITERALLY, on the caller's scratch: simData[SCBO_STEP][simStride] is read and
 * written exactly like the reference does, including the knock-out branch that records the knocked node in
 * column i instead of column nodePositions[i] (:518) -- so a knocked node's own column keeps whatever the scratch
 * held (or what another node wrote there) and column i may clobber the node recorded there.  The window search
 * (getAttract, :551) then runs over the nodePositions columns of that scratch.  simStride must exceed both
 * max(nodePositions) and nodeNum - 1. */
int scbo_cluster_literal(int sampleIndex, const int32_t *individual, int indLen, int nodeNum,
                         const int32_t *andLenList, const int32_t *individualParse,
                         const int32_t *andNodes, const int32_t *andNodeInvertList,
                         int simSteps, const int32_t *knockouts, const int32_t *knockins,
                         const int32_t *binMat, int64_t binMatStride, const int32_t *nodePositions,
                         int32_t *simData, int64_t simStride, int32_t res[2], scbo_diag *diag)
{
    if (nodeNum <= 0 || simSteps > SCBO_STEP || simStride < nodeNum) return -1;
    if (simStride <= max_pos(nodePositions, nodeNum)) return -1;
    net_t m = { nodeNum, indLen, individual, andLenList, individualParse,
                andNodes, andNodeInvertList, knockouts, knockins };
    if (diag) memset(diag, 0, sizeof(*diag));
    int32_t *old = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    int32_t *nw = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    int rc = 0;
    for (int i = 0; i < nodeNum; ++i) {                                       /* :499-504 */
        nw[i] = binMat[(int64_t)nodePositions[i] * binMatStride + sampleIndex];
        simData[nodePositions[i]] = nw[i];
    }
    for (int s = 1; s < simSteps && !rc; ++s) {                               /* :506 */
        memcpy(old, nw, sizeof(int32_t) * nodeNum);
        if (sync_step(&m, old, nw, diag)) { rc = -1; break; }
        for (int i = 0; i < nodeNum; ++i) {                                   /* writes in the loop order of :512 */
            if (knockouts[i] == 1) simData[s * simStride + i] = nw[i];        /* :518, column i (sic) */
            else simData[s * simStride + nodePositions[i]] = nw[i];
        }
    }
    if (!rc) attract_window(simData, simStride, nodePositions, nodeNum, res); /* :551 */
    free(old); free(nw);
    return rc;
}

/* ---- Python attractor path ------------------------------------------------------------ */

/* singleCell.py:924-959 __updateBool: like rule_value; the inner loop starts at input 0 and
 * guards every input with `> -1` (:944-950); flags are Python bools, compared by value. */
static int py_rule_value(const int32_t *old, const int32_t *bits, int nTerms,
                         const int32_t *an, const int32_t *ai, scbo_diag *diag)
{
    int have = 0, acc = 0;
    for (int a = 0; a < nTerms; ++a) {
        if (!bits[a]) continue;
        int v = old[an[a * 3]] != ai[a * 3];
        for (int b = 0; b < SCBO_MAX_IN; ++b)
            if (an[a * 3 + b] > -1) v = v && (old[an[a * 3 + b]] != ai[a * 3 + b]);
        acc = have ? (acc || v) : v;
        have = 1;
    }
    if (!have) { if (diag) diag->empty_rule_evals++; return 0; } /* Python: NaN -> ValueError */
    return acc;
}

int scbo_py_cluster(const int32_t *individual, int indLen, int nodeNum,
                    const int32_t *andLenList, const int32_t *individualParse,
                    const int32_t *andNodes, const int32_t *andNodeInvertList,
                    int simSteps, const int32_t *knockouts, const int32_t *knockins,
                    int lenSamples, const int32_t *binMat, int64_t binMatStride,
                    const int32_t *nodePositions, int32_t *vals, int32_t *res, scbo_diag *diag)
{
    enum { PY_STEP = 10 };                                    /* singleCell.py:963 */
    if (simSteps != PY_STEP || nodeNum <= 0) return -1;
    if (diag) memset(diag, 0, sizeof(*diag));
    int32_t *old = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    int32_t *nw = (int32_t *)malloc(sizeof(int32_t) * nodeNum);
    for (int c = 0; c < lenSamples; ++c) {
        int32_t *v = vals + (int64_t)c * simSteps * nodeNum;  /* np.full(...,0), :1163-1168 */
        memset(v, 0, sizeof(int32_t) * (size_t)simSteps * nodeNum);
        for (int i = 0; i < nodeNum; ++i) {                   /* :1011-1012 */
            nw[i] = binMat[(int64_t)nodePositions[i] * binMatStride + c];
            v[i] = nw[i];
        }
        for (int s = 1; s < simSteps; ++s) {                  /* :1014-1055 */
            memcpy(old, nw, sizeof(int32_t) * nodeNum);
            for (int i = 0; i < nodeNum; ++i) {
                int t;
                if (knockouts[i] == 1) t = 0;
                else if (knockins[i] == 1) t = 1;
                else if (andLenList[i] == 1)
                    t = old[TBL(andNodes, i, 0, 0)] != TBL(andNodeInvertList, i, 0, 0);
                else if (andLenList[i] == 0) t = old[i];      /* :1032-1033 holds own value */
                else if (i == nodeNum - 1) continue;          /* :1037-1038: never updated, never recorded (Q12) */
                else {
                    int start = individualParse[i], end = individualParse[i + 1];
                    t = py_rule_value(old, individual + start, end - start,
                                      &TBL(andNodes, i, 0, 0), &TBL(andNodeInvertList, i, 0, 0), diag);
                }
                nw[i] = t;
                v[(int64_t)s * nodeNum + i] = t;
            }
        }
        int32_t *r = res + 2 * (int64_t)c;                    /* :961-982, STEP = 10, init [-1,-1] */
        r[0] = -1; r[1] = -1;
        for (int i = PY_STEP - 1; i >= 0 && r[1] < 0; --i)
            for (int j = i - 1; j >= 0; --j)
                if (!memcmp(v + (int64_t)i * nodeNum, v + (int64_t)j * nodeNum,
                            sizeof(int32_t) * nodeNum)) { r[0] = j; r[1] = i; break; }
    }
    free(old); free(nw);
    return 0;
}

int scbo_hamming_argmin(const int32_t *attractors, int nAttr, int nodeNum, int lenSamples,
                        const int32_t *binMat, int64_t binMatStride,
                        const int32_t *nodePositions, int32_t *dist, int32_t *decider,
                        int32_t *deciderDistance)
{
    for (int c = 0; c < lenSamples; ++c) {
        int best = -1, bestD = 0;
        for (int a = 0; a < nAttr; ++a) {
            int d = 0;
            for (int i = 0; i < nodeNum; ++i)                 /* singleCell.py:1222-1227 */
                d += abs(attractors[(int64_t)a * nodeNum + i]
                         - binMat[(int64_t)nodePositions[i] * binMatStride + c]);
            if (dist) dist[(int64_t)c * nAttr + a] = d;
            if (best < 0 || d < bestD) { best = a; bestD = d; } /* idxmin: first minimum, :1235 */
        }
        if (decider) decider[c] = best;
        if (deciderDistance) deciderDistance[c] = bestD;
    }
    return 0;
}

/* ---- importanceScore (simulator.c:555-755) ---------------------------------------------- */

static int importance_pass(const net_t *m, const int32_t *knock, int simSteps, int lenSamples,
                           const int32_t *binMat, int64_t binMatStride, const int32_t *pos,
                           double *avg)
{
    int n = m->nodeNum, rc = 0;
    int32_t *traj = (int32_t *)calloc((size_t)SCBO_STEP * n, sizeof(int32_t));
    int32_t *ident = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; ++i) ident[i] = i;
    for (int c = 0; c < lenSamples && !rc; ++c) {
        for (int i = 0; i < n; ++i) traj[i] = binMat[(int64_t)pos[i] * binMatStride + c];
        for (int s = 1; s < simSteps; ++s) {
            const int32_t *old = traj + (int64_t)(s - 1) * n;
            int32_t *nw = traj + (int64_t)s * n;
            for (int i = 0; i < n; ++i) {
                if (knock[i]) nw[i] = 0;                       /* :604 and, sic, :674 */
                else if (m->andLen[i])                         /* :611 every rule node is 1-input */
                    nw[i] = old[TBL(m->an, i, 0, 0)] != TBL(m->ai, i, 0, 0);
                else nw[i] = old[TBL(m->an, i, 0, 0)];         /* :617 */
            }
        }
        int32_t w[2];
        attract_window(traj, n, ident, n, w);
        int len = w[1] - w[0];
        if (len == 0) { rc = -2; break; }                      /* reference: SIGFPE at :647 */
        for (int t = w[0]; t <= w[1]; ++t)                     /* inclusive window, :641 */
            for (int i = 0; i < n; ++i)
                avg[i] += (double)(traj[(int64_t)t * n + i] / len); /* integer division, :647 */
    }
    free(traj); free(ident);
    return rc;
}

int scbo_importance_score(const int32_t *individual, int indLen, int nodeNum,
                          const int32_t *andLenList, const int32_t *individualParse,
                          const int32_t *andNodes, const int32_t *andNodeInvertList,
                          int simSteps, const int32_t *knockouts, const int32_t *knockins,
                          int lenSamples, const int32_t *binMat, int64_t binMatStride,
                          const int32_t *nodePositions, double *score)
{
    if (nodeNum <= 0 || simSteps > SCBO_STEP) return -1;
    net_t m = { nodeNum, indLen, individual, andLenList, individualParse,
                andNodes, andNodeInvertList, knockouts, knockins };
    double *ko = (double *)calloc(nodeNum, sizeof(double));
    double *ki = (double *)calloc(nodeNum, sizeof(double));
    int rc = importance_pass(&m, knockouts, simSteps, lenSamples, binMat, binMatStride, nodePositions, ko);
    /* the knock-in pass also accumulates into the knock-OUT average (:717) */
    if (!rc) rc = importance_pass(&m, knockins, simSteps, lenSamples, binMat, binMatStride, nodePositions, ko);
    score[0] = 0.0;
    if (!rc) {
        for (int i = 0; i < nodeNum; ++i) { ki[i] /= lenSamples; ko[i] /= lenSamples; } /* :734-741 */
        for (int i = 0; i < nodeNum; ++i) score[0] += fabs(ki[i] - ko[i]);              /* :750 */
    }
    free(ko); free(ki);
    return rc;
}
