/*
 * scb_compile.h -- turn one node's reference-format rule into (<= 3 inputs, 8-bit truth table).
 *
 * Semantics restated from the reference (file:line under /root/reference/src/scBONITA/):
 *   simulator.c:383-394  knockouts[i]==1 -> 0, else knockins[i]==1 -> 1
 *   simulator.c:395-399  andLenList[i]==1 -> old[andNodes[i][0][0]] != andNodeInvertList[i][0][0]
 *   simulator.c:400-404  andLenList[i]==0 -> old[andNodes[i][0][0]]
 *   simulator.c:405-411  otherwise updateBool over bits [parse[i], parse[i+1]) (indLen for the
 *                        last node), simulator.c:10-50: OR over selected terms of the AND of
 *                        `old[in] != flag`, input 0 always, inputs 1,2 only when index > -1.
 * `old != flag` is an int comparison on 0/1 values: flag 0 = identity, 1 = NOT, anything else
 * = constant TRUE.  Because every literal is a function of one state bit, a node
 * whose selected terms touch at most three distinct state bits is exactly a 3-input truth
 * table, which the simulator evaluates with one LOP3 per 32 cells.
 */
#ifndef SCB_COMPILE_H
#define SCB_COMPILE_H
#include "scb_types.h"

struct ScbNodeFn {
    int arity;          /* number of inputs the function really depends on (0..3) */
    int in[3];
    unsigned lut;       /* 8-bit table, index (x0<<2)|(x1<<1)|x2; don't-care inputs duplicated */
    unsigned flags;     /* SCB_NF_* */
};

SCB_HD inline ScbNodeFn scb_const_fn(int v, unsigned flags)
{
    ScbNodeFn f;
    f.arity = 0; f.in[0] = f.in[1] = f.in[2] = 0;
    f.lut = v ? 0xFFu : 0x00u; f.flags = flags;
    return f;
}

/* sem selects whose update rule is restated:
 *   SCB_SEM_C      scSyncBool / cluster                   (simulator.c:377-416, :509-549)
 *   SCB_SEM_PY     singleCell.__cluster                   (singleCell.py:1014-1055): a zero-input node holds its
 *                  OWN value (:1032-1033) and the last node, unless knocked / single-input / zero-input, is
 *                  never updated (:1037-1038) -- the caller blanks its recorded rows
 *   SCB_SEM_IMPORTANCE importanceScore                    (simulator.c:600-622): `ko` is the knock array of the pass,
 *                  any nonzero entry forces 0; every node with andLen != 0 is the single literal of its first
 *                  term's first input, the individual is ignored */
#define SCB_SEM_C 0
#define SCB_SEM_PY 1
#define SCB_SEM_IMPORTANCE 2

SCB_HD inline ScbNodeFn scb_copy_fn(int src, int invert)
{
    ScbNodeFn f;
    f.arity = 1; f.in[0] = src; f.in[1] = src; f.in[2] = src;
    f.lut = invert ? 0x0Fu : 0xF0u; f.flags = 0;
    return f;
}

/* a table over vars[0..nv) -> the node function with the inputs the table does not depend on dropped
 * (a node costs only the loads it needs); don't-care positions repeat the first kept input */
SCB_HD inline ScbNodeFn scb_finish_fn(const int *vars, int nv, unsigned lut, unsigned flags)
{
    const bool dep[3] = { (((lut >> 4) ^ lut) & 0x0Fu) != 0,
                          (((lut >> 2) ^ lut) & 0x33u) != 0,
                          (((lut >> 1) ^ lut) & 0x55u) != 0 };
    int keep[3] = {0, 0, 0}, nk = 0;
    for (int k = 0; k < 3; ++k) if (k < nv && dep[k]) keep[nk++] = k;
    unsigned nl = 0;
    for (int idx = 0; idx < 8; ++idx) {
        int m = 0;
        for (int j = 0; j < nk; ++j) if ((idx >> (2 - j)) & 1) m |= 1 << (2 - keep[j]);
        nl |= ((lut >> m) & 1u) << idx;
    }
    ScbNodeFn f;
    f.arity = nk; f.flags = flags; f.lut = nl;
    const int fill = nk ? vars[keep[0]] : 0;
    for (int j = 0; j < 3; ++j) f.in[j] = j < nk ? vars[keep[j]] : fill;
    return f;
}

/* The AND-term tables of one node from its (at most three) upstream nodes, as the reference builds them:
 * every non-empty subset of the inputs in itertools.product order over ((a, empty), (b, empty), (c, empty)),
 * i.e. [abc, ab, ac, a, bc, b, c] (ruleMaker.py:188-200, :300-335), each term padded with -1 and the table
 * padded to 7 terms with [0,0,0] (__updateCpointers, ruleMaker.py:337-360).  up[3] holds the inputs as a prefix
 * (-1 after the last), upInv[3] their inversion flags.  Returns the number of terms, -1 for a malformed triple. */
SCB_HD inline int scb_expand_terms(const int32_t *up, const int32_t *upInv, int32_t *an, int32_t *ai)
{
    int k = 0;
    while (k < SCB_MAX_IN && up[k] >= 0) ++k;
    bool bad = false;
    for (int j = k; j < SCB_MAX_IN; ++j) bad = bad || up[j] >= 0;
    for (int x = 0; x < SCB_MAX_TERMS * SCB_MAX_IN; ++x) { an[x] = 0; ai[x] = 0; }
    if (bad) return -1;
    const int nt = (1 << k) - 1;
    for (int m = 0; m < nt; ++m) {
        int w = 0;
        for (int j = 0; j < k; ++j) {
            if ((m >> (k - 1 - j)) & 1) continue;                 /* input j is "empty" in this term */
            an[m * SCB_MAX_IN + w] = up[j];
            ai[m * SCB_MAX_IN + w] = upInv[j];
            ++w;
        }
        for (; w < SCB_MAX_IN; ++w) { an[m * SCB_MAX_IN + w] = -1; ai[m * SCB_MAX_IN + w] = -1; }
    }
    return nt;
}

SCB_HD inline ScbNodeFn scb_compile_node(int i, int N, int indLen, const int32_t *ind,
                                         const int32_t *andLen, const int32_t *parse,
                                         const int32_t *an, const int32_t *ai,
                                         const int32_t *ko, const int32_t *ki, int sem = SCB_SEM_C)
{
    const int32_t *an_i = an + (long long)i * (SCB_MAX_TERMS * SCB_MAX_IN);
    const int32_t *ai_i = ai + (long long)i * (SCB_MAX_TERMS * SCB_MAX_IN);
    const int al = andLen[i];
    if (sem == SCB_SEM_IMPORTANCE) {
        if (ko[i]) return scb_const_fn(0, 0);                                   /* simulator.c:604 */
        const int src = an_i[0];
        if (src < 0 || src >= N) return scb_const_fn(0, SCB_NF_BADIDX);
        if (al == 0) return scb_copy_fn(src, 0);                                /* :617 */
        const int fl = ai_i[0];                                                 /* :611 */
        if (fl != 0 && fl != 1) return scb_const_fn(1, 0);
        return scb_copy_fn(src, fl);
    }
    if (ko[i] == 1) return scb_const_fn(0, 0);
    if (ki[i] == 1) return scb_const_fn(1, 0);
    if (sem == SCB_SEM_PY && al == 0) return scb_copy_fn(i, 0);
    if (sem == SCB_SEM_PY && al != 1 && i == N - 1) return scb_copy_fn(i, 0);
    if (al == 1 || al == 0) {
        const int src = an_i[0];
        if (src < 0 || src >= N) return scb_const_fn(0, SCB_NF_BADIDX);
        const int fl = (al == 1) ? ai_i[0] : 0;
        if (fl != 0 && fl != 1) return scb_const_fn(1, 0);
        ScbNodeFn f;
        f.arity = 1; f.in[0] = src; f.in[1] = src; f.in[2] = src;
        f.lut = fl ? 0x0Fu : 0xF0u; f.flags = 0;
        return f;
    }
    unsigned flags = 0;
    const int start = parse[i];
    const int end = (i == N - 1) ? indLen : parse[i + 1];
    int nt = end - start;
    if (nt > SCB_MAX_TERMS) { flags |= SCB_NF_TERMS; nt = SCB_MAX_TERMS; }
    if (nt < 0 || start < 0 || end > indLen) { flags |= SCB_NF_TERMS; nt = 0; }

    /* One pass over the selected terms: a literal over variable k (the k-th distinct state bit met, k < 3) is the
     * 8-minterm mask VAR[k] or its complement, a term the AND of its literals, the rule the OR of its terms. */
    unsigned sel = 0u;
    for (int a = 0; a < nt; ++a) sel |= ind[start + a] ? 1u << a : 0u;
    if (!sel) return scb_const_fn(0, flags | SCB_NF_EMPTY);   /* Q9: reference reads orset[0] uninitialised */
    int v0 = 0, v1 = 0, v2 = 0, nv = 0;                  /* scalars, not an indexed array: no local memory on the device */
    bool overflow = false;
    unsigned lut = 0;
    for (int a = 0; a < nt; ++a) {
        if (!((sel >> a) & 1u)) continue;
        unsigned t = 0xFFu;
        for (int b = 0; b < SCB_MAX_IN; ++b) {
            const int src = an_i[a * 3 + b];
            if (b > 0 && src <= -1) continue;
            if (src < 0 || src >= N) { flags |= SCB_NF_BADIDX; continue; }
            const int fl = ai_i[a * 3 + b];
            if (fl != 0 && fl != 1) continue;            /* constant TRUE literal */
            int k = (nv > 0 && v0 == src) ? 0 : ((nv > 1 && v1 == src) ? 1 : ((nv > 2 && v2 == src) ? 2 : 3));
            if (k == 3) {
                if (nv == 3) { overflow = true; continue; }
                k = nv++;
                if (k == 0) v0 = src; else if (k == 1) v1 = src; else v2 = src;
            }
            const unsigned var = k == 0 ? 0xF0u : (k == 1 ? 0xCCu : 0xAAu);
            t &= fl ? ~var : var;
        }
        lut |= t & 0xFFu;
    }
    const int vars[3] = {v0, v1, v2};
    if (overflow) return scb_const_fn(0, flags | SCB_NF_UNSUPPORTED);
    return scb_finish_fn(vars, nv, lut, flags);
}

/* Constant propagation (the steady program of K2): `f` with every input whose value is known (cst[node] = 0 or 1,
 * anything else = unknown) fixed to that value, dead inputs dropped again.  Exact wherever the inputs really hold
 * those values. */
/* table `lut` with input j (0 = the 4-weight bit of the index) fixed to v */
SCB_HD inline unsigned scb_lut_fix(unsigned lut, int j, unsigned v)
{
    const unsigned var = j == 0 ? 0xF0u : (j == 1 ? 0xCCu : 0xAAu);
    const int sh = 4 >> j;
    if (v) { const unsigned hi = lut & var; return hi | (hi >> sh); }
    const unsigned lo = lut & ~var & 0xFFu;
    return lo | (lo << sh);
}

/* the table of `f` with every known input (cst[node] = 0 or 1) fixed */
SCB_HD inline unsigned scb_restrict_lut(const ScbNodeFn &f, const unsigned char *cst)
{
    unsigned lut = f.lut;
    for (int j = 0; j < f.arity; ++j) {
        const unsigned v = cst[f.in[j]];
        if (v <= 1u) lut = scb_lut_fix(lut, j, v);
    }
    return lut;
}

SCB_HD inline ScbNodeFn scb_restrict_fn(const ScbNodeFn &f, const unsigned char *cst)
{
    return scb_finish_fn(f.in, f.arity, scb_restrict_lut(f, cst), 0);
}

/* generic table lookup on 32 cells at once (host tests and the emulator; the device path uses
 * lop3 with an immediate table) */
SCB_HD inline uint32_t scb_lut_apply(unsigned lut, uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = 0;
    for (int idx = 0; idx < 8; ++idx)
        if ((lut >> idx) & 1u)
            r |= ((idx & 4) ? a : ~a) & ((idx & 2) ? b : ~b) & ((idx & 1) ? c : ~c);
    return r;
}

#endif
