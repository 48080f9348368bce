/*
 * scb_legacy.inl -- the reference's exported symbols (drop-in for ./simulator.so), included by
 * scb_api.cu.  Same argument lists as simulator.c:350 (scSyncBool), :555 (importanceScore) and
 * :487 (cluster); geometry that the reference fixes at compile time (NODE / CELL,
 * simulator.c:7-8) is a runtime setting here.
 */
namespace {
int g_legacyNode = 20000;   /* simulator.c:7 */
int g_legacyCell = 15000;   /* simulator.c:8 */
std::once_flag g_legacyEnvOnce;

void legacy_env()
{
    std::call_once(g_legacyEnvOnce, [] {
        if (const char *e = getenv("SCB_LEGACY_CELL")) { int v = atoi(e); if (v > 0) g_legacyCell = v; }
        if (const char *e = getenv("SCB_LEGACY_NODE")) { int v = atoi(e); if (v > 0) g_legacyNode = v; }
    });
}
}  // namespace

extern "C" {

void scb_set_legacy_geometry(int nodeStride, int cellStride)
{
    legacy_env();
    if (nodeStride > 0) g_legacyNode = nodeStride;
    if (cellStride > 0) g_legacyCell = cellStride;
}

/* simulator.c:350-485.  errors: localSearch == 0 -> errors[c] assigned for c < lenSamples (:481);
 * localSearch != 0 -> errors[i] accumulated for i < nodeNum (:477-479).  simData is scratch the
 * Python never reads back (ruleMaker.py:1131-1170) and is left untouched. */
void scSyncBool(void *simData, const int32_t *individual, intptr_t indLen, intptr_t nodeNum,
                const int32_t *andLenList, const int32_t *individualParse, const int32_t *andNodes,
                const int32_t *andNodeInvertList, intptr_t simSteps, const int32_t *knockouts,
                const int32_t *knockins, intptr_t lenSamples, const int32_t *binMat,
                const int32_t *nodePositions, int32_t *errors, intptr_t localSearch, intptr_t importanceScores)
{
    (void)simData; (void)importanceScores;
    legacy_env();
    const int N = (int)nodeNum, C = (int)lenSamples, L = (int)indLen;
    if ((int)simSteps != SCB_STEPS) {
        fprintf(stderr, "scbonita_b200: scSyncBool supports simSteps == 100 only (got %d); errors left untouched\n", (int)simSteps);
        return;
    }
    if (C <= 0 || N <= 0) return;
    scb_ctx *ctx = nullptr;
    int rc = scb_ctx_create(&ctx, 0, N, C, binMat, 4, g_legacyCell, nodePositions);
    if (rc == SCB_OK) {
        const int mode = localSearch ? SCB_MODE_NODEWISE : SCB_MODE_CELLWISE;
        std::vector<int32_t> out((size_t)(localSearch ? N : C));
        rc = scb_eval_population(ctx, 1, individual, L, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                                 knockouts, knockins, mode, out.data(), nullptr);
        if (rc == SCB_OK) {
            if (localSearch) for (int i = 0; i < N; ++i) errors[i] += out[i];
            else for (int c = 0; c < C; ++c) errors[c] = out[c];
        }
    }
    if (rc != SCB_OK) fprintf(stderr, "scbonita_b200: scSyncBool failed (%d): %s\n", rc, scb_last_error());
    scb_ctx_destroy(ctx);
}

/* trajectories + attractor windows of the uploaded single individual (P == 1 already on the device) */
static int run_trajectories(scb_ctx *c, int sem, int steps, int maxStates, int zeroNode)
{
    const int N = c->N, tiles = (c->W + 31) / 32, NW = (N + 31) / 32;
    int rc;
    if (scb_smem_bytes(N, 1, 2) > c->smemLimit) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d too large for the trajectory kernel", N);
    if ((rc = c->dTraj.ensure((size_t)tiles * steps * N * 32))) return rc;
    if ((rc = c->dRes.ensure((size_t)c->C * 2)) || (rc = c->dNStates.ensure(c->C))) return rc;
    if ((rc = c->dStates.ensure((size_t)c->C * std::max(maxStates, 1) * NW))) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), st));
    if ((rc = launch_compile(c, sem, c->dKo.p))) return rc;
    ScbTrajArgs ta;
    ta.X = c->dX.p; ta.N = N; ta.C = c->C; ta.Wpad = c->Wpad; ta.tiles = tiles; ta.steps = steps;
    ta.descs = c->dDescs.p; ta.counts = c->dCounts.p; ta.traj = c->dTraj.p; ta.zeroNode = zeroNode;
    const int warps = std::min(16, std::max(1, (N + 3) / 4));
    const unsigned grid = (unsigned)std::max(1, std::min(tiles, c->numSM * 2));
    SCB_LAUNCH(scb_k_traj, dim3(grid), dim3((unsigned)warps * 32), scb_smem_bytes(N, 1, 2), st, ta);
    CK(cudaGetLastError());
    ScbWindowArgs wa;
    wa.traj = c->dTraj.p; wa.N = N; wa.C = c->C; wa.tiles = tiles; wa.steps = steps; wa.sem = sem == SCB_SEM_PY ? 1 : 0;
    wa.maxStates = maxStates; wa.NW = NW; wa.res = c->dRes.p; wa.nStates = c->dNStates.p; wa.states = c->dStates.p;
    const int nPairs = wa.sem == 0 ? steps - 1 : steps * (steps - 1) / 2;
    SCB_LAUNCH(scb_k_window, dim3(grid), dim3(256), (size_t)nPairs * 128, st, wa);
    CK(cudaGetLastError());
    return SCB_OK;
}

static int check_compile_diag(scb_ctx *c)
{
    CK(cudaMemcpyAsync(&c->hDiag, c->dDiag.p, sizeof(ScbDiag), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const ScbDiag &d = c->hDiag;
    if (d.bad_index_nodes) return fail(SCB_ERR_INVALID, "%llu rules reference an input outside [0,%d)", d.bad_index_nodes, c->N);
    if (d.term_overflow_nodes) return fail(SCB_ERR_INVALID, "%llu rules span more than 7 AND-terms", d.term_overflow_nodes);
    if (d.unsupported_nodes) return fail(SCB_ERR_UNSUPPORTED, "%llu rules use more than 3 distinct inputs", d.unsupported_nodes);
    return SCB_OK;
}

int scb_attractors(scb_ctx *c, const int32_t *individual, int indLen, const int32_t *andLenList,
                   const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                   const int32_t *knockouts, const int32_t *knockins, int semantics, int maxStates,
                   int32_t *res, int32_t *nStates, uint32_t *states)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (semantics != 0 && semantics != 1) return fail(SCB_ERR_INVALID, "semantics must be 0 (C cluster) or 1 (Python __cluster)");
    if (maxStates < 0 || !res || !nStates || (maxStates > 0 && !states)) return fail(SCB_ERR_INVALID, "bad argument");
    rc = scb_pop_upload(c, 1, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0, knockouts, knockins);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    const int N = c->N, NW = (N + 31) / 32;
    int zeroNode = -1;
    if (semantics == 1) {
        /* singleCell.py:1037-1038: the last node falls through without being updated or recorded */
        const int al = andLenList[N - 1];
        if (knockouts[N - 1] != 1 && knockins[N - 1] != 1 && al != 1 && al != 0) zeroNode = N - 1;
    }
    rc = run_trajectories(c, semantics == 1 ? SCB_SEM_PY : SCB_SEM_C, semantics == 1 ? 10 : SCB_STEPS, maxStates, zeroNode);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(res, c->dRes.p, (size_t)c->C * 2 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nStates, c->dNStates.p, (size_t)c->C * 4, cudaMemcpyDeviceToHost, st));
    if (maxStates > 0) CK(cudaMemcpyAsync(states, c->dStates.p, (size_t)c->C * maxStates * NW * 4, cudaMemcpyDeviceToHost, st));
    return check_compile_diag(c);
}

/* simulator.c:487-552.  One cell (`sampleIndex`); resSubmit = attractor window; the trajectory is left in
 * simData[step][nodePositions[i]] (row stride = legacy NODE) like the reference does.  The reference's
 * knock-out branch writes column i instead of nodePositions[i] (simulator.c:518); this shim writes the true
 * trajectory at nodePositions[i] for every node, which is identical whenever no node is knocked out or
 * nodePositions is the identity. */
void cluster(void *simData, int32_t *resSubmit, intptr_t sampleIndex, const int32_t *individual, intptr_t indLen,
             intptr_t nodeNum, const int32_t *andLenList, const int32_t *individualParse, const int32_t *andNodes,
             const int32_t *andNodeInvertList, intptr_t simSteps, const int32_t *knockouts, const int32_t *knockins,
             const int32_t *binMat, const int32_t *nodePositions)
{
    legacy_env();
    const int N = (int)nodeNum;
    if ((int)simSteps != SCB_STEPS) {
        fprintf(stderr, "scbonita_b200: cluster supports simSteps == 100 only (got %d)\n", (int)simSteps);
        return;
    }
    scb_ctx *ctx = nullptr;
    int rc = scb_ctx_create(&ctx, 0, N, 1, binMat + (long long)sampleIndex, 4, g_legacyCell, nodePositions);
    if (rc == SCB_OK) {
        rc = scb_pop_upload(ctx, 1, individual, (int)indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                            knockouts, knockins);
        if (rc == SCB_OK) rc = run_trajectories(ctx, SCB_SEM_C, SCB_STEPS, 0, -1);
        std::vector<uint32_t> traj((size_t)SCB_STEPS * N * 32);
        int32_t r2[2] = {0, 0};
        if (rc == SCB_OK) {
            cudaMemcpyAsync(traj.data(), ctx->dTraj.p, traj.size() * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(r2, ctx->dRes.p, 8, cudaMemcpyDeviceToHost, ctx->stream);
            rc = check_compile_diag(ctx);
        }
        if (rc == SCB_OK) {
            int32_t *sd = static_cast<int32_t *>(simData);
            if (sd)
                for (int t = 0; t < SCB_STEPS; ++t)
                    for (int i = 0; i < N; ++i)
                        sd[(long long)t * g_legacyNode + nodePositions[i]] = (int32_t)(traj[((size_t)t * N + i) * 32] & 1u);
            resSubmit[0] = r2[0]; resSubmit[1] = r2[1];
        }
    }
    if (rc != SCB_OK) fprintf(stderr, "scbonita_b200: cluster failed (%d): %s\n", rc, scb_last_error());
    scb_ctx_destroy(ctx);
}

/* simulator.c:555-755, incl. its quirks: both passes force knocked nodes to 0 and accumulate into
 * the knock-OUT average (:674, :717), every node with andLen != 0 is a single literal, the individual is
 * ignored, (int)(state / window length) is an integer division, so only windows of length 1 (S_98 == S_99)
 * contribute, with their two rows 98 and 99.  Batched: item k is the pair (knockouts[k][N], knockins[k][N]) --
 * ruleMaker.__calcImportance makes 3 x N such calls per network (ruleMaker.py:1285-1300).  Every distinct knock
 * array is one program of ONE simulator launch (K0 with the importance semantics, K2/K3 in their importance output
 * shape); a pass whose knock-in array equals its knock-out array (what __calcImportance passes) is simulated once.
 * The integer counts come from the device; the last N divisions and additions are done here in the reference's
 * order so that every double is bit-identical.  A cell without an attractor window makes the reference divide
 * by zero (SIGFPE): that item's score is NaN and the call returns SCB_ERR_INVALID. */
int scb_importance_scores(scb_ctx *ctx, const int32_t *individual, int indLen, const int32_t *andLenList,
                          const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                          int nItems, const int32_t *knockouts, const int32_t *knockins, double *scores)
{
    int rc = check_ctx(ctx);
    if (rc) return rc;
    if (!scores || !knockouts || !knockins) return fail(SCB_ERR_INVALID, "null argument");
    if (nItems <= 0) return fail(SCB_ERR_INVALID, "nItems must be positive");
    const int N = ctx->N, C = ctx->C;
    for (int k = 0; k < nItems; ++k) scores[k] = 0.0;
    std::vector<int32_t> knock;
    std::vector<int> passOf((size_t)2 * nItems);
    knock.reserve((size_t)2 * nItems * N);
    for (int k = 0; k < nItems; ++k) {
        const int32_t *ko = knockouts + (size_t)k * N, *ki = knockins + (size_t)k * N;
        passOf[2 * k] = (int)(knock.size() / N);
        knock.insert(knock.end(), ko, ko + N);
        /* simulator.c:604 / :674 test the arrays for non-zero only */
        bool same = true;
        for (int i = 0; i < N && same; ++i) same = (ko[i] != 0) == (ki[i] != 0);
        if (same) passOf[2 * k + 1] = passOf[2 * k];
        else { passOf[2 * k + 1] = (int)(knock.size() / N); knock.insert(knock.end(), ki, ki + N); }
    }
    const int P = (int)(knock.size() / N);
    rc = pop_upload_impl(ctx, P, 1, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                         P, knock.data(), knockins);
    if (rc) return rc;
    if ((rc = pop_run_impl(ctx, SCB_MODE_IMPORTANCE, SCB_SEM_IMPORTANCE))) return rc;
    std::vector<int32_t> cnt((size_t)P * N);
    std::vector<unsigned long long> noWindow((size_t)P);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CK(cudaMemcpyAsync(cnt.data(), ctx->dNodewise.p, cnt.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(noWindow.data(), ctx->dFitness.p, noWindow.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = check_compile_diag(ctx))) return rc;
        ctx->ran = false;                       /* the buffers do not hold a population result */
    }
    unsigned long long bad = 0;
    int badItems = 0;
    for (int k = 0; k < nItems; ++k) {
        const int pa = passOf[2 * k], pb = passOf[2 * k + 1];
        const unsigned long long nf = noWindow[pa] + noWindow[pb];
        if (nf) { scores[k] = nan(""); bad += nf; ++badItems; continue; }
        double acc = 0.0;
        for (int i = 0; i < N; ++i) {
            /* rows 98 and 99 of every fixed-point cell, knock-out pass then knock-in pass (:641-648, :712-719) */
            const long long sum = 2ll * cnt[(size_t)pa * N + i] + 2ll * cnt[(size_t)pb * N + i];
            const double ki = 0.0 / C;                                   /* attractorAverage_ki stays 0 (:734-736) */
            const double ko = (double)sum / C;                           /* :739-741 */
            acc = acc + fabs(ki - ko);                                   /* :750 */
        }
        scores[k] = acc;
    }
    if (bad)
        return fail(SCB_ERR_INVALID, "%llu cell passes of %d item(s) have no attractor window; the reference divides by "
                                     "zero there (simulator.c:647)", bad, badItems);
    return SCB_OK;
}

int scb_importance_score(scb_ctx *ctx, const int32_t *individual, int indLen, const int32_t *andLenList,
                         const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                         const int32_t *knockouts, const int32_t *knockins, double *score)
{
    return scb_importance_scores(ctx, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 1,
                                 knockouts, knockins, score);
}

void importanceScore(void *simData, const int32_t *individual, intptr_t indLen, intptr_t nodeNum,
                     const int32_t *andLenList, const int32_t *individualParse, const int32_t *andNodes,
                     const int32_t *andNodeInvertList, intptr_t simSteps, const int32_t *knockouts,
                     const int32_t *knockins, intptr_t lenSamples, const int32_t *binMat,
                     const int32_t *nodePositions, double *importanceScores)
{
    (void)simData;
    legacy_env();
    const int N = (int)nodeNum, C = (int)lenSamples;
    if (!importanceScores) return;
    importanceScores[0] = 0.0;
    if ((int)simSteps != SCB_STEPS) {
        fprintf(stderr, "scbonita_b200: importanceScore supports simSteps == 100 only (got %d)\n", (int)simSteps);
        return;
    }
    if (N <= 0 || C <= 0) return;
    scb_ctx *ctx = nullptr;
    int rc = scb_ctx_create(&ctx, 0, N, C, binMat, 4, g_legacyCell, nodePositions);
    if (rc == SCB_OK)
        rc = scb_importance_score(ctx, individual, (int)indLen, andLenList, individualParse, andNodes, andNodeInvertList,
                                  knockouts, knockins, importanceScores);
    if (rc != SCB_OK) fprintf(stderr, "scbonita_b200: importanceScore failed (%d): %s\n", rc, scb_last_error());
    scb_ctx_destroy(ctx);
}

}  /* extern "C" */
