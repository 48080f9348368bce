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

/* TODO(round 1, in progress): K4/K6 kernels */
int scb_attractors(scb_ctx *, const int32_t *, int, const int32_t *, const int32_t *, const int32_t *, const int32_t *,
                   const int32_t *, const int32_t *, int, int, int32_t *, int32_t *, uint32_t *)
{
    return fail(SCB_ERR_UNSUPPORTED, "scb_attractors: not implemented yet");
}
void importanceScore(void *, const int32_t *, intptr_t, intptr_t, const int32_t *, const int32_t *, const int32_t *,
                     const int32_t *, intptr_t, const int32_t *, const int32_t *, intptr_t, const int32_t *,
                     const int32_t *, double *)
{
    fprintf(stderr, "scbonita_b200: importanceScore not implemented yet\n");
}
void cluster(void *, int32_t *, intptr_t, const int32_t *, intptr_t, intptr_t, const int32_t *, const int32_t *,
             const int32_t *, const int32_t *, intptr_t, const int32_t *, const int32_t *, const int32_t *,
             const int32_t *)
{
    fprintf(stderr, "scbonita_b200: cluster not implemented yet\n");
}

}  /* extern "C" */
