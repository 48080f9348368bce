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
/* The reference's symbols return void and have no error convention; its own failure mode is a crash.  A failed
 * evaluation must never look like a result: the unmodified Python scores `sum(errors)` and a GA that minimises
 * error would read untouched zeros as a perfect individual.  Default: message + abort().  With
 * SCB_LEGACY_ON_ERROR=sentinel the outputs are filled with values no evaluation can produce and the call returns. */
bool legacy_sentinel_mode()
{
    const char *e = getenv("SCB_LEGACY_ON_ERROR");
    return e && strcmp(e, "sentinel") == 0;
}

[[noreturn]] void legacy_die(const char *fn, const char *msg)
{
    fprintf(stderr, "scbonita_b200: %s failed: %s\n", fn, msg);
    fflush(stderr);
    abort();
}

void legacy_fatal(const char *fn, const char *msg)
{
    if (!legacy_sentinel_mode()) legacy_die(fn, msg);
    fprintf(stderr, "scbonita_b200: %s failed: %s (outputs set to sentinels)\n", fn, msg);
}

/* ---- packed contexts of the legacy symbols, cached by content -------------------------------------------------
 * The reference's Python passes the same dense matrix on every call (2 538 local-search calls per network,
 * ruleMaker.py:1247-1259); packing it again and again (cudaMalloc + H2D + K1) dominated a level-0 drop-in call.
 * Key: two independent 64-bit hashes of the N used rows (lenSamples ints each) and of nodePositions, plus the sizes
 * and the device.  A few entries, least recently used out first; an entry in use is never evicted. */
struct LegacyEntry {
    uint64_t h1 = 0, h2 = 0;
    int N = 0, C = 0, device = 0;
    scb_ctx *ctx = nullptr;
    int inUse = 0;
    unsigned long long stamp = 0;
};
std::mutex g_legacyMu;
LegacyEntry g_legacyCache[4];
unsigned long long g_legacyClock = 0;

int legacy_device()
{
    const char *e = getenv("SCB_DEVICE");
    return e ? atoi(e) : 0;
}

void legacy_hash(const int32_t *binMat, long long rowStride, const int32_t *pos, int N, int C, uint64_t *h1, uint64_t *h2)
{
    uint64_t a = 0x9E3779B97F4A7C15ull ^ (uint64_t)N, b = 0xC2B2AE3D27D4EB4Full ^ (uint64_t)C;
    for (int i = 0; i < N; ++i) {
        const int32_t *row = binMat + (long long)pos[i] * rowStride;
        a = (a ^ (uint64_t)(uint32_t)pos[i]) * 0x100000001B3ull;
        b = (b + (uint64_t)(uint32_t)pos[i]) * 0xFF51AFD7ED558CCDull; b ^= b >> 29;
        int c = 0;
        for (; c + 1 < C; c += 2) {
            const uint64_t v = (uint64_t)(uint32_t)row[c] | ((uint64_t)(uint32_t)row[c + 1] << 32);
            a = (a ^ v) * 0x100000001B3ull; a ^= a >> 31;
            b = (b + v) * 0xFF51AFD7ED558CCDull; b ^= b >> 29;
        }
        if (c < C) { a = (a ^ (uint64_t)(uint32_t)row[c]) * 0x100000001B3ull; b = (b + (uint64_t)(uint32_t)row[c]) * 0xFF51AFD7ED558CCDull; }
    }
    *h1 = a; *h2 = b;
}

/* a context for (binMat rows nodePositions, lenSamples), from the cache or new; release with legacy_release */
int legacy_ctx(scb_ctx **out, int N, int C, const int32_t *binMat, const int32_t *nodePositions)
{
    *out = nullptr;
    if (!binMat || !nodePositions) return fail(SCB_ERR_INVALID, "null binMat / nodePositions");
    for (int i = 0; i < N; ++i) if (nodePositions[i] < 0) return fail(SCB_ERR_INVALID, "nodePositions[%d] is negative", i);
    const int device = legacy_device();
    uint64_t h1, h2;
    legacy_hash(binMat, g_legacyCell, nodePositions, N, C, &h1, &h2);
    {
        std::lock_guard<std::mutex> lk(g_legacyMu);
        for (auto &e : g_legacyCache)
            if (e.ctx && e.h1 == h1 && e.h2 == h2 && e.N == N && e.C == C && e.device == device) {
                ++e.inUse; e.stamp = ++g_legacyClock;
                *out = e.ctx;
                return SCB_OK;
            }
    }
    scb_ctx *ctx = nullptr;
    int rc = scb_ctx_create(&ctx, device, N, C, binMat, 4, g_legacyCell, nodePositions);
    if (rc) return rc;
    scb_ctx *evict = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_legacyMu);
        LegacyEntry *slot = nullptr;
        for (auto &e : g_legacyCache) if (!e.ctx) { slot = &e; break; }
        if (!slot)
            for (auto &e : g_legacyCache) if (e.inUse == 0 && (!slot || e.stamp < slot->stamp)) slot = &e;
        if (slot) {
            evict = slot->ctx;
            slot->h1 = h1; slot->h2 = h2; slot->N = N; slot->C = C; slot->device = device; slot->ctx = ctx;
            slot->inUse = 1; slot->stamp = ++g_legacyClock;
        }
        /* every slot busy: the context stays private to this call and is destroyed on release */
    }
    if (evict) scb_ctx_destroy(evict);
    *out = ctx;
    return SCB_OK;
}

void legacy_release(scb_ctx *ctx)
{
    if (!ctx) return;
    {
        std::lock_guard<std::mutex> lk(g_legacyMu);
        for (auto &e : g_legacyCache) if (e.ctx == ctx) { --e.inUse; return; }
    }
    scb_ctx_destroy(ctx);
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
    if (C <= 0 || N <= 0) return;                       /* the reference's loops do not run either */
    int rc = SCB_OK;
    scb_ctx *ctx = nullptr;
    if ((int)simSteps != SCB_STEPS)
        rc = fail(SCB_ERR_UNSUPPORTED, "simSteps must be 100 (getAttractCycle2 always scans 100 rows, simulator.c:306-331), got %d", (int)simSteps);
    else if (C > g_legacyCell)
        rc = fail(SCB_ERR_INVALID, "lenSamples %d exceeds the binMat row stride %d (SCB_LEGACY_CELL / scb_set_legacy_geometry)", C, g_legacyCell);
    else rc = legacy_ctx(&ctx, N, C, binMat, nodePositions);
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
    legacy_release(ctx);
    if (rc != SCB_OK) {
        legacy_fatal("scSyncBool", scb_last_error());
        /* sentinel mode: no evaluation can reach these (an error is at most nodeNum per cell, lenSamples per node);
         * the sum over the array still fits an int64 */
        const int n = localSearch ? N : C;
        for (int i = 0; i < n; ++i) errors[i] = 0x7fffffff / (n > 0 ? n : 1);
    }
}

/* trajectories + attractor windows of the uploaded single individual (P == 1 already on the device).
 * recOn / recBits (host, may be null): nodes whose RECORDED rows of steps >= 1 are overridden, see ScbTrajArgs */
static int run_trajectories(scb_ctx *c, int sem, int steps, int maxStates, const unsigned char *recOn = nullptr,
                            const uint32_t *recBits = nullptr)
{
    const int N = c->N, tiles = (c->W + 31) / 32, NW = (N + 31) / 32;
    int rc;
    const int warps = std::min(16, std::max(1, (N + 3) / 4));
    if (scb_smem_bytes(N, 1, 2, warps) > c->smemLimit) return fail(SCB_ERR_UNSUPPORTED, "nodeNum %d too large for the trajectory kernel", N);
    if ((rc = c->dTraj.ensure((size_t)tiles * steps * N * 32))) return rc;
    if ((rc = c->dRes.ensure((size_t)c->C * 2)) || (rc = c->dNStates.ensure(c->C))) return rc;
    if ((rc = c->dStates.ensure((size_t)c->C * std::max(maxStates, 1) * NW))) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->dDiag.p, 0, sizeof(ScbDiag), st));
    if ((rc = launch_compile(c, sem, c->dKo.p))) return rc;
    ScbTrajArgs ta;
    ta.X = c->dX.p; ta.N = N; ta.C = c->C; ta.Wpad = c->Wpad; ta.tiles = tiles; ta.steps = steps;
    ta.descs = c->dDescs.p; ta.counts = c->dCounts.p; ta.traj = c->dTraj.p; ta.recOn = nullptr; ta.recBits = nullptr;
    if (recOn) {
        if ((rc = c->dRecOn.ensure(N)) || (rc = c->dRecBits.ensure((size_t)N * 4))) return rc;
        CK(cudaMemcpyAsync(c->dRecOn.p, recOn, (size_t)N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->dRecBits.p, recBits, (size_t)N * 16, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                /* the host arrays are the caller's temporaries */
        ta.recOn = c->dRecOn.p; ta.recBits = c->dRecBits.p;
    }
    const unsigned grid = (unsigned)std::max(1, std::min(tiles, c->numSM * 2));
    SCB_LAUNCH(scb_k_traj, dim3(grid), dim3((unsigned)warps * 32), scb_smem_bytes(N, 1, 2, warps), st, ta);
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
    std::lock_guard<std::recursive_mutex> lk(c->mu);        /* upload + kernels + copies under one lock */
    rc = scb_pop_upload(c, 1, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0, knockouts, knockins);
    if (rc) return rc;
    const int N = c->N, NW = (N + 31) / 32;
    std::vector<unsigned char> recOn;
    std::vector<uint32_t> recBits;
    if (semantics == 1) {
        /* singleCell.py:1037-1038: the last node falls through without being updated or recorded: its rows stay 0 */
        const int al = andLenList[N - 1];
        if (knockouts[N - 1] != 1 && knockins[N - 1] != 1 && al != 1 && al != 0) {
            recOn.assign((size_t)N, 0); recBits.assign((size_t)N * 4, 0u);
            recOn[N - 1] = 1;
        }
    }
    rc = run_trajectories(c, semantics == 1 ? SCB_SEM_PY : SCB_SEM_C, semantics == 1 ? 10 : SCB_STEPS, maxStates,
                          recOn.empty() ? nullptr : recOn.data(), recOn.empty() ? nullptr : recBits.data());
    if (rc) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(res, c->dRes.p, (size_t)c->C * 2 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nStates, c->dNStates.p, (size_t)c->C * 4, cudaMemcpyDeviceToHost, st));
    if (maxStates > 0) CK(cudaMemcpyAsync(states, c->dStates.p, (size_t)c->C * maxStates * NW * 4, cudaMemcpyDeviceToHost, st));
    return check_compile_diag(c);
}

int scb_attractor_states(scb_ctx *c, const int32_t *individual, int indLen, const int32_t *andLenList,
                         const int32_t *individualParse, const int32_t *andNodes, const int32_t *andNodeInvertList,
                         const int32_t *knockouts, const int32_t *knockins, int semantics, int maxStates,
                         uint32_t *unique, int cap, int *nUnique)
{
    int rc = check_ctx(c);
    if (rc) return rc;
    if (semantics != 0 && semantics != 1) return fail(SCB_ERR_INVALID, "semantics must be 0 (C cluster) or 1 (Python __cluster)");
    if (maxStates <= 0 || cap < 0 || !nUnique || (cap > 0 && !unique)) return fail(SCB_ERR_INVALID, "bad argument");
    std::lock_guard<std::recursive_mutex> lk(c->mu);
    rc = scb_pop_upload(c, 1, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0, knockouts, knockins);
    if (rc) return rc;
    const int N = c->N, NW = (N + 31) / 32;
    std::vector<unsigned char> recOn;
    std::vector<uint32_t> recBits;
    if (semantics == 1) {
        const int al = andLenList[N - 1];                   /* singleCell.py:1037-1038, as in scb_attractors */
        if (knockouts[N - 1] != 1 && knockins[N - 1] != 1 && al != 1 && al != 0) {
            recOn.assign((size_t)N, 0); recBits.assign((size_t)N * 4, 0u);
            recOn[N - 1] = 1;
        }
    }
    rc = run_trajectories(c, semantics == 1 ? SCB_SEM_PY : SCB_SEM_C, semantics == 1 ? 10 : SCB_STEPS, maxStates,
                          recOn.empty() ? nullptr : recOn.data(), recOn.empty() ? nullptr : recBits.data());
    if (rc) return rc;
    const long long rows = (long long)c->C * maxStates;
    size_t tsize = 1024;
    while (tsize < (size_t)rows * 2) tsize <<= 1;
    if ((rc = c->dDedupeTable.ensure(tsize)) || (rc = c->dDedupeFlags.ensure((size_t)rows)) || (rc = c->dDedupeOut.ensure((size_t)std::max(cap, 1) * NW))) return rc;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->dDedupeTable.p, 0xff, tsize * 4, st));
    CK(cudaMemsetAsync(c->dDedupeFlags.p, 0, (size_t)rows, st));
    ScbDedupeArgs da;
    da.states = c->dStates.p; da.nStates = c->dNStates.p; da.C = c->C; da.maxStates = maxStates; da.NW = NW;
    da.table = c->dDedupeTable.p; da.mask = (unsigned)(tsize - 1); da.flags = c->dDedupeFlags.p;
    da.count = c->dCounters.p + 11; da.out = c->dDedupeOut.p; da.cap = cap;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    SCB_LAUNCH(scb_k_dedupe_insert, dim3(grid), dim3(256), 0, st, da);
    SCB_LAUNCH(scb_k_dedupe_flag, dim3(grid), dim3(256), 0, st, da);
    const int cthr = c->optMaxCtas > 0 ? 64 : 1024;         /* (the emulator runs a thread per CUDA thread) */
    SCB_LAUNCH(scb_k_dedupe_compact, dim3(1), dim3((unsigned)cthr), (size_t)(cthr + 1) * 4, st, da);
    CK(cudaGetLastError());
    unsigned int n = 0;
    CK(cudaMemcpyAsync(&n, c->dCounters.p + 11, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *nUnique = (int)n;
    const size_t take = std::min<size_t>(n, (size_t)cap);
    if (take) CK(cudaMemcpyAsync(unique, c->dDedupeOut.p, take * NW * 4, cudaMemcpyDeviceToHost, st));
    return check_compile_diag(c);
}

/* simulator.c:487-552.  One cell (`sampleIndex`); resSubmit = attractor window; the scratch `simData[step][column]`
 * (row stride = legacy NODE) is left exactly as the reference leaves it.  That includes the slip of its knock-out
 * branch, which records a knocked node i in column i instead of column nodePositions[i] (simulator.c:518): the
 * knocked node's own column keeps what the caller's scratch held (or the 0 another knocked node wrote there), and
 * column i may overwrite the node recorded there when i comes later in the loop.  The window search runs over the
 * columns as recorded, so the kernel is told which nodes' recorded rows differ from their true rows (the dynamics
 * never do).  The caller's scratch must hold 0/1 where it shows through (the reference's Python passes zeros). */
void cluster(void *simData, int32_t *resSubmit, intptr_t sampleIndex, const int32_t *individual, intptr_t indLen,
             intptr_t nodeNum, const int32_t *andLenList, const int32_t *individualParse, const int32_t *andNodes,
             const int32_t *andNodeInvertList, intptr_t simSteps, const int32_t *knockouts, const int32_t *knockins,
             const int32_t *binMat, const int32_t *nodePositions)
{
    legacy_env();
    const int N = (int)nodeNum;
    if (!resSubmit) legacy_die("cluster", "null resSubmit");
    /* sentinel mode: a window no trajectory has */
#define SCB_CLUSTER_FAIL(msg) do { legacy_fatal("cluster", msg); resSubmit[0] = resSubmit[1] = -0x7fffffff; return; } while (0)
    if ((int)simSteps != SCB_STEPS) SCB_CLUSTER_FAIL("simSteps must be 100 (simulator.c:306-331 always scans 100 rows)");
    if (N <= 0) SCB_CLUSTER_FAIL("nodeNum must be positive");
    int32_t *sd = static_cast<int32_t *>(simData);
    std::vector<unsigned char> recOn((size_t)N, 0);
    std::vector<uint32_t> recBits((size_t)N * 4, 0u);
    bool anyRec = false;
    for (int k = 0; k < N; ++k) {
        const int col = nodePositions[k];
        const bool koK = knockouts[k] == 1, koCol = col < N && knockouts[col] == 1;
        if (!koK && !(koCol && col > k)) continue;            /* the true trajectory is what the column shows */
        recOn[k] = 1; anyRec = true;
        if (koCol) continue;                                  /* a knocked node's 0 is the last write: bits stay 0 */
        for (int t = 1; t < SCB_STEPS && sd; ++t) {           /* nobody writes the column: the caller's scratch shows through */
            const int32_t v = sd[(long long)t * g_legacyNode + col];
            if (v != 0 && v != 1) SCB_CLUSTER_FAIL("the scratch holds a value other than 0/1 in a column the knock-out branch leaves unwritten");
            if (v) recBits[(size_t)k * 4 + (t >> 5)] |= 1u << (t & 31);
        }
    }
    scb_ctx *ctx = nullptr;
    int rc = legacy_ctx(&ctx, N, 1, binMat + (long long)sampleIndex, nodePositions);
    std::vector<uint32_t> traj((size_t)SCB_STEPS * N * 32);
    int32_t r2[2] = {0, 0};
    if (rc == SCB_OK)
        rc = scb_pop_upload(ctx, 1, individual, (int)indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                            knockouts, knockins);
    if (rc == SCB_OK) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        rc = run_trajectories(ctx, SCB_SEM_C, SCB_STEPS, 0, anyRec ? recOn.data() : nullptr, anyRec ? recBits.data() : nullptr);
        if (rc == SCB_OK) {
            cudaMemcpyAsync(traj.data(), ctx->dTraj.p, traj.size() * 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(r2, ctx->dRes.p, 8, cudaMemcpyDeviceToHost, ctx->stream);
            rc = check_compile_diag(ctx);
        }
    }
    legacy_release(ctx);
    if (rc != SCB_OK) SCB_CLUSTER_FAIL(scb_last_error());
#undef SCB_CLUSTER_FAIL
    if (sd) {
        for (int t = 0; t < SCB_STEPS; ++t) {
            if (t > 0)
                for (int i = 0; i < N; ++i)
                    if (knockouts[i] == 1) sd[(long long)t * g_legacyNode + i] = 0;              /* :518 */
            for (int i = 0; i < N; ++i)                                                         /* the columns as recorded */
                sd[(long long)t * g_legacyNode + nodePositions[i]] = (int32_t)(traj[((size_t)t * N + i) * 32] & 1u);
        }
    }
    resSubmit[0] = r2[0]; resSubmit[1] = r2[1];
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
    std::lock_guard<std::recursive_mutex> lkAll(ctx->mu);   /* upload + run + fetch under one lock */
    rc = pop_upload_impl(ctx, P, 1, individual, indLen, andLenList, individualParse, andNodes, andNodeInvertList, 0,
                         P, knock.data(), knockins);
    if (rc) return rc;
    if ((rc = pop_run_impl(ctx, SCB_MODE_IMPORTANCE, SCB_SEM_IMPORTANCE))) return rc;
    std::vector<int32_t> cnt((size_t)P * N);
    std::vector<unsigned long long> noWindow((size_t)P);
    {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
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
    if (N <= 0 || C <= 0) return;
    scb_ctx *ctx = nullptr;
    int rc;
    if ((int)simSteps != SCB_STEPS) rc = fail(SCB_ERR_UNSUPPORTED, "simSteps must be 100, got %d", (int)simSteps);
    else if (C > g_legacyCell) rc = fail(SCB_ERR_INVALID, "lenSamples %d exceeds the binMat row stride %d", C, g_legacyCell);
    else rc = legacy_ctx(&ctx, N, C, binMat, nodePositions);
    if (rc == SCB_OK)
        rc = scb_importance_score(ctx, individual, (int)indLen, andLenList, individualParse, andNodes, andNodeInvertList,
                                  knockouts, knockins, importanceScores);
    legacy_release(ctx);
    /* a cell without an attractor window: the reference divides by zero (SIGFPE); here the score is NaN, which no
     * caller can mistake for a result */
    if (rc != SCB_OK && !(rc == SCB_ERR_INVALID && importanceScores[0] != importanceScores[0])) {
        legacy_fatal("importanceScore", scb_last_error());
        importanceScores[0] = nan("");
    }
}

}  /* extern "C" */
