// cl2_driver.cu -- the per-file work unit of callLoops in one native call: the whole eps x minPts sweep of clustering
// runs, the per-run candidate filters, filterLoopsByDis, estLoopSig and the de-duplication of combineLoops /
// filterDupLoops, for one chromosome (pair) resident in HBM.  This is what the reference hands to a joblib worker
// (callCisLoops.py:134-137 + 561-572, callTransLoops.py:95-97 + 237-243); keeping the loop over the runs on this
// side of the ABI means the host language only sees one call per file.
#include "cl2_points.cuh"
#include <algorithm>
#include <chrono>
#include <new>
#include <unordered_map>

struct cl2_loops {
    int npass = 0;
    std::vector<int64_t> pass_cand, pass_pets;   // per run, in the order given: candidates kept, PETs in them
    int64_t n_tested = 0;                        // candidates that went into estLoopSig (after filterLoopsByDis)
    int64_t n_unique = -1;                       // distinct coordinates among them (only when asked for)
    int64_t n = 0;                               // survivors, de-duplicated, in candidate order
    int64_t n_pets = 0;
    int64_t ext[4] = {0, 0, 0, 0};
    std::vector<int64_t> coords, core;
    std::vector<int32_t> pass_id;
    std::vector<double> hyp, stats;
    double seconds[6] = {0, 0, 0, 0, 0, 0};      // clustering runs, index build, tests: host wall time, then thread CPU time
};

static inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#include <time.h>
static inline double cpu_s() {              // CPU time of the calling thread (waits that spin count, waits that sleep do not)
    struct timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int64_t gcd_i64(int64_t a, int64_t b) {
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}
// may run (e, m2) skip the rows in noise cells removed by run (E, m)?  (cl2_points_sweep_hint)
static bool rows_carry_over(int64_t E, int32_t m, int64_t e, int32_t m2) {
    return e < E && m2 >= m && E >= 3 * e - gcd_i64(e, E);
}

struct Key4 {
    int64_t v[4];
    bool operator==(const Key4 &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
};
struct Key4Hash {
    size_t operator()(const Key4 &k) const {
        uint64_t h = (uint64_t)k.v[0] * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)k.v[1] * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.v[2] * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.v[3] * 0x27D4EB2F165667C5ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

// ---- device-resident candidate list ------------------------------------------------------------------------------
// per cluster box of one run: the per-run candidate filters (cis: callCisLoops.py:97-99, :114, :116; trans:
// callTransLoops.py:62-64), the run's counters as runCisDBSCANLoops logs them (:122), then filterLoopsByDis (distance
// > cut, callCisLoops.py:146-156; centres as the reference computes them, (start + end) / 2 in floating point,
// :107-112) and the append
__global__ void __launch_bounds__(256) k_cand_emit(const int *__restrict__ boxes, const uint64_t *__restrict__ keys, int32_t ncomp,
                                                    int mode, long long cut, int pass, longlong4 *__restrict__ cand,
                                                    uint64_t *__restrict__ key, int32_t *__restrict__ pid,
                                                    int32_t *__restrict__ cursor, long long *__restrict__ pass_stat) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool cnt = false, keep = false;
    long long x0 = 0, x1 = 0, y0 = 0, y1 = 0, size = 0;
    if (k < ncomp) {
        x0 = boxes[5 * (size_t)k]; x1 = boxes[5 * (size_t)k + 1]; y0 = boxes[5 * (size_t)k + 2]; y1 = boxes[5 * (size_t)k + 3];
        size = (unsigned)boxes[5 * (size_t)k + 4];
        cnt = !(x0 == x1 || y0 == y1);
        if (cnt && mode == CL2_MODE_CIS) {
            if ((x1 - x0) + (y1 - y0) < 200) cnt = false;
            if (!(x1 < y0)) cnt = false;
        }
        keep = cnt;
        if (keep && mode == CL2_MODE_CIS) {
            const double xc = (double)(x0 + x1) / 2.0, yc = (double)(y0 + y1) / 2.0;
            const double d = yc - xc;
            keep = (d < 0 ? -d : d) > (double)cut;
        }
    }
    // the run's counters: one atomic pair per warp
    const unsigned mc = __ballot_sync(0xffffffffu, cnt);
    long long sz = cnt ? size : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) sz += __shfl_xor_sync(0xffffffffu, sz, o);
    if ((threadIdx.x & 31) == 0 && mc) {
        atomicAdd((unsigned long long *)&pass_stat[2 * pass], (unsigned long long)__popc(mc));
        atomicAdd((unsigned long long *)&pass_stat[2 * pass + 1], (unsigned long long)sz);
    }
    // the append: one cursor bump per warp
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && mk) base = atomicAdd(cursor, __popc(mk));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const int slot = base + __popc(mk & ((1u << (threadIdx.x & 31)) - 1u));
        cand[slot] = make_longlong4(x0, x1, y0, y1);
        key[slot] = keys ? keys[k] : (uint64_t)k;
        pid[slot] = pass;
    }
}

int cl2_sink_emit(cl2_ctx *ctx, cl2_cand_sink *sink, const int *boxes, const uint64_t *keys, int32_t ncomp) {
    if (ncomp <= 0) return CL2_OK;
    if (sink->bound + ncomp > sink->cap) {
        // grow (the rows used so far are at most `bound`)
        int64_t cap = sink->cap ? sink->cap : 1 << 16;
        while (cap < sink->bound + ncomp) cap *= 2;
        longlong4 *c2 = nullptr;
        uint64_t *k2 = nullptr;
        int32_t *p2 = nullptr;
        CL2_TRY(cl2_dev_alloc(ctx, &c2, (size_t)cap));
        CL2_TRY(cl2_dev_alloc(ctx, &k2, (size_t)cap));
        CL2_TRY(cl2_dev_alloc(ctx, &p2, (size_t)cap));
        if (sink->bound > 0) {
            CL2_CUDA(ctx, cudaMemcpyAsync(c2, sink->cand, (size_t)sink->bound * 32, cudaMemcpyDeviceToDevice, ctx->stream));
            CL2_CUDA(ctx, cudaMemcpyAsync(k2, sink->key, (size_t)sink->bound * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CL2_CUDA(ctx, cudaMemcpyAsync(p2, sink->pid, (size_t)sink->bound * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        cl2_dev_free(ctx, sink->cand);
        cl2_dev_free(ctx, sink->key);
        cl2_dev_free(ctx, sink->pid);
        sink->cand = c2;
        sink->key = k2;
        sink->pid = p2;
        sink->cap = cap;
    }
    k_cand_emit<<<cl2_grid(ncomp, 256), 256, 0, ctx->stream>>>(boxes, keys, ncomp, sink->mode, (long long)sink->cut, sink->pass,
                                                               sink->cand, sink->key, sink->pid, sink->cursor, sink->pass_stat);
    ctx->launches++;
    sink->bound += ncomp;
    return CL2_OK;
}

void cl2_sink_free(cl2_ctx *ctx, cl2_cand_sink *sink) {
    cl2_dev_free(ctx, sink->cand);
    cl2_dev_free(ctx, sink->key);
    cl2_dev_free(ctx, sink->pid);
    cl2_dev_free(ctx, sink->cursor);
    *sink = cl2_cand_sink();
}

static int call_loops_body(cl2_ctx *ctx, cl2_points *pts, int32_t mode, int32_t hic, int32_t npass, const int64_t *eps,
                           const int32_t *minPts, int64_t cut, int32_t min_pts_test, int32_t win, double countDiffCut,
                           double pseudo, double peakPcut, int32_t want_unique, cl2_loops *r, cl2_cand_sink &sink) {
    const bool cis = mode == CL2_MODE_CIS;
    const bool use_c2 = cis && hic;                          // trans always clusters with blockDBSCAN (callTransLoops.py:29)
    double t0 = now_s(), c0 = cpu_s();
    // cursor + the runs' counters in one small device block
    CL2_TRY(cl2_dev_alloc(ctx, &sink.cursor, (size_t)4 * ((size_t)npass + 1)));
    CL2_CUDA(ctx, cudaMemsetAsync(sink.cursor, 0, (size_t)16 * ((size_t)npass + 1), ctx->stream));
    sink.pass_stat = reinterpret_cast<long long *>(sink.cursor + 4);
    sink.mode = mode;
    sink.cut = cut;

    // Order of execution.  The result of a run does not depend on it; the blockDBSCAN runs go from the largest eps down
    // (smallest minPts first) so that each tells the later ones which rows sit in noise cells it removed.
    std::vector<int> order((size_t)npass);
    for (int i = 0; i < npass; i++) order[i] = i;
    bool hint = false;
    if (!use_c2)
        for (int i = 1; i < npass; i++)
            if (eps[i] != eps[0]) hint = true;
    if (hint)
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (eps[a] != eps[b]) return eps[a] > eps[b];
            return minPts[a] < minPts[b];
        });
    for (int pos = 0; pos < npass; pos++) {
        const int i = order[pos];
        if (hint) {
            bool useful = false;                             // remembering rows costs a pass over the points
            for (int q = pos + 1; q < npass; q++)
                if (rows_carry_over(eps[i], minPts[i], eps[order[q]], minPts[order[q]])) useful = true;
            cl2_points_sweep_hint(ctx, pts, useful ? 1 : 2);
        }
        int32_t ncl = 0;
        sink.pass = i;
        // the clusters of the run go straight into the device-resident list (blockDBSCAN: unordered, each with the key
        // that ranks it in the reference's numbering; only the few survivors are put in order, at the end)
        CL2_TRY(use_c2 ? cl2_cdbscan2_impl(ctx, pts, eps[i], minPts[i], nullptr, &ncl, &sink)
                       : cl2_block_dbscan_impl(ctx, pts, eps[i], minPts[i], nullptr, &ncl, false, &sink));
    }
    // one read-back for all runs: the rows in the list and the runs' counters
    std::vector<long long> hs((size_t)2 * (size_t)npass + 2);
    CL2_CUDA(ctx, cudaMemcpyAsync(hs.data(), sink.cursor, hs.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += (int64_t)hs.size() * 8;
    const int64_t L = (int64_t)(int32_t)(hs[0] & 0xffffffffll);
    for (int i = 0; i < npass; i++) {
        r->pass_cand[(size_t)i] = hs[2 + 2 * (size_t)i];
        r->pass_pets[(size_t)i] = hs[3 + 2 * (size_t)i];
    }
    r->n_tested = L;
    if (want_unique && L > 0) {
        std::vector<int64_t> all((size_t)L * 4);
        CL2_CUDA(ctx, cudaMemcpyAsync(all.data(), sink.cand, (size_t)L * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2h_bytes += L * 32;
        std::unordered_map<Key4, int, Key4Hash> seen;
        seen.reserve((size_t)L * 2);
        for (int64_t k = 0; k < L; k++) {
            Key4 key = {{all[4 * k], all[4 * k + 1], all[4 * k + 2], all[4 * k + 3]}};
            seen.emplace(key, 0);
        }
        r->n_unique = (int64_t)seen.size();
    }
    double t1 = now_s(), c1 = cpu_s();
    r->seconds[0] = t1 - t0;
    r->seconds[3] = c1 - c0;
    if (L == 0) {
        cl2_points_drop_cache(ctx, pts);
        return CL2_OK;
    }
    cl2_index *idx = nullptr;
    CL2_TRY(cl2_index_build(ctx, pts, &idx));
    double t2 = now_s(), c2 = cpu_s();
    r->seconds[1] = t2 - t1;
    r->seconds[4] = c2 - c1;
    cl2_est_out eo;
    const double rpb = (double)pts->n / (double)(pts->ext[3] - pts->ext[0]);       // callCisLoops.py:231
    int rc = cl2_est_loops_device(ctx, idx, sink.cand, L, win, mode, min_pts_test, hic, countDiffCut, pseudo, peakPcut, rpb,
                                  sink.pid, sink.key, true, eo);
    cl2_index_free(ctx, idx);
    cl2_points_drop_cache(ctx, pts);                          // eps grid + y-order: derived data, rebuilt by the next call
    if (rc != CL2_OK) return rc;
    const int64_t nk = (int64_t)eo.sel.size();
    // survivors in candidate order: run by run as given, ascending cluster id inside a run
    std::vector<int64_t> sorder((size_t)nk);
    for (int64_t q = 0; q < nk; q++) sorder[(size_t)q] = q;
    std::sort(sorder.begin(), sorder.end(), [&](int64_t a, int64_t b) {
        if (eo.pid[(size_t)a] != eo.pid[(size_t)b]) return eo.pid[(size_t)a] < eo.pid[(size_t)b];
        if (eo.key[(size_t)a] != eo.key[(size_t)b]) return eo.key[(size_t)a] < eo.key[(size_t)b];
        return eo.sel[(size_t)a] < eo.sel[(size_t)b];
    });
    // combineLoops (geo.py:90-112) drops a candidate whose coordinates were already present from an EARLIER run;
    // the cis path then also applies filterDupLoops (callCisLoops.py:159-169: first position of every coordinate
    // tuple).  A candidate's test result depends on its coordinates only, so both are applied to the survivors.
    std::unordered_map<Key4, int32_t, Key4Hash> first_pass;
    first_pass.reserve((size_t)nk * 2 + 1);
    for (int64_t qq = 0; qq < nk; qq++) {
        const size_t q = (size_t)sorder[(size_t)qq];
        Key4 key = {{eo.coords[4 * q], eo.coords[4 * q + 1], eo.coords[4 * q + 2], eo.coords[4 * q + 3]}};
        auto it = first_pass.find(key);
        if (it != first_pass.end()) {
            if (cis || it->second != eo.pid[q]) continue;
        } else {
            first_pass.emplace(key, eo.pid[q]);
        }
        r->coords.insert(r->coords.end(), eo.coords.begin() + 4 * q, eo.coords.begin() + 4 * q + 4);
        r->pass_id.push_back(eo.pid[q]);
        r->core.insert(r->core.end(), eo.core.begin() + 4 * q, eo.core.begin() + 4 * q + 4);
        r->hyp.push_back(eo.hyp[q]);
        r->stats.insert(r->stats.end(), eo.stats.begin() + 13 * q, eo.stats.begin() + 13 * q + 13);
    }
    r->n = (int64_t)r->pass_id.size();
    r->seconds[2] = now_s() - t2;
    r->seconds[5] = cpu_s() - c2;
    return CL2_OK;
}

extern "C" int cl2_call_loops(cl2_ctx *ctx, cl2_points *pts, int32_t mode, int32_t hic, int32_t npass, const int64_t *eps,
                              const int32_t *minPts, int64_t cut, int32_t min_pts_test, int32_t win, double countDiffCut,
                              double pseudo, double peakPcut, int32_t want_unique, cl2_loops **out) {
    if (!ctx || !pts || !out || npass < 0 || (npass > 0 && (!eps || !minPts)))
        return cl2_fail(ctx, CL2_ERR_ARG, "cl2_call_loops: bad argument");
    if (mode != CL2_MODE_CIS && mode != CL2_MODE_TRANS) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_call_loops: bad mode");
    *out = nullptr;
    cl2_loops *r = new (std::nothrow) cl2_loops();
    if (!r) return CL2_ERR_NOMEM;
    r->npass = npass;
    r->pass_cand.assign((size_t)npass, 0);
    r->pass_pets.assign((size_t)npass, 0);
    r->n_pets = pts->n;
    for (int i = 0; i < 4; i++) r->ext[i] = pts->ext[i];
    *out = r;
    if (pts->n == 0 || npass == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    cl2_cand_sink sink;
    pts->want_xorder = 1;              // the first blockDBSCAN grid over every row also leaves the x-order for the index
    const int rc = call_loops_body(ctx, pts, mode, hic, npass, eps, minPts, cut, min_pts_test, win, countDiffCut, pseudo,
                                   peakPcut, want_unique, r, sink);
    pts->want_xorder = 0;
    cl2_sink_free(ctx, &sink);
    return rc;
}

extern "C" int64_t cl2_loops_count(const cl2_loops *r) { return r ? r->n : -1; }

extern "C" int cl2_loops_info(const cl2_loops *r, int64_t *pass_cand, int64_t *pass_pets, int64_t *totals, double *seconds) {
    if (!r) return CL2_ERR_ARG;
    for (int i = 0; i < r->npass; i++) {
        if (pass_cand) pass_cand[i] = r->pass_cand[(size_t)i];
        if (pass_pets) pass_pets[i] = r->pass_pets[(size_t)i];
    }
    if (totals) {
        totals[0] = r->n_pets;
        totals[1] = r->n_tested;
        totals[2] = r->n_unique;
        totals[3] = r->n;
        for (int i = 0; i < 4; i++) totals[4 + i] = r->ext[i];
    }
    if (seconds)
        for (int i = 0; i < 6; i++) seconds[i] = r->seconds[i];
    return CL2_OK;
}

extern "C" int cl2_loops_read(const cl2_loops *r, int64_t *coords, int32_t *pass_id, int64_t *core, double *hyp,
                              double *stats) {
    if (!r) return CL2_ERR_ARG;
    const size_t n = (size_t)r->n;
    if (n == 0) return CL2_OK;
    if (coords) memcpy(coords, r->coords.data(), n * 32);
    if (pass_id) memcpy(pass_id, r->pass_id.data(), n * 4);
    if (core) memcpy(core, r->core.data(), n * 32);
    if (hyp) memcpy(hyp, r->hyp.data(), n * 8);
    if (stats) memcpy(stats, r->stats.data(), n * 13 * 8);
    return CL2_OK;
}

extern "C" void cl2_loops_free(cl2_loops *r) { delete r; }
