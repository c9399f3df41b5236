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
    double seconds[3] = {0, 0, 0};               // clustering runs, index build, tests
};

static inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
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
    const bool cis = mode == CL2_MODE_CIS;
    const bool use_c2 = cis && hic;                          // trans always clusters with blockDBSCAN (callTransLoops.py:29)
    double t0 = now_s();

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
    // per run: kept boxes, 4 values each, and the key that ranks each box in the reference's cluster numbering (the
    // blockDBSCAN runs report their clusters unordered: only the few survivors are put in order, at the end)
    std::vector<std::vector<int64_t>> cand((size_t)npass);
    std::vector<std::vector<uint64_t>> ckey((size_t)npass);
    int rc = CL2_OK;
    for (int pos = 0; pos < npass && rc == CL2_OK; pos++) {
        const int i = order[pos];
        if (hint) {
            bool useful = false;                             // remembering rows costs a pass over the points
            for (int q = pos + 1; q < npass; q++)
                if (rows_carry_over(eps[i], minPts[i], eps[order[q]], minPts[order[q]])) useful = true;
            cl2_points_sweep_hint(ctx, pts, useful ? 1 : 2);
        }
        int32_t ncl = 0;
        rc = use_c2 ? cl2_cdbscan2(ctx, pts, eps[i], minPts[i], nullptr, &ncl)
                    : cl2_block_dbscan_impl(ctx, pts, eps[i], minPts[i], nullptr, &ncl, false);
        if (rc != CL2_OK) break;
        std::vector<int64_t> &c = cand[(size_t)i];
        std::vector<uint64_t> &ck = ckey[(size_t)i];
        c.reserve((size_t)ncl * 4);
        ck.reserve((size_t)ncl);
        int64_t pets = 0;
        for (int32_t k = 0; k < ncl; k++) {
            const int64_t x0 = pts->bbox[4 * (size_t)k], x1 = pts->bbox[4 * (size_t)k + 1];
            const int64_t y0 = pts->bbox[4 * (size_t)k + 2], y1 = pts->bbox[4 * (size_t)k + 3];
            if (x0 == x1 || y0 == y1) continue;              // callCisLoops.py:97-99 / callTransLoops.py:62-64
            if (cis) {
                if ((x1 - x0) + (y1 - y0) < 200) continue;   // :114
                if (!(x1 < y0)) continue;                    // :116
            }
            c.push_back(x0); c.push_back(x1); c.push_back(y0); c.push_back(y1);
            ck.push_back(use_c2 ? (uint64_t)k : pts->ckeys[(size_t)k]);
            pets += pts->size[(size_t)k];
        }
        r->pass_cand[(size_t)i] = (int64_t)c.size() / 4;
        r->pass_pets[(size_t)i] = pets;
    }
    if (rc != CL2_OK) return rc;
    // candidates of all runs in sweep order; filterLoopsByDis (callCisLoops.py:146-156) keeps distance > cut, with the
    // centres as the reference computes them, (start + end) / 2 in floating point (:107-112)
    std::vector<int64_t> all;
    std::vector<int32_t> pid;
    std::vector<uint64_t> akey;
    for (int i = 0; i < npass; i++) {
        const std::vector<int64_t> &c = cand[(size_t)i];
        for (size_t k = 0; k + 3 < c.size(); k += 4) {
            if (cis) {
                const double xc = (double)(c[k] + c[k + 1]) / 2.0, yc = (double)(c[k + 2] + c[k + 3]) / 2.0;
                const double d = yc - xc;
                if (!((d < 0 ? -d : d) > (double)cut)) continue;
            }
            all.insert(all.end(), c.begin() + k, c.begin() + k + 4);
            pid.push_back(i);
            akey.push_back(ckey[(size_t)i][k / 4]);
        }
    }
    const int64_t L = (int64_t)pid.size();
    r->n_tested = L;
    if (want_unique) {
        std::unordered_map<Key4, int, Key4Hash> seen;
        seen.reserve((size_t)L * 2);
        for (int64_t k = 0; k < L; k++) {
            Key4 key = {{all[4 * k], all[4 * k + 1], all[4 * k + 2], all[4 * k + 3]}};
            seen.emplace(key, 0);
        }
        r->n_unique = (int64_t)seen.size();
    }
    double t1 = now_s();
    r->seconds[0] = t1 - t0;
    if (L == 0) {
        cl2_points_drop_cache(ctx, pts);
        return CL2_OK;
    }
    cl2_index *idx = nullptr;
    rc = cl2_index_build(ctx, pts, &idx);
    if (rc != CL2_OK) return rc;
    double t2 = now_s();
    r->seconds[1] = t2 - t1;
    std::vector<int64_t> sel((size_t)L), core((size_t)L * 4);
    std::vector<double> hyp((size_t)L), stats((size_t)L * 13);
    int64_t nk = 0;
    const double rpb = (double)pts->n / (double)(pts->ext[3] - pts->ext[0]);       // callCisLoops.py:231
    rc = cl2_est_loops(ctx, idx, all.data(), L, win, mode, min_pts_test, hic, countDiffCut, pseudo, peakPcut, rpb, &nk,
                       sel.data(), core.data(), hyp.data(), stats.data());
    cl2_index_free(ctx, idx);
    cl2_points_drop_cache(ctx, pts);                          // eps grid + y-order: derived data, rebuilt by the next call
    if (rc != CL2_OK) return rc;
    // combineLoops (geo.py:90-112) drops a candidate whose coordinates were already present from an EARLIER run;
    // the cis path then also applies filterDupLoops (callCisLoops.py:159-169: first position of every coordinate
    // tuple).  A candidate's test result depends on its coordinates only, so both are applied to the survivors.
    // survivors in candidate order: run by run as given, ascending cluster id inside a run
    std::vector<int64_t> sorder((size_t)nk);
    for (int64_t q = 0; q < nk; q++) sorder[(size_t)q] = q;
    std::sort(sorder.begin(), sorder.end(), [&](int64_t a, int64_t b) {
        const int64_t ka = sel[(size_t)a], kb = sel[(size_t)b];
        if (pid[(size_t)ka] != pid[(size_t)kb]) return pid[(size_t)ka] < pid[(size_t)kb];
        if (akey[(size_t)ka] != akey[(size_t)kb]) return akey[(size_t)ka] < akey[(size_t)kb];
        return ka < kb;
    });
    std::unordered_map<Key4, int32_t, Key4Hash> first_pass;
    first_pass.reserve((size_t)nk * 2 + 1);
    for (int64_t qq = 0; qq < nk; qq++) {
        const int64_t q = sorder[(size_t)qq];
        const int64_t k = sel[(size_t)q];
        Key4 key = {{all[4 * k], all[4 * k + 1], all[4 * k + 2], all[4 * k + 3]}};
        auto it = first_pass.find(key);
        if (it != first_pass.end()) {
            if (cis || it->second != pid[(size_t)k]) continue;
        } else {
            first_pass.emplace(key, pid[(size_t)k]);
        }
        r->coords.insert(r->coords.end(), all.begin() + 4 * k, all.begin() + 4 * k + 4);
        r->pass_id.push_back(pid[(size_t)k]);
        r->core.insert(r->core.end(), core.begin() + 4 * q, core.begin() + 4 * q + 4);
        r->hyp.push_back(hyp[(size_t)q]);
        r->stats.insert(r->stats.end(), stats.begin() + 13 * q, stats.begin() + 13 * q + 13);
    }
    r->n = (int64_t)r->pass_id.size();
    r->seconds[2] = now_s() - t2;
    return CL2_OK;
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
        for (int i = 0; i < 3; i++) seconds[i] = r->seconds[i];
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
