// cl2_host.cu -- native host-side pieces of the path that are order-dependent and sequential in the reference.
#include "cl2_common.cuh"
#include <algorithm>
#include "cl2_stats.cuh"

static inline bool anchor_ov(int64_t xa, int64_t xb, int64_t ya, int64_t yb) {
    // geo.py:65-73, closed intervals
    return (ya <= xa && xa <= yb) || (ya <= xb && xb <= yb) || (xa <= ya && ya <= xb) || (xa <= yb && yb <= xb);
}
static inline bool loop_ov(const int64_t *a, const int64_t *b) {
    return anchor_ov(a[0], a[1], b[0], b[1]) && anchor_ov(a[2], a[3], b[2], b[3]);
}

// selSigLoops, callCisLoops.py:379-422.  Exact list-order semantics; the only shortcut is an x-interval pre-test
// against the running bounding interval of the group before the member scan.
extern "C" int cl2_sel_sig_loops(const int64_t *coords, const double *es, int64_t L, int64_t *out_idx, int64_t *out_n) {
    if (L < 0 || (L > 0 && (!coords || !es || !out_idx)) || !out_n) return CL2_ERR_ARG;
    std::vector<char> skip((size_t)L, 0);
    std::vector<int64_t> reps;
    std::vector<int64_t> grp;
    for (int64_t i = 0; i < L; i++) {
        if (skip[i]) continue;
        grp.clear();
        grp.push_back(i);
        int64_t gx0 = coords[4 * i], gx1 = coords[4 * i + 1], gy0 = coords[4 * i + 2], gy1 = coords[4 * i + 3];
        for (int64_t j = i + 1; j < L; j++) {
            const int64_t *cj = coords + 4 * j;
            // no member can overlap j unless j overlaps the union box of the members
            if (cj[1] < gx0 || cj[0] > gx1 || cj[3] < gy0 || cj[2] > gy1) continue;
            for (size_t p = 0; p < grp.size(); p++) {
                if (loop_ov(coords + 4 * grp[p], cj)) {
                    grp.push_back(j);
                    skip[j] = 1;
                    gx0 = std::min(gx0, cj[0]); gx1 = std::max(gx1, cj[1]);
                    gy0 = std::min(gy0, cj[2]); gy1 = std::max(gy1, cj[3]);
                    break;
                }
            }
        }
        // swap-based selection (:405-412): slot 0 ends up holding the first member attaining the maximum ES
        int64_t best = grp[0];
        for (size_t p = 1; p < grp.size(); p++)
            if (es[best] < es[grp[p]]) best = grp[p];
        reps.push_back(best);
    }
    // "search again" (:414-421): reps grows while scanning
    for (int64_t a = 0; a < L; a++) {
        bool hit = false;
        for (size_t r = 0; r < reps.size(); r++)
            if (loop_ov(coords + 4 * a, coords + 4 * reps[r])) { hit = true; break; }
        if (!hit) reps.push_back(a);
    }
    for (size_t r = 0; r < reps.size(); r++) out_idx[r] = reps[r];
    *out_n = (int64_t)reps.size();
    return CL2_OK;
}

// Host entry points to the fp64 tail functions the counting kernel evaluates on the device (same source, cl2_stats.cuh);
// scipy-style arguments: sf(k) = P(X > k).
extern "C" double cl2_stat_poisson_sf(double k, double mu) { return cl2_poisson_sf_ge(floor(k) + 1.0, mu); }
extern "C" double cl2_stat_binom_sf(double k, double n, double p) { return cl2_binom_sf_ge(floor(k) + 1.0, n, p); }
extern "C" double cl2_stat_hypergeom_sf(double k, double M, double n, double N) {
    return cl2_hypergeom_sf_ge(floor(k) + 1.0, M, n, N);
}
