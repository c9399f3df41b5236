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

// selSigLoops, callCisLoops.py:379-422, with the reference's exact list-order semantics but without its O(L^2) scans.
//  * overlap graph: sweep over the loops sorted by x_start, edge iff both anchors overlap (geo.py:65-87)
//  * group of seed i (:388-398): j > i joins iff it overlaps a member already in the group when j is examined; members
//    are appended in ascending j, so membership(j) = some neighbour m < j is a member -> ordered closure with a
//    min-heap of pending neighbours.  A loop already claimed by an earlier group can still join later groups.
//  * representative (:405-412): first member attaining the maximum ES
//  * "search again" (:414-421): a loop is appended unless it overlaps a representative chosen so far
extern "C" int cl2_sel_sig_loops(const int64_t *coords, const double *es, int64_t L, int64_t *out_idx, int64_t *out_n) {
    if (L < 0 || (L > 0 && (!coords || !es || !out_idx)) || !out_n) return CL2_ERR_ARG;
    const size_t n = (size_t)L;
    std::vector<int64_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = (int64_t)i;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        return coords[4 * a] != coords[4 * b] ? coords[4 * a] < coords[4 * b] : a < b;
    });
    std::vector<std::vector<int64_t>> adj(n);
    for (size_t p = 0; p < n; p++) {
        const int64_t a = order[p];
        const int64_t *ca = coords + 4 * a;
        for (size_t q = p + 1; q < n; q++) {
            const int64_t b = order[q];
            if (coords[4 * b] > ca[1]) break;            // x_start beyond a's x_end: no later loop overlaps in x
            if (loop_ov(ca, coords + 4 * b)) {
                adj[a].push_back(b);
                adj[b].push_back(a);
            }
        }
    }
    std::vector<char> skip(n, 0);
    std::vector<int64_t> reps;
    std::vector<int64_t> heap;                            // min-heap of pending members (indices > seed)
    std::vector<int64_t> stamp(n, -1);
    auto cmp = [](int64_t a, int64_t b) { return a > b; };
    for (int64_t i = 0; i < L; i++) {
        if (skip[i]) continue;
        heap.clear();
        stamp[i] = i;
        int64_t best = i;
        for (int64_t nb : adj[i])
            if (nb > i && stamp[nb] != i) { stamp[nb] = i; heap.push_back(nb); std::push_heap(heap.begin(), heap.end(), cmp); }
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            int64_t j = heap.back();
            heap.pop_back();
            skip[j] = 1;                                  // j joins the group (in ascending order)
            if (es[best] < es[j]) best = j;               // strict '<': the first maximum stays
            for (int64_t nb : adj[j])
                if (nb > j && stamp[nb] != i) { stamp[nb] = i; heap.push_back(nb); std::push_heap(heap.begin(), heap.end(), cmp); }
        }
        reps.push_back(best);
    }
    std::vector<char> is_rep(n, 0);
    for (int64_t r : reps) is_rep[r] = 1;
    for (int64_t a = 0; a < L; a++) {
        bool hit = is_rep[a];
        if (!hit)
            for (int64_t nb : adj[a])
                if (is_rep[nb]) { hit = true; break; }
        if (!hit) { is_rep[a] = 1; reps.push_back(a); }
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
