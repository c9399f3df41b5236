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

// ---- text of the loop table -----------------------------------------------------------------------------------
// loops2txt (cLoops2/io.py:517-561) writes every field through Python's str().  For floats that is repr(): the
// shortest digit string that round-trips, in fixed notation with at least one decimal for 1e-4 <= |x| < 1e16 and as
// d.ddde-XX (two exponent digits at least) otherwise.  std::to_chars yields the same shortest digits; the layout is
// redone here to Python's rules so the file is byte-identical without a Python-level loop over 24 fields per row.
#include <charconv>

static char *put_str(char *o, const char *s) {
    while (*s) *o++ = *s++;
    return o;
}
static char *put_i64(char *o, int64_t v) {
    auto r = std::to_chars(o, o + 24, v);
    return r.ptr;
}
// Python's repr(float)
static char *put_pyfloat(char *o, double x) {
    if (x != x) return put_str(o, "nan");
    if (x == HUGE_VAL) return put_str(o, "inf");
    if (x == -HUGE_VAL) return put_str(o, "-inf");
    if (x == 0.0) return put_str(o, std::signbit(x) ? "-0.0" : "0.0");
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);     // [-]d[.ddd]e[+-]XX, shortest
    const char *p = buf, *end = r.ptr;
    if (*p == '-') *o++ = *p++;
    char digits[24];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        p++;
    }
    int exp10 = 0;
    if (p < end) {
        p++;                                                   // 'e'
        bool neg = *p == '-';
        if (*p == '-' || *p == '+') p++;
        while (p < end) exp10 = exp10 * 10 + (*p++ - '0');
        if (neg) exp10 = -exp10;
    }
    const int decpt = exp10 + 1;                               // position of the decimal point relative to the digits
    if (decpt > 16 || decpt < -3) {                            // exponent form: d[.ddd]e+XX
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            for (int i = 1; i < nd; i++) *o++ = digits[i];
        }
        *o++ = 'e';
        int e = decpt - 1;
        *o++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        if (e < 10) *o++ = '0';
        auto r2 = std::to_chars(o, o + 8, e);
        return r2.ptr;
    }
    if (decpt <= 0) {                                          // 0.000ddd
        *o++ = '0';
        *o++ = '.';
        for (int i = 0; i < -decpt; i++) *o++ = '0';
        for (int i = 0; i < nd; i++) *o++ = digits[i];
        return o;
    }
    for (int i = 0; i < decpt; i++) *o++ = i < nd ? digits[i] : '0';
    *o++ = '.';
    if (nd > decpt) {
        for (int i = decpt; i < nd; i++) *o++ = digits[i];
    } else {
        *o++ = '0';
    }
    return o;
}

extern "C" int64_t cl2_format_floats(const double *v, int64_t n, char *out, int64_t cap) {
    if (n < 0 || (n > 0 && (!v || !out))) return CL2_ERR_ARG;
    char *o = out;
    for (int64_t i = 0; i < n; i++) {
        if (o - out + 40 > cap) return CL2_ERR_RANGE;
        o = put_pyfloat(o, v[i]);
        *o++ = '\n';
    }
    return o - out;
}

extern "C" int64_t cl2_loops_text(const char *chrA, const char *chrB, int64_t first_id, int64_t n, int32_t cis,
                                  const int64_t *coords, const int64_t *ra, const int64_t *rb, const int64_t *rab,
                                  const double *density, const double *es, const double *p2ll, const double *fdr,
                                  const double *nbp, const double *hyp, const double *pop, const double *px,
                                  const double *py, char *out, int64_t cap) {
    if (!chrA || !chrB || n < 0 || (n > 0 && (!coords || !ra || !rb || !rab || !density || !es || !p2ll || !fdr || !nbp ||
                                               !hyp || !pop || !px || !py || !out)))
        return CL2_ERR_ARG;
    const int64_t need = 2 * ((int64_t)strlen(chrA) + (int64_t)strlen(chrB)) + 640;
    char *o = out;
    for (int64_t i = 0; i < n; i++) {
        if (o - out + need > cap) return CL2_ERR_RANGE;
        const int64_t xs = coords[4 * i], xe = coords[4 * i + 1], ys = coords[4 * i + 2], ye = coords[4 * i + 3];
        o = put_str(o, "loop_"); o = put_str(o, chrA); *o++ = '-'; o = put_str(o, chrB); *o++ = '-';
        o = put_i64(o, first_id + i); *o++ = '\t';
        o = put_str(o, chrA); *o++ = '\t';
        o = put_i64(o, xs); *o++ = '\t';
        o = put_i64(o, xe); *o++ = '\t';
        o = put_str(o, chrB); *o++ = '\t';
        o = put_i64(o, ys); *o++ = '\t';
        o = put_i64(o, ye); *o++ = '\t';
        const double xc = (double)(xs + xe) / 2.0, yc = (double)(ys + ye) / 2.0;      // callCisLoops.py:107-112
        if (cis) { const double d = yc - xc; o = put_pyfloat(o, d < 0 ? -d : d); }
        else o = put_str(o, "-1");                                                   // callTransLoops.py:78
        *o++ = '\t';
        o = put_pyfloat(o, xc); *o++ = '\t';
        o = put_pyfloat(o, yc); *o++ = '\t';
        o = put_i64(o, ra[i]); *o++ = '\t';
        o = put_i64(o, rb[i]); *o++ = '\t';
        o = put_str(o, cis ? "True" : "False"); *o++ = '\t';
        o = put_i64(o, rab[i]); *o++ = '\t';
        const double *cols[9] = {density, es, p2ll, fdr, nbp, hyp, pop, px, py};
        for (int k = 0; k < 9; k++) { o = put_pyfloat(o, cols[k][i]); *o++ = '\t'; }
        *o++ = '1';
        *o++ = '\n';
    }
    return o - out;
}
