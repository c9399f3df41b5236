// cl2_count.cu -- device form of cLoops2/ds.py:141-198 (XY) and the per-candidate counting behind estLoopSig
// (callCisLoops.py:173-354, callTransLoops.py:106-193).
//
// XY keeps sorted xs / ys plus value->row dicts and answers queryPeak(l,r) = {rows with l<=X<=r or l<=Y<=r} as Python
// sets.  Here the index is two arrays, (x,y) sorted by x and (y,x) sorted by y, and every set size the reference takes
// becomes a binary-search range plus a warp-strided scan of the range members:
//     |P(I)|            = |Xrange(I)| + #{j in Yrange(I) : x_j not in I}
//     |P(I1) & P(I2)|   = #{members of P(I1) that have an end in I2}
// The reference's window bounds are multiples of 0.25 evaluated in fp64 (exact); they are carried here as integers
// scaled by 4, so l <= X <= r is decided in exact integer arithmetic.  One warp per candidate loop.
#include "cl2_points.cuh"
#include "cl2_stats.cuh"
#include <math_constants.h>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------ index build
__global__ void __launch_bounds__(256) k_index_gather(const int2 *__restrict__ xy, const uint32_t *__restrict__ row,
                                                       int2 *__restrict__ out, int64_t n, int which) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int2 p = __ldg(&xy[row[i]]);
    out[i] = which == 0 ? p : make_int2(p.y, p.x);
}

// bucket table over a coordinate-sorted array: tab[b] = first position whose coordinate is >= b << shift; tab[nb] = n
__global__ void __launch_bounds__(256) k_bucket_table(const int2 *__restrict__ a, int64_t n, int shift, int nb,
                                                       uint32_t *__restrict__ tab) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bi = a[i].x >> shift;
    const int bp = i > 0 ? a[i - 1].x >> shift : -1;
    for (int b = bp + 1; b <= bi; b++) tab[b] = (uint32_t)i;
    if (i == n - 1)
        for (int b = bi + 1; b <= nb; b++) tab[b] = (uint32_t)n;
}
static int bucket_shift(int64_t maxcoord, int64_t n) {        // about 16 rows per bucket
    int64_t want = n / 16 > 1024 ? n / 16 : 1024;
    int s = 0;
    while ((maxcoord >> s) + 1 > want) s++;
    return s;
}
static int build_table(cl2_ctx *ctx, const int2 *a, int64_t n, int64_t maxcoord, uint32_t **tab, int *shift, int *nb) {
    *tab = nullptr;
    if (n <= 0 || n > (int64_t)UINT32_MAX) return CL2_OK;
    *shift = bucket_shift(maxcoord, n);
    *nb = (int)(maxcoord >> *shift) + 1;
    CL2_TRY(cl2_dev_alloc(ctx, tab, (size_t)*nb + 1));
    FamTimer t(ctx, FAM_INDEX, 8.0 * (double)n + 4.0 * (double)*nb);
    k_bucket_table<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(a, n, *shift, *nb, *tab);
    return CL2_OK;
}

// (y, x) sorted by y plus the file row of every sorted position; kept on the points object
int cl2_points_yorder(cl2_ctx *ctx, cl2_points *pts) {
    if (pts->byy || pts->n == 0) return CL2_OK;
    const int64_t n = pts->n;
    Scratch sc(ctx);
    uint64_t *k0, *k1, *ks;
    uint32_t *v0, *v1, *vs;
    CL2_TRY(sc.get(&k0, (size_t)n));
    CL2_TRY(sc.get(&k1, (size_t)n));
    CL2_TRY(sc.get_pool(&v0, (size_t)n));           // one of the two is kept as pts->rowy
    CL2_TRY(sc.get_pool(&v1, (size_t)n));
    cl2_keygen gen;
    gen.kind = CL2_KEY_Y;
    gen.xy = pts->xy;
    CL2_TRY(cl2_radix_sort_gen(ctx, gen, k0, k1, v0, v1, n, bits_for((uint64_t)pts->ext[3]), &ks, &vs));
    CL2_TRY(cl2_dev_alloc(ctx, &pts->byy, (size_t)n));
    {
        FamTimer t(ctx, FAM_INDEX, 20.0 * (double)n);
        k_index_gather<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(pts->xy, vs, pts->byy, n, 1);
    }
    sc.release(vs);
    pts->rowy = vs;
    CL2_CUDA(ctx, cudaGetLastError());
    return CL2_OK;
}

extern "C" int cl2_index_build(cl2_ctx *ctx, cl2_points *pts, cl2_index **out) {
    if (!ctx || !pts || !out) return CL2_ERR_ARG;
    *out = nullptr;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    cl2_index *ix = new (std::nothrow) cl2_index();
    if (!ix) return CL2_ERR_NOMEM;
    const int64_t n = pts->n;
    ix->n = n;
    ix->small = (pts->ext[1] < (1ll << 29) && pts->ext[3] < (1ll << 29)) ? 1 : 0;
    ix->ymax = pts->ext[3];
    int rc = CL2_OK;
    if (n > 0) {
        Scratch sc(ctx);
        uint64_t *k0 = nullptr, *k1 = nullptr, *ks;
        uint32_t *v0 = nullptr, *v1 = nullptr, *vs;
        do {
            if ((rc = cl2_points_yorder(ctx, pts))) break;
            // the x-order a clustering run left behind (cl2_block.cu: k_xorder_from_grid), unless a column was too long
            if (pts->byx) {
                int32_t bad = 1;
                if ((rc = cl2_read_i32(ctx, pts->xorder_bad, &bad))) break;
                if (bad == 0) {
                    ix->byx = pts->byx;
                    pts->byx = nullptr;
                }
                cl2_points_drop_xorder(ctx, pts);
            }
            if ((rc = cl2_dev_alloc(ctx, &ix->byy, (size_t)n))) break;
            // the index owns a copy of the y-order so it may outlive the points object
            if (cudaMemcpyAsync(ix->byy, pts->byy, (size_t)n * sizeof(int2), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
                rc = cl2_fail(ctx, CL2_ERR_CUDA, "cl2_index_build: copy failed");
                break;
            }
            if (!ix->byx) {
                if ((rc = sc.get(&k0, (size_t)n)) || (rc = sc.get(&k1, (size_t)n)) || (rc = sc.get(&v0, (size_t)n)) ||
                    (rc = sc.get(&v1, (size_t)n)))
                    break;
                if ((rc = cl2_dev_alloc(ctx, &ix->byx, (size_t)n))) break;
                cl2_keygen gen;
                gen.kind = CL2_KEY_X;
                gen.xy = pts->xy;
                rc = cl2_radix_sort_gen(ctx, gen, k0, k1, v0, v1, n, bits_for((uint64_t)pts->ext[1]), &ks, &vs);
                if (rc) break;
                FamTimer t(ctx, FAM_INDEX, 20.0 * (double)n);
                k_index_gather<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(pts->xy, vs, ix->byx, n, 0);
            }
            if ((rc = build_table(ctx, ix->byx, n, pts->ext[1], &ix->tx, &ix->sx, &ix->nbx))) break;
            if ((rc = build_table(ctx, ix->byy, n, pts->ext[3], &ix->ty, &ix->sy, &ix->nby))) break;
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
                rc = cl2_fail(ctx, CL2_ERR_CUDA, "cl2_index_build: kernel failure");
            }
        } while (0);
    }
    if (rc != CL2_OK) {
        cl2_dev_free(ctx, ix->byx);
        cl2_dev_free(ctx, ix->byy);
        cl2_dev_free(ctx, ix->tx);
        cl2_dev_free(ctx, ix->ty);
        delete ix;
        return rc;
    }
    *out = ix;
    return CL2_OK;
}

extern "C" void cl2_index_free(cl2_ctx *ctx, cl2_index *idx) {
    if (!ctx || !idx) return;
    cudaSetDevice(ctx->device);
    cl2_dev_free(ctx, idx->byx);
    cl2_dev_free(ctx, idx->byy);
    cl2_dev_free(ctx, idx->tx);
    cl2_dev_free(ctx, idx->ty);
    cl2_dev_free(ctx, idx->wm);
    delete idx;
}

// ------------------------------------------------------------------------------------------------ counting
#define CNT_WARPS 8
#define CNT_MAXW 16   // 2*win <= 16

// What the kernels see of a cl2_index.  tx / ty are bucket tables over the sorted coordinate: t[b] = first position
// whose coordinate is >= b << shift (nb + 1 entries, t[nb] = n), so a search starts from a ~16-row bracket instead
// of the whole array.
struct IdxView {
    const int2 *byx, *byy;
    const uint32_t *tx, *ty;
    int64_t n;
    int sx, sy, nbx, nby;
    int lg;               // ceil(log2 n): depth of one binary search in SURVEY.md 8d's byte count
    int small;            // every coordinate is below 2^29 (4 * c fits an int: closed-form window masks)
    const uint2 *wm;      // wavelet matrix over y in x order (null until built), see cl2_index
    int wm_L;
    uint32_t wm_nw;
    uint32_t wm_z[32];
};
struct Arr {
    const int2 *a;
    const uint32_t *tab;
    int shift, nb;
};
static IdxView cl2_view(const cl2_index *idx) {
    IdxView v = {idx->byx, idx->byy, idx->tx, idx->ty, idx->n, idx->sx, idx->sy, idx->nbx, idx->nby, 1, idx->small,
                 idx->wm, idx->wm_L, (uint32_t)idx->wm_nw, {0}};
    for (int l = 0; l < 32; l++) v.wm_z[l] = idx->wm_z[l];
    while (((int64_t)1 << v.lg) < idx->n) v.lg++;
    return v;
}
__device__ __forceinline__ Arr arr_x(const IdxView &ix) { Arr r = {ix.byx, ix.tx, ix.sx, ix.nbx}; return r; }
__device__ __forceinline__ Arr arr_y(const IdxView &ix) { Arr r = {ix.byy, ix.ty, ix.sy, ix.nby}; return r; }

struct Iv {           // closed integer interval; empty when lo > hi
    long long lo, hi;
};
__device__ __forceinline__ long long floordiv4(long long a) { return a >> 2; }              // arithmetic shift = floor
__device__ __forceinline__ long long ceildiv4(long long a) { return -((-a) >> 2); }
// interval with bounds given scaled by 4: {v integer : lo4 <= 4v <= hi4}
__device__ __forceinline__ Iv iv4(long long lo4, long long hi4) {
    Iv r;
    r.lo = ceildiv4(lo4);
    r.hi = floordiv4(hi4);
    return r;
}
// The same interval cut to the coordinate domain [0, INT_MAX) (cl2_points_upload refuses anything else), held as
// {lo, width} so that membership is one subtract and one unsigned compare.  Empty: lo = INT_MAX, which no coordinate
// reaches.
struct Iv32 {
    int lo;
    unsigned wd;
};
__device__ __forceinline__ Iv32 iv32(long long lo, long long hi) {
    Iv32 r;
    if (lo < 0) lo = 0;
    if (hi > (long long)INT_MAX) hi = INT_MAX;
    if (lo > hi) { r.lo = INT_MAX; r.wd = 0; }
    else { r.lo = (int)lo; r.wd = (unsigned)(hi - lo); }
    return r;
}
__device__ __forceinline__ Iv32 iv32(const Iv &I) { return iv32(I.lo, I.hi); }
__device__ __forceinline__ bool in32(const Iv32 &I, int v) { return (unsigned)(v - I.lo) <= I.wd; }

// Bracket [lo, hi] that holds "first index with first-coordinate >= v" (np.searchsorted side="left", ds.py:170)
__device__ __forceinline__ void bracket(const Arr &A, int64_t n, long long v, int64_t &lo, int64_t &hi) {
    if (v <= 0) { lo = hi = 0; return; }                    // coordinates are >= 0
    if (v > (long long)INT_MAX) { lo = hi = n; return; }    // and < INT_MAX
    lo = 0;
    hi = n;
    if (A.tab) {
        int b = (int)v >> A.shift;
        if (b >= A.nb) { lo = hi = n; }
        else { lo = A.tab[b]; hi = A.tab[b + 1]; }
    }
}
// The search inside a bracket, by the whole warp: every step 32 lanes probe 32 evenly spaced positions and a ballot
// picks the sub-range, so even a search over 10^8 rows is 6 dependent loads instead of 27; with the bucket table the
// loop body rarely runs at all.  All lanes call with the same arguments and get the same result.
__device__ __forceinline__ void narrow(const int2 *__restrict__ a, int v, int64_t &lo, int64_t &hi, int lane) {
    while (hi - lo > 32) {
        int64_t step = (hi - lo) >> 5;
        int64_t pos = lo + step * (lane + 1) - 1;
        bool less = a[pos].x < v;
        int k = __popc(__ballot_sync(0xffffffffu, less));   // monotone: lanes 0..k-1 are "less"
        int64_t nlo = k > 0 ? lo + step * k : lo;
        int64_t nhi = k < 32 ? lo + step * (k + 1) - 1 : hi;
        lo = nlo;
        hi = nhi;
    }
}
struct Range {
    int64_t b, e;
};
// positions of the rows whose sort coordinate lies in I; both ends are searched side by side so their loads overlap
__device__ __forceinline__ Range range_of(const Arr &A, int64_t n, const Iv &I, int lane) {
    Range r;
    if (I.lo > I.hi) { r.b = r.e = 0; return r; }
    int64_t lo1, hi1, lo2, hi2;
    bracket(A, n, I.lo, lo1, hi1);
    bracket(A, n, I.hi + 1, lo2, hi2);                       // side="right" on integers
    const int v1 = (int)(I.lo < 0 ? 0 : (I.lo > (long long)INT_MAX ? INT_MAX : I.lo));
    const int v2 = (int)(I.hi + 1 < 0 ? 0 : (I.hi + 1 > (long long)INT_MAX ? INT_MAX : I.hi + 1));
    narrow(A.a, v1, lo1, hi1, lane);
    narrow(A.a, v2, lo2, hi2, lane);
    const int64_t p1 = lo1 + lane, p2 = lo2 + lane;
    const bool l1 = p1 < hi1 && A.a[p1].x < v1;
    const bool l2 = p2 < hi2 && A.a[p2].x < v2;
    r.b = lo1 + __popc(__ballot_sync(0xffffffffu, l1));
    r.e = lo2 + __popc(__ballot_sync(0xffffffffu, l2));
    return r;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// ---- counting on the X-sorted list only ---------------------------------------------------------------------
// Every set size of the reference is a combination of 1-D range counts (pure position arithmetic on the two sorted
// arrays) and sums over the rows of an X-range (checked against the oracle for ordered, unordered, overlapping and
// clipped intervals):
//     |P(I)|          = cX(I) + cY(I) - #{x in I and y in I}
//     |P(U) & P(V)|   = cY(U & V) + sum over rows with x in U | V of  [(x in U or y in U) and (x in V or y in V)]
//                                                                   - [y in U and y in V]
// (a row whose x lies outside U | V can only be in both sets through y, which is what cY(U & V) counts; the rows
// scanned are subtracted from it again).  The Y-sorted array is only ever searched, never scanned, and no row is
// visited through two lists, so there is no de-duplication.

// lower bound (first position whose sort coordinate is >= v) searched by ONE lane for its own value: the bucket
// table brackets the answer to ~16 rows, then a plain binary search.  32 lanes run 32 independent searches.
__device__ __forceinline__ uint32_t lane_lb(const Arr &A, int64_t n, long long v) {
    if (v <= 0) return 0u;                                   // coordinates are >= 0
    if (v > (long long)INT_MAX) return (uint32_t)n;          // and < INT_MAX
    uint32_t lo = 0, hi = (uint32_t)n;
    if (A.tab) {
        const int b = (int)v >> A.shift;
        if (b >= A.nb) return (uint32_t)n;
        lo = A.tab[b];
        hi = A.tab[b + 1];
    }
    const int key = (int)v;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;                 // n < 2^31: no overflow
        if (A.a[mid].x < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// one bound of the closed interval [lo, hi] in one of the two lists: end 0 = first position inside, 1 = one past the
// last (searchsorted left / right, ds.py:170-171); an empty interval gets the empty range [0, 0)
__device__ __forceinline__ uint32_t bound_task(const IdxView &ix, long long lo, long long hi, int list, int end) {
    if (lo > hi) return 0u;
    return list ? lane_lb(arr_y(ix), ix.n, end ? hi + 1 : lo) : lane_lb(arr_x(ix), ix.n, end ? hi + 1 : lo);
}
struct IvPos {        // X-list range [xb, xe) and Y-list range [yb, ye) of an interval
    uint32_t xb, xe, yb, ye;
};
// lanes 4k .. 4k+3 hold xb, xe, yb, ye of interval k
__device__ __forceinline__ IvPos gather_pos(uint32_t pos, int k) {
    IvPos P;
    P.xb = __shfl_sync(0xffffffffu, pos, 4 * k);
    P.xe = __shfl_sync(0xffffffffu, pos, 4 * k + 1);
    P.yb = __shfl_sync(0xffffffffu, pos, 4 * k + 2);
    P.ye = __shfl_sync(0xffffffffu, pos, 4 * k + 3);
    return P;
}
__device__ __forceinline__ long long pos_max(const IvPos &P) {   // lower bound of |P(I)| (SURVEY.md 8d's R / W)
    const long long a = (long long)P.xe - P.xb, b = (long long)P.ye - P.yb;
    return a > b ? a : b;
}
// cY(U & V) from the Y ranges of U and V (lower / upper bounds are monotone in the coordinate)
__device__ __forceinline__ long long ycount_both(const IvPos &U, const IvPos &V) {
    const long long b = U.yb > V.yb ? U.yb : V.yb, e = U.ye < V.ye ? U.ye : V.ye;
    return e > b ? e - b : 0;
}

// The warps that work on one candidate together.  Candidates whose regions hold few rows are handled by one warp each;
// the rare large ones (a 1 Mb anchor of Hi-C-like data spans millions of rows) would leave that warp running long after
// the rest of the grid has finished, so they are handed to a second kernel in which all warps of a block share the scan
// of one candidate and combine their partial sums through shared memory.
struct Coop {
    int nw;      // warps sharing the candidate (1: no shared reduction, no block barrier)
    int part;    // this warp's index among them
    int *red;    // shared scratch of COOP_RED ints (nw > 1)
    bool defer;  // a single warp may decline a large candidate (the caller has the cooperative kernel to hand it to)
};
__device__ __forceinline__ Coop coop_solo() { Coop c = {1, 0, nullptr, false}; return c; }
#define COOP_RED 128
#define BIG_ROWS 8192        // rows (X-range lengths) from which a candidate goes to the cooperative kernel
// vals[0..k): warp-uniform partial sums on entry, the sums over the cooperating warps on return (all threads call)
template <int K>
__device__ __forceinline__ void coop_sum(const Coop &C, int (&vals)[K], int lane) {
    if (C.nw == 1) return;
    __syncthreads();
    if ((int)threadIdx.x < K) C.red[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < K; i++) atomicAdd(&C.red[i], vals[i]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; i++) vals[i] = C.red[i];
}

// sum over the rows with x in U | V:  cnt = [(x|y in U) and (x|y in V)] - [y in U and y in V],
// eu = [x in U and y in U], ev = [x in V and y in V]
__device__ __forceinline__ void pair_scan(const int2 *__restrict__ byx, const Iv32 U, const Iv32 V, const IvPos &PU,
                                          const IvPos &PV, int lane, const Coop &C, int &cnt, int &eu, int &ev) {
    uint32_t b0 = PU.xb, e0 = PU.xe, b1 = PV.xb, e1 = PV.xe;
    if (b0 <= e1 && b1 <= e0) {                              // overlapping or touching ranges: one scan
        b0 = b0 < b1 ? b0 : b1;
        e0 = e0 > e1 ? e0 : e1;
        b1 = e1 = 0;
    }
    int c = 0, u = 0, v = 0;
    const uint32_t stride = 32u * (uint32_t)C.nw;
#pragma unroll 1
    for (int rr = 0; rr < 2; rr++) {
        const uint32_t b = rr ? b1 : b0, e = rr ? e1 : e0;
        for (uint32_t p = b + 32u * (uint32_t)C.part + lane; p < e; p += stride) {
            const int2 P = byx[p];
            const bool px = in32(U, P.x), qy = in32(U, P.y), rx = in32(V, P.x), sy = in32(V, P.y);
            c += (int)((px || qy) && (rx || sy)) - (int)(qy && sy);
            u += (int)(px && qy);
            v += (int)(rx && sy);
        }
    }
    int r[3] = {warp_sum(c), warp_sum(u), warp_sum(v)};
    coop_sum(C, r, lane);
    cnt = r[0];
    eu = r[1];
    ev = r[2];
}
// #{x in I and y in I} over the X range of I
__device__ __forceinline__ int both_scan(const int2 *__restrict__ byx, const Iv32 I, const IvPos &P, int lane, const Coop &C) {
    int c = 0;
    for (uint32_t p = P.xb + 32u * (uint32_t)C.part + lane; p < P.xe; p += 32u * (uint32_t)C.nw) c += (int)in32(I, byx[p].y);
    int r[1] = {warp_sum(c)};
    coop_sum(C, r, lane);
    return r[0];
}

struct WarpAcc {
    int na[CNT_MAXW];
    int nb[CNT_MAXW];
    int nab[CNT_MAXW * CNT_MAXW];
    // permuted windows as {lo, width} (Iv32 layout, one 8-byte load per test) and their upper bounds
    int2 aw[CNT_MAXW], bw[CNT_MAXW];
    int ahi[CNT_MAXW], bhi[CNT_MAXW];
    // 2*win <= 10: window bounds (a-windows then b-windows) and their positions xb, xe, yb, ye in the two lists
    long long wlo[2 * CNT_MAXW], whi[2 * CNT_MAXW];
    uint32_t pos[4][2 * CNT_MAXW];
};

// bit t set iff window t holds c
__device__ __noinline__ unsigned coord_mask(const int2 *__restrict__ w, int W, int c) {
    unsigned m = 0;
    if (W == 10) {
#pragma unroll
        for (int t = 0; t < 10; t++) {
            const int2 q = w[t];
            m |= ((unsigned)(c - q.x) <= (unsigned)q.y) ? (1u << t) : 0u;
        }
    } else {
        for (int t = 0; t < W; t++) {
            const int2 q = w[t];
            m |= ((unsigned)(c - q.x) <= (unsigned)q.y) ? (1u << t) : 0u;
        }
    }
    return m;
}

// na / nb live in registers: lane t carries the count of window t
__device__ __forceinline__ void window_accumulate(WarpAcc &A, int W, unsigned ma, unsigned mb, int lane, int &na_r,
                                                  int &nb_r) {
    // only the windows some lane of this batch falls into are visited (a sorted batch touches 1-3 of them)
    unsigned ua = __reduce_or_sync(0xffffffffu, ma), ub = __reduce_or_sync(0xffffffffu, mb);
    while (ua) {
        int t = __ffs(ua) - 1;
        ua &= ua - 1;
        int c = __popc(__ballot_sync(0xffffffffu, (ma >> t) & 1u));
        na_r += lane == t ? c : 0;
    }
    while (ub) {
        int t = __ffs(ub) - 1;
        ub &= ub - 1;
        int c = __popc(__ballot_sync(0xffffffffu, (mb >> t) & 1u));
        nb_r += lane == t ? c : 0;
    }
    if (ma && mb) {
        unsigned a = ma;
        while (a) {
            int i = __ffs(a) - 1;
            a &= a - 1;
            unsigned b = mb;
            while (b) {
                int j = __ffs(b) - 1;
                b &= b - 1;
                atomicAdd(&A.nab[i * W + j], 1);
            }
        }
    }
}

// ---- per-loop building blocks (one warp) -------------------------------------------------------------------
// g_bytes: algorithmic bytes of the work done for this candidate: 16 B per member of every region
// queried, members counted from below as max(cX, cY), plus 2 arrays x 2 searches x ceil(log2 N) x 8 B per interval
// query; every lane carries the same value.
// queryLoop counts + the P2LL window (cis callCisLoops.py:287-289, 302-305; trans callTransLoops.py:124-126, 142-144)
// Returns false (nothing scanned, nothing counted) when a single warp is asked for a candidate with more than BIG_ROWS
// rows to scan: the caller hands it to the cooperative kernel.
__device__ bool loop_core(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int mode, int lane,
                          const Coop &C, int64_t out[4], long long *g_bytes) {
    // intervals 0..3: A, B, the lower-left pair A2, B2; lane 4k + 2*list + end searches one bound of interval k
    const int k = (lane >> 2) & 3;
    long long lo, hi;
    if (k == 0) { lo = xs; hi = xe; }
    else if (k == 1) { lo = ys; hi = ye; }
    else if (k == 2) {
        if (mode == CL2_MODE_CIS) { lo = xe; hi = xe + (xe - xs); }
        else { lo = xs - (xe - xs); hi = xs; }
    } else { lo = ys - (ye - ys); hi = ys; }
    const uint32_t pos = lane < 16 ? bound_task(ix, lo, hi, (lane >> 1) & 1, lane & 1) : 0u;
    const IvPos PA = gather_pos(pos, 0), PB = gather_pos(pos, 1), PA2 = gather_pos(pos, 2), PB2 = gather_pos(pos, 3);
    if (C.defer && (long long)(PA.xe - PA.xb) + (PB.xe - PB.xb) + (PA2.xe - PA2.xb) + (PB2.xe - PB2.xb) > BIG_ROWS) return false;
    const Iv32 A = iv32(xs, xe), B = iv32(ys, ye);
    Iv32 A2 = mode == CL2_MODE_CIS ? iv32(xe, xe + (xe - xs)) : iv32(xs - (xe - xs), xs);
    const Iv32 B2 = iv32(ys - (ye - ys), ys);
    int cab, ea, eb, clow, t1, t2;
    pair_scan(ix.byx, A, B, PA, PB, lane, C, cab, ea, eb);
    pair_scan(ix.byx, A2, B2, PA2, PB2, lane, C, clow, t1, t2);
    out[0] = ((int64_t)PA.xe - PA.xb) + ((int64_t)PA.ye - PA.yb) - ea;
    out[1] = ((int64_t)PB.xe - PB.xb) + ((int64_t)PB.ye - PB.yb) - eb;
    out[2] = ycount_both(PA, PB) + cab;
    out[3] = ycount_both(PA2, PB2) + clow;
    *g_bytes += 16 * (pos_max(PA) + pos_max(PB) + pos_max(PA2) + pos_max(PB2)) + 4 * 32 * (long long)ix.lg;
    return true;
}

// estAnchorSig counts of both anchors (callCisLoops.py:226-247): count = |P(l,r)|, wide = |P(max(0, m-5len), m+5len)|
__device__ void loop_anchors(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int lane,
                             const Coop &C, int64_t *cx, int64_t *wx, int64_t *cy, int64_t *wy, long long *g_bytes) {
    // intervals 0..3: anchor x, its wide window, anchor y, its wide window
    const int k = (lane >> 2) & 3;
    const long long l = k < 2 ? xs : ys, r = k < 2 ? xe : ye;
    const long long m4 = 2 * (l + r), len = r - l;
    long long s4 = m4 - 20 * len;
    if (s4 < 0) s4 = 0;
    const Iv Wd = iv4(s4, m4 + 20 * len);
    Iv I;
    if (k & 1) I = Wd; else { I.lo = l; I.hi = r; }
    const uint32_t pos = lane < 16 ? bound_task(ix, I.lo, I.hi, (lane >> 1) & 1, lane & 1) : 0u;
    int64_t res[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const IvPos P = gather_pos(pos, q);
        const long long il = __shfl_sync(0xffffffffu, I.lo, 4 * q), ih = __shfl_sync(0xffffffffu, I.hi, 4 * q);
        res[q] = ((int64_t)P.xe - P.xb) + ((int64_t)P.ye - P.yb) - both_scan(ix.byx, iv32(il, ih), P, lane, C);
        if (q & 1) *g_bytes += 16 * pos_max(P) + 2 * 32 * (long long)ix.lg;
    }
    *cx = res[0]; *wx = res[1]; *cy = res[2]; *wy = res[3];
}

// transpose of the 32 x 32 bit matrix held one row per lane: afterwards lane t holds column t (bit l = bit t of lane
// l's word), i.e. the ballot of bit t, for all 32 bits in 5 exchange steps
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, int lane) {
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const unsigned m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
    }
    return x;
}

// membership masks of one coordinate in the W windows of a family ({lo, width} pairs in shared memory)
__device__ __forceinline__ unsigned win_mask(const int2 *__restrict__ w, int W, int c) {
    unsigned m = 0;
#pragma unroll
    for (int t = 0; t < 10; t++) {
        if (t < W) {
            const int2 q = w[t];
            m |= ((unsigned)(c - q.x) <= (unsigned)q.y) ? (1u << t) : 0u;
        }
    }
    return m;
}
// The same mask in closed form.  The windows of a family are regular: window k (k = 0..2*win, k = win being the
// anchor itself, which is not a window) covers base4 + k*s4 <= 4c <= base4 + k*s4 + w2, so with t = 4c - base4 the
// windows holding c are k = ceil((t - w2) / s4) .. floor(t / s4); the two quotients come from a multiply-shift by the
// reciprocal of s4 (exact for numerators below 2^31).  Valid when no window was clipped at 0 on its upper end
// (callCisLoops.py:193-196 clips to max(0, .), which turns such a window into [0, 0]), coordinates stay below 2^29 and
// the family spans less than 2^30 quarter base pairs; otherwise win_mask() is used.
struct FamDiv {
    int base4;
    unsigned tmax, w2;
};
struct WinDiv {
    FamDiv a, b;
    unsigned s4m1, magic, valid11, lowmask;
    int shift, win;
};
__device__ __forceinline__ unsigned div_mask(const FamDiv &F, const WinDiv &D, int c) {
    const unsigned t = (unsigned)(4 * c - F.base4);
    const unsigned khi = (unsigned)(((unsigned long long)t * D.magic) >> D.shift);
    int u = (int)(t - F.w2 + D.s4m1);
    u = u < 0 ? 0 : u;
    const unsigned klo = (unsigned)(((unsigned long long)(unsigned)u * D.magic) >> D.shift);
    unsigned m = ((2u << (khi & 31u)) - 1u) & (0xFFFFFFFFu << (klo & 31u)) & D.valid11;
    m = t <= F.tmax ? m : 0u;
    return (m & D.lowmask) | ((m >> (D.win + 1)) << D.win);
}

// The 2*win permuted windows of both anchors (callCisLoops.py:179-198): bounds into S.aw / S.bw / S.wlo / S.whi, their
// positions in the two sorted lists into S.pos (one search per lane and round), the closed-form mask parameters into D;
// returns whether the closed form may be used.
__device__ __forceinline__ bool windows_setup(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int win,
                                              int lane, WarpAcc &S, WinDiv &D) {
    const int W = 2 * win;
    bool regular = ix.small != 0;
    {
        const long long ca4 = 2 * (xs + xe), cb4 = 2 * (ys + ye), sa4 = 2 * (xe - xs), sb4 = 2 * (ye - ys);
        const long long step4 = (xe - xs) + (ye - ys);
        if (lane < W) {
            // windows i = -win..-1, 1..win (callCisLoops.py:191-198), all bounds scaled by 4
            const int i = lane < win ? lane - win : lane - win + 1;
            long long al = ca4 + i * step4 - sa4, ah = ca4 + i * step4 + sa4;
            long long bl = cb4 + i * step4 - sb4, bh = cb4 + i * step4 + sb4;
            if (al < 0) al = 0;
            if (ah < 0) ah = 0;
            if (bl < 0) bl = 0;
            if (bh < 0) bh = 0;
            const Iv a = iv4(al, ah), b = iv4(bl, bh);
            const Iv32 a32 = iv32(a), b32 = iv32(b);
            S.aw[lane] = make_int2(a32.lo, (int)a32.wd);
            S.bw[lane] = make_int2(b32.lo, (int)b32.wd);
            S.wlo[lane] = a.lo; S.whi[lane] = a.hi;
            S.wlo[W + lane] = b.lo; S.whi[W + lane] = b.hi;
        }
        // closed-form masks: every lane derives the same family parameters
        const long long basea = ca4 - win * step4 - sa4, baseb = cb4 - win * step4 - sb4;
        const long long tmaxa = 2 * win * step4 + 2 * sa4, tmaxb = 2 * win * step4 + 2 * sb4;
        const long long lim = 1ll << 30;
        if (step4 < 1 || step4 >= lim || basea + 2 * sa4 < 0 || baseb + 2 * sb4 < 0 || basea <= -lim || baseb <= -lim ||
            basea >= lim || baseb >= lim || tmaxa >= lim || tmaxb >= lim)
            regular = false;
        D.a.base4 = (int)basea; D.a.tmax = (unsigned)tmaxa; D.a.w2 = (unsigned)(2 * sa4);
        D.b.base4 = (int)baseb; D.b.tmax = (unsigned)tmaxb; D.b.w2 = (unsigned)(2 * sb4);
        const unsigned d = (unsigned)(step4 < 1 ? 1 : (step4 >= lim ? 1 : step4));
        int l = 0;
        while ((1u << l) < d) l++;                           // d < 2^30
        D.shift = 31 + l;
        D.magic = (unsigned)(((1ull << (31 + l)) + d - 1) / d);
        D.s4m1 = d - 1;
        D.win = win;
        D.lowmask = (1u << win) - 1u;
        D.valid11 = ((2u << (2 * win)) - 1u) & ~(1u << win);
    }
    __syncwarp();
    // 2W windows x 2 lists x 2 bounds: one search per lane and round
    for (int t = lane; t < 8 * W; t += 32) {
        const int w = t >> 2;
        S.pos[t & 3][w] = bound_task(ix, S.wlo[w], S.whi[w], (t >> 1) & 1, t & 1);
    }
    __syncwarp();
    return regular;
}

// getPerRegions + getLoopNearbyPETs counts into the warp's accumulator (callCisLoops.py:173-223), W = 2*win <= 10.
// With ax / ay / bx / by the masks of the a- and b-windows holding a row's x and y (ma = ax | ay, mb = bx | by):
//     na[i]     = cX(a_i) + cY(a_i) - sum [ax_i and ay_i]
//     nab[i][j] = cY(a_i & b_j) + sum ([ma_i and mb_j] - [ay_i and by_j])       sums over the rows with x in any window
// One X-range scan; per batch of 32 rows the three 10-bit fields (ma, mb, ax & ay) and (ay, by, bx & by) are
// transposed across the warp so that lane t owns the ballot of bit t, and every lane accumulates its share of the
// W*W pair counters in registers from shuffled columns -- no shared-memory atomics, no data-dependent loops.
// fdr_stop > 0: the scan may stop as soon as fdr_stop window pairs are known to hold more than `rab` rows (the
// counters only grow), which decides the reference's `fdr >= 0.1` cut (callCisLoops.py:322); returns false then.
// Returns 1 when the counts are complete, 0 when the fdr cut was decided early, -1 when a single warp was asked for more
// than BIG_ROWS rows (nothing scanned: the candidate goes to the cooperative kernel).
__device__ int loop_windows10(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int win, int lane,
                              const Coop &C, WarpAcc &S, long long *g_bytes, int fdr_stop = 0, long long rab = 0) {
    const int W = 2 * win;
    WinDiv D;
    const bool regular = windows_setup(ix, xs, xe, ys, ye, win, lane, S, D);
    // X ranges of the two spans (window centres are monotone in i, so are the positions)
    uint32_t b0 = S.pos[0][0], e0 = S.pos[1][W - 1], b1 = S.pos[0][W], e1 = S.pos[1][2 * W - 1];
    const long long span_rows = (long long)max(e0 > b0 ? e0 - b0 : 0u, S.pos[3][W - 1] > S.pos[2][0] ? S.pos[3][W - 1] - S.pos[2][0] : 0u) +
                                (long long)max(e1 > b1 ? e1 - b1 : 0u, S.pos[3][2 * W - 1] > S.pos[2][W] ? S.pos[3][2 * W - 1] - S.pos[2][W] : 0u);
    if (e0 < b0) e0 = b0;
    if (e1 < b1) e1 = b1;
    if (b0 <= e1 && b1 <= e0) {
        b0 = b0 < b1 ? b0 : b1;
        e0 = e0 > e1 ? e0 : e1;
        b1 = e1 = 0;
    }
    const long long scan_rows = (long long)(e0 - b0) + (long long)(e1 - b1);
    if (C.defer && scan_rows > BIG_ROWS) return -1;
    *g_bytes += (2 * W + 1) * 32 * (long long)ix.lg;
    // this lane's pairs: lanes 0..29 own a-window lane / 3 and up to four b-windows from 4 * (lane % 3)
    const int pi = lane / 3, pj0 = 4 * (lane - 3 * pi);
    const bool own = lane < 30 && pi < W;
    int cy[4];                           // cY(a_i & b_j) of the lane's pairs
#pragma unroll
    for (int q = 0; q < 4; q++) {
        cy[q] = 0;
        if (own && pj0 + q < W) {
            const int i = pi, j = W + pj0 + q;
            const uint32_t yb = S.pos[2][i] > S.pos[2][j] ? S.pos[2][i] : S.pos[2][j];
            const uint32_t ye_ = S.pos[3][i] < S.pos[3][j] ? S.pos[3][i] : S.pos[3][j];
            cy[q] = ye_ > yb ? (int)(ye_ - yb) : 0;
        }
    }
    const int srca = own ? pi : 0;
    int acc[4] = {0, 0, 0, 0};
    int e_a = 0, e_b = 0;                // lanes 20 .. 20+W-1: sum [ax_i and ay_i], sum [bx_j and by_j]
    const int2 *__restrict__ byx = ix.byx;
    bool stopped = false;
    long long done_rows = 0;
    // cooperating warps: C.red[i*10 + j] collects the pair counters (what each warp has already added is in sent[]),
    // C.red[100..] / [110..] the e_a / e_b sums at the end, C.red[127] the stop flag
    int sent[4] = {0, 0, 0, 0};
    if (C.nw > 1) {
        __syncthreads();
        for (int t = threadIdx.x; t < COOP_RED; t += blockDim.x) C.red[t] = 0;
        __syncthreads();
    }
#pragma unroll 1
    for (int rr = 0; rr < 2 && !stopped; rr++) {
        const uint32_t rb = rr ? b1 : b0, re = rr ? e1 : e0;
        const uint32_t stride = 32u * (uint32_t)C.nw;
        const uint32_t nbatch = (re - rb + 31u) >> 5;
        const uint32_t iters = (nbatch + (uint32_t)C.nw - 1u) / (uint32_t)C.nw;      // the same for every cooperating warp
        const uint32_t first = rb + 32u * (uint32_t)C.part;
        int2 pnext = (first + lane < re) ? byx[first + lane] : make_int2(0, 0);
        int since = 0;
        for (uint32_t it = 0; it < iters; it++) {
            const uint32_t p0 = first + it * stride;
            const bool valid = p0 + lane < re;
            const int2 P = pnext;
            pnext = (p0 + stride + lane < re) ? byx[p0 + stride + lane] : make_int2(0, 0);
            unsigned ax, ay, bx, by;
            if (regular) {
                ax = div_mask(D.a, D, P.x); ay = div_mask(D.a, D, P.y);
                bx = div_mask(D.b, D, P.x); by = div_mask(D.b, D, P.y);
            } else {
                ax = win_mask(S.aw, W, P.x); ay = win_mask(S.aw, W, P.y);
                bx = win_mask(S.bw, W, P.x); by = win_mask(S.bw, W, P.y);
            }
            unsigned w1 = (ax | ay) | ((bx | by) << 10) | ((ax & ay) << 20);
            w1 = valid ? w1 : 0u;
            if (__any_sync(0xffffffffu, w1 != 0)) {          // else: rows in the gaps between windows
                unsigned w2 = ay | (by << 10) | ((bx & by) << 20);
                w2 = valid ? w2 : 0u;
                const unsigned c1 = warp_transpose32(w1, lane), c2 = warp_transpose32(w2, lane);
                e_a += __popc(c1);                           // meaningful in lanes 20..29 only
                e_b += __popc(c2);
                const unsigned A = __shfl_sync(0xffffffffu, c1, srca), Cc = __shfl_sync(0xffffffffu, c2, srca);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int sj = 10 + ((pj0 + q) < 10 ? pj0 + q : 9);
                    const unsigned B = __shfl_sync(0xffffffffu, c1, sj), Dd = __shfl_sync(0xffffffffu, c2, sj);
                    acc[q] += __popc(A & B & ~(Cc & Dd));
                }
            }
            if (fdr_stop <= 0) continue;
            if (C.nw == 1) {
                if (++since == 4) {
                    since = 0;
                    int over = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) over += (own && pj0 + q < W && (long long)acc[q] + cy[q] > rab) ? 1 : 0;
                    if (warp_sum(over) >= fdr_stop) {
                        stopped = true;
                        done_rows += (long long)(p0 + 32 - rb);
                        break;
                    }
                }
            } else if ((it & 63u) == 63u) {
                // every warp is at the same iteration: pool what was counted since the last check and let the totals decide
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (own && pj0 + q < W && acc[q] != sent[q]) {
                        atomicAdd(&C.red[pi * 10 + pj0 + q], acc[q] - sent[q]);
                        sent[q] = acc[q];
                    }
                __syncthreads();
                if (C.part == 0) {
                    int over = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        over += (own && pj0 + q < W && (long long)C.red[pi * 10 + pj0 + q] + cy[q] > rab) ? 1 : 0;
                    if (warp_sum(over) >= fdr_stop && lane == 0) C.red[127] = 1;
                }
                __syncthreads();
                if (C.red[127]) {
                    stopped = true;
                    const long long d = (long long)(it + 1) * stride;
                    done_rows += d < (long long)(re - rb) ? d : (long long)(re - rb);
                    break;
                }
            }
        }
        if (!stopped) done_rows += (long long)(re - rb);
    }
    // rows of the regions actually needed: all of them, or the part scanned before the fdr cut was decided
    *g_bytes += 16 * (stopped && scan_rows > 0 ? (span_rows * done_rows) / scan_rows : span_rows);
    if (stopped) return 0;
    if (C.nw > 1) {
        // what is left of the partial counters -> shared slots, then every warp takes the totals
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (own && pj0 + q < W && acc[q] != sent[q]) atomicAdd(&C.red[pi * 10 + pj0 + q], acc[q] - sent[q]);
        if (lane >= 20 && lane < 20 + W) {
            if (e_a) atomicAdd(&C.red[100 + lane - 20], e_a);
            if (e_b) atomicAdd(&C.red[110 + lane - 20], e_b);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) acc[q] = (own && pj0 + q < W) ? C.red[pi * 10 + pj0 + q] : 0;
        if (lane >= 20 && lane < 30) { e_a = C.red[100 + lane - 20]; e_b = C.red[110 + lane - 20]; }
    }
    const int src = 20 + (lane < W ? lane : 0);
    const int ea_i = __shfl_sync(0xffffffffu, e_a, src), eb_i = __shfl_sync(0xffffffffu, e_b, src);
    if (lane < W) {
        S.na[lane] = (int)(((long long)S.pos[1][lane] - S.pos[0][lane]) + ((long long)S.pos[3][lane] - S.pos[2][lane]) - ea_i);
        S.nb[lane] = (int)(((long long)S.pos[1][W + lane] - S.pos[0][W + lane]) +
                           ((long long)S.pos[3][W + lane] - S.pos[2][W + lane]) - eb_i);
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (own && pj0 + q < W) S.nab[pi * W + pj0 + q] = acc[q] + cy[q];
    __syncwarp();
    return 1;
}

// ---- the same counts for 2*win > 10 (up to 16): four range scans with de-duplication --------------------------
// getPerRegions + getLoopNearbyPETs counts into the warp's accumulator (callCisLoops.py:173-223)
__device__ void loop_windows_wide(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int win, int lane,
                             WarpAcc &S, long long *g_bytes) {
    const int W = 2 * win;
    for (int t = lane; t < W * W; t += 32) S.nab[t] = 0;
    long long a_lo = 0, a_hi = -1, b_lo = 0, b_hi = -1;
    int na_r = 0, nb_r = 0;
    if (lane < W) {
        // windows i = -win..-1, 1..win (callCisLoops.py:191-198), all bounds scaled by 4
        int i = lane < win ? lane - win : lane - win + 1;
        long long ca4 = 2 * (xs + xe), cb4 = 2 * (ys + ye), sa4 = 2 * (xe - xs), sb4 = 2 * (ye - ys);
        long long step4 = (xe - xs) + (ye - ys);
        long long al = ca4 + i * step4 - sa4, ah = ca4 + i * step4 + sa4;
        long long bl = cb4 + i * step4 - sb4, bh = cb4 + i * step4 + sb4;
        if (al < 0) al = 0;
        if (ah < 0) ah = 0;
        if (bl < 0) bl = 0;
        if (bh < 0) bh = 0;
        Iv a = iv4(al, ah), b = iv4(bl, bh);
        a_lo = a.lo; a_hi = a.hi; b_lo = b.lo; b_hi = b.hi;
        const Iv32 a32 = iv32(a), b32 = iv32(b);
        S.aw[lane] = make_int2(a32.lo, (int)a32.wd);
        S.bw[lane] = make_int2(b32.lo, (int)b32.wd);
        S.ahi[lane] = a32.lo == INT_MAX ? -1 : (int)(a32.lo + a32.wd);
        S.bhi[lane] = b32.lo == INT_MAX ? -1 : (int)(b32.lo + b32.wd);
    }
    // spans covering all a-windows / all b-windows (window centres are monotone in i)
    Iv SA, SB;
    SA.lo = __shfl_sync(0xffffffffu, a_lo, 0); SA.hi = __shfl_sync(0xffffffffu, a_hi, W - 1);
    SB.lo = __shfl_sync(0xffffffffu, b_lo, 0); SB.hi = __shfl_sync(0xffffffffu, b_hi, W - 1);
    const Iv32 SA32 = iv32(SA), SB32 = iv32(SB);
    __syncwarp();
    Range rxa = range_of(arr_x(ix), ix.n, SA, lane), rya = range_of(arr_y(ix), ix.n, SA, lane);
    Range rxb = range_of(arr_x(ix), ix.n, SB, lane), ryb = range_of(arr_y(ix), ix.n, SB, lane);
    {
        const long long ma_ = max(rxa.e - rxa.b, rya.e - rya.b), mb_ = max(rxb.e - rxb.b, ryb.e - ryb.b);
        *g_bytes += 16 * (ma_ + mb_) + (2 * W + 1) * 32 * (long long)ix.lg;
    }
    // four lists; a PET is handled by the first list that contains it.  Each list is sorted by the coordinate that put
    // the PET into it ("primary": x for the X-sorted lists, y for the Y-sorted ones), so the 32 PETs of a batch can
    // only fall into the few windows [t_lo, t_hi] that meet [first, last] of the batch -- found with two ballots over
    // the window bounds.  The other three coordinate / window-set combinations are rare and go through a span
    // pre-test and coord_mask().
    for (int pass = 0; pass < 4; pass++) {
        const int2 *arr = (pass & 1) ? ix.byy : ix.byx;
        const Range r = pass == 0 ? rxa : pass == 1 ? rya : pass == 2 ? rxb : ryb;
        const bool prim_a = pass < 2;                       // primary coordinate is tested against the a-windows
        const int2 *pw = prim_a ? S.aw : S.bw, *ow = prim_a ? S.bw : S.aw;
        const int *phi = prim_a ? S.ahi : S.bhi;
        const Iv32 PS = prim_a ? SA32 : SB32, OS = prim_a ? SB32 : SA32;   // spans of the primary / the other set
        int2 pnext = (r.b + lane < r.e) ? arr[r.b + lane] : make_int2(0, 0);
        for (int64_t i0 = r.b; i0 < r.e; i0 += 32) {
            const int64_t i = i0 + lane;
            const bool valid = i < r.e;
            const int2 p = pnext;
            pnext = (i + 32 < r.e) ? arr[i + 32] : make_int2(0, 0);     // next batch in flight while this one is tested
            const int prim = p.x;                            // first component = the sort coordinate of this list
            const int64_t left = r.e - i0;
            const int first = __shfl_sync(0xffffffffu, prim, 0);
            const int last = __shfl_sync(0xffffffffu, prim, left >= 32 ? 31 : (int)left - 1);
            const unsigned bl = __ballot_sync(0xffffffffu, lane < W && phi[lane < W ? lane : 0] >= first);
            const unsigned bh = __ballot_sync(0xffffffffu, lane < W && pw[lane < W ? lane : 0].x <= last);
            unsigned ma = 0, mb = 0;
            if (valid) {
                const int x = (pass & 1) ? p.y : p.x, y = (pass & 1) ? p.x : p.y;
                bool dup = false;
                if (pass >= 1 && in32(SA32, x)) dup = true;                      // seen in list 0
                if (pass >= 2 && in32(SA32, y)) dup = true;                      // seen in list 1
                if (pass == 3 && in32(SB32, x)) dup = true;                      // seen in list 2
                if (!dup) {
                    unsigned mp = 0;
                    if (bl && bh) {
                        const int t_lo = __ffs(bl) - 1, t_hi = 31 - __clz(bh);
                        for (int t = t_lo; t <= t_hi; t++) {
                            const int2 q = pw[t];
                            mp |= ((unsigned)(prim - q.x) <= (unsigned)q.y) ? (1u << t) : 0u;
                        }
                    }
                    const int sec = p.y;                    // the other coordinate
                    unsigned mo = 0;
                    if (in32(PS, sec)) mp |= coord_mask(pw, W, sec);
                    if (in32(OS, prim)) mo |= coord_mask(ow, W, prim);
                    if (in32(OS, sec)) mo |= coord_mask(ow, W, sec);
                    ma = prim_a ? mp : mo;
                    mb = prim_a ? mo : mp;
                }
            }
            if (__any_sync(0xffffffffu, (ma | mb) != 0)) window_accumulate(S, W, ma, mb, lane, na_r, nb_r);
        }
    }
    if (lane < W) { S.na[lane] = na_r; S.nb[lane] = nb_r; }
    __syncwarp();
}

__device__ __forceinline__ int loop_windows(const IdxView &ix, long long xs, long long xe, long long ys, long long ye,
                                            int win, int lane, const Coop &C, WarpAcc &S, long long *g_bytes,
                                            int fdr_stop = 0, long long rab = 0) {
    if (2 * win <= 10) return loop_windows10(ix, xs, xe, ys, ye, win, lane, C, S, g_bytes, fdr_stop, rab);
    if (C.part == 0) loop_windows_wide(ix, xs, xe, ys, ye, win, lane, S, g_bytes);       // not shared between warps
    return 1;
}

// ---- wavelet matrix: 2-D range counts without scanning ---------------------------------------------------------
// Every count of the path is a combination of 1-D range counts and rectangle counts R(I, J) = #{x in I and y in J}
// (checked against the oracle for ordered, unordered, overlapping and clipped intervals):
//     |P(I)|        = cX(I) + cY(I) - R(I, I)
//     |P(U) & P(V)| = cX(UV) + cY(UV) + R(U,V) + R(V,U) - R(U,UV) - R(V,UV) - R(UV,U) - R(UV,V) + R(UV,UV),  UV = U & V
// The scans above find R by visiting the rows of an X range, which is the cheapest way while a range holds a few
// hundred rows (Hi-TrAC / ChIA-PET like data).  In Hi-C like data a 10 kb window holds thousands of rows and the +-5
// step region of a candidate 10^5: there R(I, J) = F(p, q, J.hi + 1) - F(p, q, J.lo) with [p, q) the X range of I and
// F(p, q, v) = #{rows in [p, q) with y < v} comes from a wavelet matrix over the y values in x order in wm_L steps,
// each two 8-byte reads -- independent of the number of rows.
// gpu-scope descriptor access for the chained scan below (coherent in L2, never served from L1)
__device__ __forceinline__ unsigned wm_ld(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wm_st(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#define WM_AGG 0x40000000u
#define WM_INC 0x80000000u
#define WM_VAL 0x3FFFFFFFu
#define WM_NT 256
#define WM_IPT 8
#define WM_TILE (WM_NT * WM_IPT)

// y values in x order + the number of ones of the top level
__global__ void __launch_bounds__(256) k_wm_init(const int2 *__restrict__ byx, uint32_t *__restrict__ cur, int64_t n, int shift,
                                                  unsigned long long *__restrict__ ones0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bit = 0;
    if (i < n) {
        const uint32_t y = (uint32_t)byx[i].y;
        cur[i] = y;
        bit = (y >> shift) & 1u;
    }
    const int c = __popc(__ballot_sync(0xffffffffu, bit));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(ones0, (unsigned long long)c);
}
// One level in one pass: the level's {bits, ones before} table, the stable partition by the bit (zeros first) and the
// number of ones of the NEXT level (counted while the values pass through).  The ones before a tile come from a chained
// scan over the tiles (ticket order, decoupled look-back by the first warp); ones_here = this level's total, known from
// the previous pass, gives the start of the ones.
__global__ void __launch_bounds__(WM_NT) k_wm_level(const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, uint32_t n,
                                                     int shift, uint32_t nw, uint2 *__restrict__ level,
                                                     const unsigned long long *__restrict__ ones_here,
                                                     unsigned long long *__restrict__ ones_next, unsigned *__restrict__ desc,
                                                     unsigned *__restrict__ ticket) {
    __shared__ unsigned s_tile, s_excl;
    __shared__ int wtot[WM_NT / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const uint32_t i0 = tile * (uint32_t)WM_TILE + (uint32_t)w * (32 * WM_IPT) + lane;
    uint32_t v[WM_IPT];
    unsigned word[WM_IPT];
    int mine = 0;
#pragma unroll
    for (int r = 0; r < WM_IPT; r++) {
        const uint32_t i = i0 + r * 32;
        v[r] = i < n ? cur[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < WM_IPT; r++) {
        word[r] = __ballot_sync(0xffffffffu, (i0 + r * 32 < n) && ((v[r] >> shift) & 1u));
        mine += __popc(word[r]);
    }
    if (lane == 0) wtot[w] = mine;
    __syncthreads();
    if (w == 0) {
        int tot = 0;
#pragma unroll
        for (int k = 0; k < WM_NT / 32; k++) tot += wtot[k];
        if (lane == 0) wm_st(&desc[tile], (unsigned)tot | WM_AGG);
        // look back: lane j reads the descriptor of tile t - j; the nearest tiles are summed up to the first inclusive one
        unsigned excl = 0;
        int t = (int)tile - 1;
        while (t >= 0) {
            const int tt = t - lane;
            const unsigned d = tt >= 0 ? wm_ld(&desc[tt]) : WM_INC;
            const unsigned inc = __ballot_sync(0xffffffffu, (d & WM_INC) != 0);
            const unsigned nr = __ballot_sync(0xffffffffu, (d & (WM_INC | WM_AGG)) == 0);
            const int first_inc = inc ? __ffs(inc) - 1 : 32, first_nr = nr ? __ffs(nr) - 1 : 32;
            const int upto = first_inc < first_nr ? first_inc : first_nr;          // lanes below `upto` hold aggregates
            const bool take = lane < upto || (lane == upto && first_inc == upto);
            excl += __reduce_add_sync(0xffffffffu, take ? (d & WM_VAL) : 0u);
            if (first_inc == upto && upto < 32) break;                             // reached an inclusive prefix
            t -= upto;                                                             // not published yet: read again from there
        }
        if (lane == 0) {
            wm_st(&desc[tile], (excl + (unsigned)tot) | WM_INC);
            s_excl = excl;
        }
    }
    __syncthreads();
    unsigned before = s_excl;
    for (int k = 0; k < w; k++) before += (unsigned)wtot[k];
    const uint32_t z = n - (uint32_t)*ones_here;
    const unsigned lt = (1u << lane) - 1u;
    int nb = 0;
#pragma unroll
    for (int r = 0; r < WM_IPT; r++) {
        const uint32_t i = i0 + r * 32;
        if (lane == 0 && (i >> 5) < nw) level[i >> 5] = make_uint2(word[r], before);
        if (i < n) {
            const uint32_t r1 = before + __popc(word[r] & lt);
            nxt[((word[r] >> lane) & 1u) ? z + r1 : i - r1] = v[r];
            if (shift > 0) nb += (v[r] >> (shift - 1)) & 1u;
        }
        before += __popc(word[r]);
    }
    if (shift > 0) {
        nb = __reduce_add_sync(0xffffffffu, nb);
        if (lane == 0 && nb) atomicAdd(ones_next, (unsigned long long)nb);
    }
}

static int cl2_index_wavelet(cl2_ctx *ctx, cl2_index *idx) {
    if (idx->wm || idx->n <= 0) return CL2_OK;
    const int64_t n = idx->n;
    const int L = bits_for((uint64_t)idx->ymax);
    const int64_t nw = n / 32 + 1;
    if (L > 32 || n >= (int64_t)WM_VAL) return cl2_fail(ctx, CL2_ERR_RANGE, "wavelet matrix: too many rows");
    const int ntiles = (int)((nw * 32 + WM_TILE - 1) / WM_TILE);
    Scratch sc(ctx);
    uint32_t *cur, *nxt;
    unsigned long long *tot;          // ones of every level
    unsigned *desc, *ticket;
    CL2_TRY(sc.get(&cur, (size_t)n));
    CL2_TRY(sc.get(&nxt, (size_t)n));
    CL2_TRY(sc.get(&tot, (size_t)L + 1));
    CL2_TRY(sc.get(&desc, (size_t)L * (size_t)ntiles + (size_t)L));
    ticket = desc + (size_t)L * (size_t)ntiles;
    uint2 *wm;
    CL2_TRY(cl2_dev_alloc(ctx, &wm, (size_t)L * (size_t)nw));
    CL2_CUDA(ctx, cudaMemsetAsync(tot, 0, ((size_t)L + 1) * 8, ctx->stream));
    CL2_CUDA(ctx, cudaMemsetAsync(desc, 0, ((size_t)L * (size_t)ntiles + (size_t)L) * 4, ctx->stream));
    {
        FamTimer t(ctx, FAM_INDEX, 12.0 * (double)n);
        k_wm_init<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(idx->byx, cur, n, L - 1, tot);
    }
    for (int l = 0; l < L; l++) {
        FamTimer t(ctx, FAM_INDEX, 8.25 * (double)n);
        k_wm_level<<<ntiles, WM_NT, 0, ctx->stream>>>(cur, nxt, (uint32_t)n, L - 1 - l, (uint32_t)nw, wm + (size_t)l * (size_t)nw,
                                                       tot + l, tot + l + 1, desc + (size_t)l * (size_t)ntiles, ticket + l);
        uint32_t *t2 = cur; cur = nxt; nxt = t2;
    }
    std::vector<unsigned long long> ht((size_t)L);
    CL2_CUDA(ctx, cudaMemcpyAsync(ht.data(), tot, (size_t)L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CL2_CUDA(ctx, cudaGetLastError());
    for (int l = 0; l < L; l++) idx->wm_z[l] = (uint32_t)(n - (int64_t)ht[(size_t)l]);
    idx->wm = wm;
    idx->wm_L = L;
    idx->wm_nw = nw;
    return CL2_OK;
}

// F(p, q, vv) = #{rows at x positions [p, q) with y < vv}, one lane on its own query
__device__ __forceinline__ int wm_less(const IdxView &ix, uint32_t p, uint32_t q, long long vv) {
    if (q <= p || vv <= 0) return 0;
    const int L = ix.wm_L;
    if (vv >= (1ll << L)) return (int)(q - p);
    const uint32_t v = (uint32_t)vv;
    int cnt = 0;
    const uint2 *lev = ix.wm;
    for (int l = 0; l < L; l++, lev += ix.wm_nw) {
        const uint2 ep = lev[p >> 5], eq = lev[q >> 5];
        const uint32_t r1p = ep.y + __popc(ep.x & ((1u << (p & 31)) - 1u));
        const uint32_t r1q = eq.y + __popc(eq.x & ((1u << (q & 31)) - 1u));
        if ((v >> (L - 1 - l)) & 1u) {
            cnt += (int)((q - r1q) - (p - r1p));             // the rows with a 0 here are smaller
            p = ix.wm_z[l] + r1p;
            q = ix.wm_z[l] + r1q;
        } else {
            p -= r1p;
            q -= r1q;
        }
        if (p == q) break;
    }
    return cnt;
}

struct WmTab {                 // per-warp tables of the wavelet form of the window counts
    int T[2 * 10][4 * 10];     // T[w][t] = F(X range of window w, threshold t); t < 2W: lo of window t, else hi + 1 of t - 2W
    int O[100][4];             // per window pair (i, j): F(X range of a_i & b_j, {a.lo, a.hi + 1, b.lo, b.hi + 1})
};

// |P(U)|, |P(V)|-style terms for the four core intervals A, B, A2, B2 in one round of 32 queries:
// lanes 0-1 R(A,A), 2-3 R(B,B), 4-17 the seven rectangles of (A, B), 18-31 those of (A2, B2)
__device__ void core_wm(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int mode, int lane,
                        int64_t out[4], long long *g_bytes) {
    const int k = (lane >> 2) & 3;
    long long lo, hi;
    if (k == 0) { lo = xs; hi = xe; }
    else if (k == 1) { lo = ys; hi = ye; }
    else if (k == 2) {
        if (mode == CL2_MODE_CIS) { lo = xe; hi = xe + (xe - xs); }
        else { lo = xs - (xe - xs); hi = xs; }
    } else { lo = ys - (ye - ys); hi = ys; }
    const uint32_t pos = lane < 16 ? bound_task(ix, lo, hi, (lane >> 1) & 1, lane & 1) : 0u;
    IvPos P[4];
    Iv I[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        P[q] = gather_pos(pos, q);
        I[q].lo = __shfl_sync(0xffffffffu, lo, 4 * q);
        I[q].hi = __shfl_sync(0xffffffffu, hi, 4 * q);
    }
    // this lane's query: X range [p, q), threshold, sign, destination (0 R(A,A), 1 R(B,B), 2 pair(A,B), 3 pair(A2,B2))
    uint32_t p = 0, q = 0;
    long long thr = 0;
    int sign = 0, dst = 0;
    if (lane < 4) {
        const int w = lane >> 1;
        p = P[w].xb; q = P[w].xe;
        thr = (lane & 1) ? I[w].hi + 1 : I[w].lo;
        sign = (lane & 1) ? 1 : -1;
        dst = w;
    } else {
        const int g = lane < 18 ? 0 : 1, r = ((lane - 4) % 14) >> 1, side = (lane - 4) & 1;
        const Iv U = I[2 * g], V = I[2 * g + 1];
        const IvPos PU = P[2 * g], PV = P[2 * g + 1];
        Iv UV;
        UV.lo = U.lo > V.lo ? U.lo : V.lo;
        UV.hi = U.hi < V.hi ? U.hi : V.hi;
        IvPos PUV;
        PUV.xb = PU.xb > PV.xb ? PU.xb : PV.xb;
        PUV.xe = PU.xe < PV.xe ? PU.xe : PV.xe;
        // rectangles: 0 R(U,V) 1 R(V,U) 2 R(U,UV) 3 R(V,UV) 4 R(UV,U) 5 R(UV,V) 6 R(UV,UV); signs + + - - - - +
        const IvPos X = (r == 0 || r == 2) ? PU : ((r == 1 || r == 3) ? PV : PUV);
        const Iv J = (r == 0 || r == 5) ? V : ((r == 1 || r == 4) ? U : UV);
        p = X.xb; q = X.xe;
        const bool empty = J.lo > J.hi || (r >= 2 && UV.lo > UV.hi);
        if (empty) q = p;
        thr = side ? J.hi + 1 : J.lo;
        sign = (side ? 1 : -1) * ((r == 0 || r == 1 || r == 6) ? 1 : -1);
        dst = 2 + g;
    }
    const int f = sign * wm_less(ix, p, q, thr);
    int r4[4];
#pragma unroll
    for (int d = 0; d < 4; d++) r4[d] = warp_sum(dst == d ? f : 0);
    out[0] = ((int64_t)P[0].xe - P[0].xb) + ((int64_t)P[0].ye - P[0].yb) - r4[0];
    out[1] = ((int64_t)P[1].xe - P[1].xb) + ((int64_t)P[1].ye - P[1].yb) - r4[1];
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const IvPos PU = P[2 * g], PV = P[2 * g + 1];
        const long long xb = PU.xb > PV.xb ? PU.xb : PV.xb, xe_ = PU.xe < PV.xe ? PU.xe : PV.xe;
        out[2 + g] = (xe_ > xb ? xe_ - xb : 0) + ycount_both(PU, PV) + r4[2 + g];
    }
    *g_bytes += 16 * (pos_max(P[0]) + pos_max(P[1]) + pos_max(P[2]) + pos_max(P[3])) + 4 * 32 * (long long)ix.lg;
}

// The window counts of loop_windows10 from rectangle counts: S.na, S.nb, S.nab as that function leaves them.
__device__ void windows_wm(const IdxView &ix, long long xs, long long xe, long long ys, long long ye, int win, int lane,
                           WarpAcc &S, WmTab &Tb, long long *g_bytes) {
    const int W = 2 * win;
    WinDiv D;
    windows_setup(ix, xs, xe, ys, ye, win, lane, S, D);
    {
        const uint32_t b0 = S.pos[0][0], e0 = S.pos[1][W - 1], b1 = S.pos[0][W], e1 = S.pos[1][2 * W - 1];
        const long long span_rows = (long long)max(e0 > b0 ? e0 - b0 : 0u, S.pos[3][W - 1] > S.pos[2][0] ? S.pos[3][W - 1] - S.pos[2][0] : 0u) +
                                    (long long)max(e1 > b1 ? e1 - b1 : 0u, S.pos[3][2 * W - 1] > S.pos[2][W] ? S.pos[3][2 * W - 1] - S.pos[2][W] : 0u);
        *g_bytes += 16 * span_rows + (2 * W + 1) * 32 * (long long)ix.lg;
    }
    // T[w][t]: 2W windows x 4W thresholds
    for (int q = lane; q < 2 * W * 4 * W; q += 32) {
        const int w = q / (4 * W), t = q - w * 4 * W;
        const long long thr = t < 2 * W ? S.wlo[t] : S.whi[t - 2 * W] + 1;
        Tb.T[w][t] = wm_less(ix, S.pos[0][w], S.pos[1][w], thr);
    }
    // O[pair][k]: the X range of a_i & b_j against the four bounds of the pair (empty intersection: zeros)
    for (int q = lane; q < W * W * 4; q += 32) {
        const int pr = q >> 2, kk = q & 3, i = pr / W, j = W + pr % W;
        const uint32_t p = S.pos[0][i] > S.pos[0][j] ? S.pos[0][i] : S.pos[0][j];
        const uint32_t e = S.pos[1][i] < S.pos[1][j] ? S.pos[1][i] : S.pos[1][j];
        const long long thr = kk == 0 ? S.wlo[i] : kk == 1 ? S.whi[i] + 1 : kk == 2 ? S.wlo[j] : S.whi[j] + 1;
        Tb.O[pr][kk] = wm_less(ix, p, e > p ? e : p, thr);
    }
    __syncwarp();
    if (lane < W) {
        const int i = lane, j = W + lane;
        S.na[lane] = (int)(((long long)S.pos[1][i] - S.pos[0][i]) + ((long long)S.pos[3][i] - S.pos[2][i]) -
                           (Tb.T[i][2 * W + i] - Tb.T[i][i]));
        S.nb[lane] = (int)(((long long)S.pos[1][j] - S.pos[0][j]) + ((long long)S.pos[3][j] - S.pos[2][j]) -
                           (Tb.T[j][2 * W + j] - Tb.T[j][j]));
    }
    for (int pr = lane; pr < W * W; pr += 32) {
        const int i = pr / W, j = W + pr % W;
        // a = window i, b = window j; ab = a & b
        const long long alo = S.wlo[i], ahi = S.whi[i], blo = S.wlo[j], bhi = S.whi[j];
        const bool lo_from_a = alo >= blo, hi_from_a = ahi <= bhi;
        const long long ablo = lo_from_a ? alo : blo, abhi = hi_from_a ? ahi : bhi;
        const int tlo = lo_from_a ? i : j, thi = 2 * W + (hi_from_a ? i : j);          // thresholds of ab in T's columns
        long long v = (Tb.T[i][2 * W + j] - Tb.T[i][j]) + (Tb.T[j][2 * W + i] - Tb.T[j][i]);     // R(a,b) + R(b,a)
        if (ablo <= abhi) {
            const uint32_t xb = S.pos[0][i] > S.pos[0][j] ? S.pos[0][i] : S.pos[0][j];
            const uint32_t xe_ = S.pos[1][i] < S.pos[1][j] ? S.pos[1][i] : S.pos[1][j];
            const uint32_t yb = S.pos[2][i] > S.pos[2][j] ? S.pos[2][i] : S.pos[2][j];
            const uint32_t ye_ = S.pos[3][i] < S.pos[3][j] ? S.pos[3][i] : S.pos[3][j];
            v += (xe_ > xb ? (long long)(xe_ - xb) : 0) + (ye_ > yb ? (long long)(ye_ - yb) : 0);      // cX(ab) + cY(ab)
            v -= (Tb.T[i][thi] - Tb.T[i][tlo]) + (Tb.T[j][thi] - Tb.T[j][tlo]);                        // R(a,ab) + R(b,ab)
            const int *o = Tb.O[pr];
            v -= (o[1] - o[0]) + (o[3] - o[2]);                                                        // R(ab,a) + R(ab,b)
            v += (hi_from_a ? o[1] : o[3]) - (lo_from_a ? o[0] : o[2]);                                // R(ab,ab)
        }
        S.nab[pr] = (int)v;
    }
    __syncwarp();
}

// ---- raw counts (cl2_loop_counts) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CNT_WARPS * 32) k_loop_counts(const IdxView ix, const longlong4 *__restrict__ loops, int64_t L,
                                                                 int win, int mode, long long *__restrict__ core,
                                                                 long long *__restrict__ na_out, long long *__restrict__ nb_out,
                                                                 long long *__restrict__ nab_out,
                                                                 long long *__restrict__ anchors) {
    __shared__ WarpAcc acc[CNT_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * CNT_WARPS + w;
    if (k >= L) return;
    const longlong4 lp = loops[k];
    const long long xs = lp.x, xe = lp.y, ys = lp.z, ye = lp.w;
    long long scanned = 0;
    if (core) {
        int64_t c[4];
        loop_core(ix, xs, xe, ys, ye, mode, lane, coop_solo(), c, &scanned);
        if (lane == 0) { core[4 * k] = c[0]; core[4 * k + 1] = c[1]; core[4 * k + 2] = c[2]; core[4 * k + 3] = c[3]; }
    }
    if (anchors) {
        int64_t cx, wx, cy, wy;
        loop_anchors(ix, xs, xe, ys, ye, lane, coop_solo(), &cx, &wx, &cy, &wy, &scanned);
        if (lane == 0) { anchors[4 * k] = cx; anchors[4 * k + 1] = wx; anchors[4 * k + 2] = cy; anchors[4 * k + 3] = wy; }
    }
    if (na_out || nb_out || nab_out) {
        const int W = 2 * win;
        WarpAcc &S = acc[w];
        loop_windows(ix, xs, xe, ys, ye, win, lane, coop_solo(), S, &scanned);
        if (na_out && lane < W) na_out[(int64_t)W * k + lane] = S.na[lane];
        if (nb_out && lane < W) nb_out[(int64_t)W * k + lane] = S.nb[lane];
        if (nab_out)
            for (int t = lane; t < W * W; t += 32) nab_out[(int64_t)W * W * k + t] = S.nab[t];
    }
}

// ---- fused counting + statistics ----------------------------------------------------------------------------
// phase 1: ra, rb, rab, lower and the hypergeometric tail hypergeom.sf(rab-1, N, ra, rb) (callCisLoops.py:310)
__global__ void __launch_bounds__(CNT_WARPS * 32) k_loop_core(const IdxView ix, const longlong4 *__restrict__ loops, int64_t L,
                                                               int mode, long long *__restrict__ core,
                                                               double *__restrict__ hyp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * CNT_WARPS + w;
    if (k >= L) return;
    const longlong4 lp = loops[k];
    int64_t c[4];
    long long scanned = 0;
    loop_core(ix, lp.x, lp.y, lp.z, lp.w, mode, lane, coop_solo(), c, &scanned);
    if (lane == 0) {
        core[4 * k] = c[0]; core[4 * k + 1] = c[1]; core[4 * k + 2] = c[2]; core[4 * k + 3] = c[3];
        hyp[k] = cl2_hypergeom_sf_ge((double)c[2], (double)ix.n, (double)c[0], (double)c[1]);
    }
}

// numpy's pairwise summation of a contiguous float64 vector (np.mean / np.add.reduce), n <= 256
__device__ double np_pairwise_base(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}
__device__ double np_pairwise_sum(const double *a, int n) {
    if (n <= 128) return np_pairwise_base(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_base(a, n2) + np_pairwise_base(a + n2, n - n2);
}

struct WarpVals {
    double v[CNT_MAXW * CNT_MAXW];
};

// median of the m non-negative values in S.v (np.median: mean of the two middle order statistics), by selection:
// keep an open value bracket (lo, hi) around the wanted order statistic, take some element inside it as pivot, count
// the values below / not above the pivot, shrink.  The pivot choice only affects the number of rounds (about log m).
__device__ double warp_median(WarpVals &S, int m, int lane) {
    const int k1 = (m - 1) / 2, k2 = m / 2;
    double lo = -1.0, hi = CUDART_INF;
    double p = 0.0;
    int le = 0;
    for (int round = 0;; round++) {
        double cand = 0.0;
        bool has = false;
        for (int i = lane; i < m; i += 32) {
            const double y = S.v[i];
            if (!has && y > lo && y < hi) { cand = y; has = true; }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, has);     // never empty: the answer itself is inside
        const int rot = (round * 11 + 5) & 31;                    // vary the lane the pivot comes from
        const int src = (__ffs(__funnelshift_r(bal, bal, rot)) - 1 + rot) & 31;
        p = __shfl_sync(0xffffffffu, cand, src);
        int lt = 0;
        le = 0;
        for (int i = lane; i < m; i += 32) {
            const double y = S.v[i];
            lt += y < p ? 1 : 0;
            le += y <= p ? 1 : 0;
        }
        lt = warp_sum(lt);
        le = warp_sum(le);
        if (k1 < lt) hi = p;
        else if (k1 >= le) lo = p;
        else break;
    }
    double q = p;                                                  // order statistic k2 = k1 or k1 + 1
    if (k2 >= le) {
        q = CUDART_INF;                                            // smallest value above p
        for (int i = lane; i < m; i += 32) {
            const double y = S.v[i];
            if (y > p && y < q) q = y;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) q = fmin(q, __shfl_xor_sync(0xffffffffu, q, o));
    }
    return (p + q) / 2.0;
}

#define CL2_NSTAT 13
// Permuted-window background, enrichment tests and anchor tests of one loop (one warp); results into o[13] (shared):
// mrabs, mbps, fdr, poisson sf, binomial sf, px, esx, py, esy, count_x, wide_x, count_y, wide_y
// From the window counts in S (na, nb, nab) to mrabs, mbps, fdr and the background cut: the part of loop_background
// after the counting.  Returns 1 keep, 0 drop.
__device__ int background_finish(double rab, int win, int mode, WarpAcc &S, WarpVals &V, int lane, bool bg_cut, double pseudo,
                                 double *o_mrabs, double *o_mbps, double *o_fdr) {
    const int W = 2 * win, m = W * W;
    // rabs: counts of the 4*win^2 window pairs; fdr = #{rabs > rab} / m
    int gt = 0, nzero = 0;
    for (int t = lane; t < m; t += 32) {
        V.v[t] = (double)S.nab[t];
        gt += ((double)S.nab[t] > rab) ? 1 : 0;
        nzero += S.nab[t] == 0 ? 1 : 0;
    }
    gt = warp_sum(gt);
    nzero = warp_sum(nzero);
    __syncwarp();
    // cis medians: when more than half of the window pairs are empty both middle order statistics are 0, for the
    // counts and for the densities alike (a density is 0 exactly where the count is 0) -- the common case
    const bool zero_median = mode == CL2_MODE_CIS && nzero >= m / 2 + 1;
    const double fdr = (double)gt / (double)m;
    if (bg_cut && fdr >= 0.1) return 0;                    // the cheap half of the background cut first
    double mrabs, mbps;
    if (zero_median) mrabs = 0.0;
    else if (mode == CL2_MODE_CIS) mrabs = warp_median(V, m, lane);
    else mrabs = np_pairwise_sum(V.v, m) / (double)m;     // every lane computes the same value
    __syncwarp();
    if (bg_cut) {
        if (mrabs >= rab || fdr >= 0.1) return 0;
        if (rab / (mrabs > pseudo ? mrabs : pseudo) < 1.0) return 0;
    }
    if (zero_median) {
        mbps = 0.0;
    } else {
        // nbps: nrab / (nac * nbc), zero rule differs between cis (:213-221) and trans (callTransLoops.py:165-169)
        for (int t = lane; t < m; t += 32) {
            double nrab = (double)S.nab[t], nac = (double)S.na[t / W], nbc = (double)S.nb[t % W];
            double d;
            if (mode == CL2_MODE_CIS) d = nrab > 0 ? nrab / (nac * nbc) : 0.0;
            else d = (nac > 0 && nbc > 0) ? nrab / (nac * nbc) : 0.0;
            V.v[t] = d;
        }
        __syncwarp();
        if (mode == CL2_MODE_CIS) mbps = warp_median(V, m, lane);
        else mbps = np_pairwise_sum(V.v, m) / (double)m;
        __syncwarp();
    }
    *o_mrabs = mrabs; *o_mbps = mbps; *o_fdr = fdr;
    return 1;
}


// Background of one loop (one warp): permuted windows -> mrabs, mbps, fdr (callCisLoops.py:207-223, 315-321).
// bg_cut (cis): return false as soon as `mrabs >= rab or fdr >= 0.1` or `es < 1` (:322-327) is known; the reference
// applies these cuts whatever `pseudo` is, which only enters the enrichment score.
// Returns 1 keep, 0 drop, -1 hand the candidate to the cooperative kernel, -2 (cooperating warps other than the first)
// nothing left to do for this warp.
__device__ int loop_background(const IdxView &ix, long long xs,
                               long long xe, long long ys, long long ye, double rab, int win, int mode, const Coop &C,
                               WarpAcc &S, WarpVals &V, int lane, bool bg_cut, double pseudo, double *o_mrabs,
                               double *o_mbps, double *o_fdr, long long *g_scan) {
    const int W = 2 * win, m = W * W;
    // smallest number of window pairs above rab that makes fdr = gt / m >= 0.1 (the comparison the reference makes)
    int fdr_stop = 0;
    if (bg_cut) {
        fdr_stop = m / 10;
        while ((double)fdr_stop / (double)m < 0.1) fdr_stop++;
        while (fdr_stop > 1 && (double)(fdr_stop - 1) / (double)m >= 0.1) fdr_stop--;
    }
    const int wr = loop_windows(ix, xs, xe, ys, ye, win, lane, C, S, g_scan, fdr_stop, (long long)rab);
    if (wr <= 0) return wr;
    if (C.nw > 1) {
        __syncthreads();                                   // the counters in S are complete for every warp
        if (C.part != 0) return -2;                        // the medians are one warp's work
    }
    return background_finish(rab, win, mode, S, V, lane, bg_cut, pseudo, o_mrabs, o_mbps, o_fdr);
}

// Anchor windows and the tail probabilities of one loop (one warp): o[3..12] = poisson sf, binomial sf, px, esx, py,
// esy, count_x, wide_x, count_y, wide_y  (callCisLoops.py:226-247, 329-333)
__device__ void loop_tail(const IdxView &ix, long long xs,
                          long long xe, long long ys, long long ye, double ra, double rb, double rab, double mrabs,
                          double mbps, int mode, double rpb, int lane, double *o, long long *g_scan) {
    int64_t cx, wx, cy, wy;
    loop_anchors(ix, xs, xe, ys, ye, lane, coop_solo(), &cx, &wx, &cy, &wy, g_scan);
    if (lane == 0) {
        o[3] = cl2_poisson_sf_ge(rab, mrabs);                                            // :329
        o[9] = (double)cx; o[10] = (double)wx; o[11] = (double)cy; o[12] = (double)wy;
    } else if (lane == 1) {
        double nn = mode == CL2_MODE_CIS ? ra * rb - rab : ra * rb;                      // :332 / callTransLoops.py:181
        o[4] = cl2_binom_sf_ge(rab, nn, mbps);
    } else if (lane == 2 || lane == 3) {
        // estAnchorSig (:226-247): c = max(local mean of the +-5x window, global rate * length, 1)
        double count = (double)(lane == 2 ? cx : cy), wide = (double)(lane == 2 ? wx : wy);
        double len = (double)(lane == 2 ? xe - xs : ye - ys);
        double r = (wide - count) / 5.0 / 2.0;
        double c = fmax(fmax(r, rpb * len), 1.0);
        o[lane == 2 ? 5 : 7] = cl2_poisson_sf_ge(count, c);
        o[lane == 2 ? 6 : 8] = count / c;
    }
    __syncwarp();
}

// both parts (the two-phase API, cl2_loop_tests)
__device__ void loop_tests_warp(const IdxView &ix, long long xs,
                                long long xe, long long ys, long long ye, double ra, double rb, double rab, int win,
                                int mode, double rpb, WarpAcc &S, WarpVals &V, int lane, double *o, long long *g_scan) {
    double mrabs, mbps, fdr;
    loop_background(ix, xs, xe, ys, ye, rab, win, mode, coop_solo(), S, V, lane, false, 1.0, &mrabs, &mbps, &fdr, g_scan);
    if (lane == 0) { o[0] = mrabs; o[1] = mbps; o[2] = fdr; }
    loop_tail(ix, xs, xe, ys, ye, ra, rb, rab, mrabs, mbps, mode, rpb, lane, o, g_scan);
}

// phase 2 as a separate launch (cl2_loop_tests)
__global__ void __launch_bounds__(CNT_WARPS * 32) k_loop_tests(const IdxView ix, const longlong4 *__restrict__ loops,
                                                                const long long *__restrict__ core, int64_t L, int win,
                                                                int mode, double rpb, double *__restrict__ stats) {
    __shared__ WarpAcc acc[CNT_WARPS];
    __shared__ WarpVals vals[CNT_WARPS];
    __shared__ double ost[CNT_WARPS][CL2_NSTAT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * CNT_WARPS + w;
    if (k >= L) return;
    const longlong4 lp = loops[k];
    long long scanned = 0;
    loop_tests_warp(ix, lp.x, lp.y, lp.z, lp.w, (double)core[4 * k], (double)core[4 * k + 1],
                    (double)core[4 * k + 2], win, mode, rpb, acc[w], vals[w], lane, ost[w], &scanned);
    if (lane < CL2_NSTAT) stats[CL2_NSTAT * k + lane] = ost[w][lane];
}

// estLoopSig on the device, kernel A (every candidate, one warp each; integer counting + medians only, so the code
// stays small): the count filters of the reference's cascade (cis callCisLoops.py:282-307, trans callTransLoops.py:
// 127-139), the permuted-window background and, for cis, the background cut (:322-327).  Survivors get flag = 1 and
// core[4], stats[0..2] = mrabs, mbps, fdr.  The hypergeometric cut (:311) is applied by kernel B: the cascade is a
// conjunction, so the order of the tests does not change which candidates survive.
// One candidate by the warps of C: returns 1 keep, 0 drop, -1 too large for a single warp (nothing was counted).
__device__ int est_candidate(const IdxView &ix, const longlong4 lp, int win, int mode, int minPts, int hic, double cdc,
                             double pseudo, const Coop &C, WarpAcc &S, WarpVals &V, int lane, int64_t c[4], double st[3],
                             long long *bytes) {
    const long long xs = lp.x, xe = lp.y, ys = lp.z, ye = lp.w;
    const double la = (double)(xe - xs), lb = (double)(ye - ys);
    const bool cis = mode == CL2_MODE_CIS;
    if (la / lb > cdc || lb / la > cdc) return 0;                                      // :282-286 / trans :135-139
    if (!loop_core(ix, xs, xe, ys, ye, mode, lane, C, c, bytes)) return -1;
    const double ra = (double)c[0], rb = (double)c[1], rab = (double)c[2];
    if (c[2] < minPts) return 0;                                                       // :290 / trans :127
    if (cis && (c[0] == c[2] || c[1] == c[2])) return 0;                               // :293
    if (ra / rb > cdc || rb / ra > cdc) return 0;                                      // :296 / trans :133
    const double lower = (double)c[3];
    const double p2ll = rab / (lower > pseudo ? lower : pseudo);                       // :306
    if (cis && hic && p2ll < 1.0) return 0;                                            // :307
    return loop_background(ix, xs, xe, ys, ye, rab, win, mode, C, S, V, lane, cis, pseudo, &st[0], &st[1], &st[2], bytes);
}

template <int MINB>
__global__ void __launch_bounds__(CNT_WARPS * 32, MINB) k_est_bg(const IdxView ix, const longlong4 *__restrict__ loops, int64_t L,
                                                            int win, int mode, int minPts, int hic, double cdc,
                                                            double pseudo, long long *__restrict__ core_out,
                                                            double *__restrict__ stats_out, int32_t *__restrict__ flag,
                                                            unsigned long long *__restrict__ scanned_total,
                                                            int32_t *__restrict__ biglist, int32_t *__restrict__ nbig) {
    __shared__ WarpAcc acc[CNT_WARPS];
    __shared__ WarpVals vals[CNT_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * CNT_WARPS + w;
    if (k >= L) return;
    long long scanned = 0;
    int64_t c[4];
    double st[3];
    const Coop C = {1, 0, nullptr, true};
    const int r = est_candidate(ix, loops[k], win, mode, minPts, hic, cdc, pseudo, C, acc[w], vals[w], lane, c, st, &scanned);
    if (lane == 0) {
        if (r == 1) {
            core_out[4 * k] = c[0]; core_out[4 * k + 1] = c[1]; core_out[4 * k + 2] = c[2]; core_out[4 * k + 3] = c[3];
            stats_out[CL2_NSTAT * k] = st[0]; stats_out[CL2_NSTAT * k + 1] = st[1]; stats_out[CL2_NSTAT * k + 2] = st[2];
        }
        flag[k] = r == 1 ? 1 : 0;
        if (r < 0) biglist[atomicAdd(nbig, 1)] = (int32_t)k;        // decided by k_est_bg_big
        else if (scanned) atomicAdd(scanned_total, (unsigned long long)scanned);
    }
}

// The candidates kernel A declined (regions with more than BIG_ROWS rows), one block each: its warps share every scan.
// The blocks of a fixed-size grid stride over the list, whose length stays on the device.
template <int BIG_THREADS, int MINB>
__global__ void __launch_bounds__(BIG_THREADS, MINB) k_est_bg_big(const IdxView ix, const longlong4 *__restrict__ loops, int win, int mode,
                                                            int minPts, int hic, double cdc, double pseudo,
                                                            long long *__restrict__ core_out, double *__restrict__ stats_out,
                                                            int32_t *__restrict__ flag,
                                                            unsigned long long *__restrict__ scanned_total,
                                                            const int32_t *__restrict__ biglist,
                                                            const int32_t *__restrict__ nbig) {
    __shared__ WarpAcc acc;
    __shared__ WarpVals vals;
    __shared__ int red[COOP_RED];
    const int lane = threadIdx.x & 31;
    const Coop C = {BIG_THREADS / 32, (int)(threadIdx.x >> 5), red, false};
    const int n = *nbig;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        __syncthreads();                                   // the shared accumulators of the previous candidate are free
        const int64_t k = biglist[b];
        long long scanned = 0;
        int64_t c[4];
        double st[3];
        const int r = est_candidate(ix, loops[k], win, mode, minPts, hic, cdc, pseudo, C, acc, vals, lane, c, st, &scanned);
        // every warp sees the same outcome except in loop_background, where warps 1.. leave with -2 once the counting is
        // done: the first warp reports
        if (threadIdx.x == 0) {
            if (r == 1) {
                core_out[4 * k] = c[0]; core_out[4 * k + 1] = c[1]; core_out[4 * k + 2] = c[2]; core_out[4 * k + 3] = c[3];
                stats_out[CL2_NSTAT * k] = st[0]; stats_out[CL2_NSTAT * k + 1] = st[1]; stats_out[CL2_NSTAT * k + 2] = st[2];
            }
            flag[k] = r == 1 ? 1 : 0;
            if (scanned) atomicAdd(scanned_total, (unsigned long long)scanned);
        }
    }
}

// The candidates kernel A declined, wavelet form: one warp each, the lanes working on different range counts; the cost
// of a candidate no longer depends on how many rows its regions hold.
#define WM_WARPS 4
__global__ void __launch_bounds__(WM_WARPS * 32) k_est_bg_wm(const IdxView ix, const longlong4 *__restrict__ loops, int win, int mode,
                                                             int minPts, int hic, double cdc, double pseudo,
                                                             long long *__restrict__ core_out, double *__restrict__ stats_out,
                                                             int32_t *__restrict__ flag,
                                                             unsigned long long *__restrict__ scanned_total,
                                                             const int32_t *__restrict__ biglist,
                                                             const int32_t *__restrict__ nbig) {
    __shared__ WarpAcc acc[WM_WARPS];
    __shared__ WarpVals vals[WM_WARPS];
    __shared__ WmTab tab[WM_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n = *nbig;
    const bool cis = mode == CL2_MODE_CIS;
    for (int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < n; b += nwarps) {
        const int64_t k = biglist[b];
        const longlong4 lp = loops[k];
        const long long xs = lp.x, xe = lp.y, ys = lp.z, ye = lp.w;
        const double la = (double)(xe - xs), lb = (double)(ye - ys);
        long long scanned = 0;
        int64_t c[4];
        double st[3] = {0.0, 0.0, 0.0};
        int r = 1;
        if (la / lb > cdc || lb / la > cdc) r = 0;                                         // :282-286 / trans :135-139
        if (r) {
            core_wm(ix, xs, xe, ys, ye, mode, lane, c, &scanned);
            const double ra = (double)c[0], rb = (double)c[1], rab = (double)c[2];
            if (c[2] < minPts) r = 0;                                                      // :290 / trans :127
            if (cis && (c[0] == c[2] || c[1] == c[2])) r = 0;                              // :293
            if (ra / rb > cdc || rb / ra > cdc) r = 0;                                     // :296 / trans :133
            const double lower = (double)c[3];
            const double p2ll = rab / (lower > pseudo ? lower : pseudo);                   // :306
            if (cis && hic && p2ll < 1.0) r = 0;                                           // :307
            if (r) {
                windows_wm(ix, xs, xe, ys, ye, win, lane, acc[w], tab[w], &scanned);
                r = background_finish(rab, win, mode, acc[w], vals[w], lane, cis, pseudo, &st[0], &st[1], &st[2]);
            }
        }
        if (lane == 0) {
            if (r == 1) {
                core_out[4 * k] = c[0]; core_out[4 * k + 1] = c[1]; core_out[4 * k + 2] = c[2]; core_out[4 * k + 3] = c[3];
                stats_out[CL2_NSTAT * k] = st[0]; stats_out[CL2_NSTAT * k + 1] = st[1]; stats_out[CL2_NSTAT * k + 2] = st[2];
            }
            flag[k] = r == 1 ? 1 : 0;
            if (scanned) atomicAdd(scanned_total, (unsigned long long)scanned);
        }
        __syncwarp();
    }
}

// kernel B (survivors of kernel A, compacted): hypergeometric / Poisson / binomial tails, anchor
// windows and tests, and the remaining cuts (:311 hypergeometric > 1e-2, :347 anchor peaks unless -hic).
// Kernel B runs in two launches so that the fp64 series of the five tests do not run one after the other on a single
// lane: B1 (one block per survivor, its warps sharing the scans: the wide anchor windows of a large loop hold millions
// of rows) counts the anchor windows, B2 gives every survivor eight threads, five of which evaluate one tail
// probability each (same routines, same arguments as loop_tail()).
// rows the anchor counts of one loop have to scan (the X ranges of both anchors and their wide windows)
__device__ __forceinline__ long long anchor_rows(const IdxView &ix, const longlong4 lp, int lane) {
    const int k = (lane >> 1) & 3;                           // lanes 0..7: interval k, lower / upper bound in the X list
    const long long l = k < 2 ? lp.x : lp.z, r = k < 2 ? lp.y : lp.w;
    const long long m4 = 2 * (l + r), len = r - l;
    long long s4 = m4 - 20 * len;
    if (s4 < 0) s4 = 0;
    const Iv Wd = iv4(s4, m4 + 20 * len);
    Iv I;
    if (k & 1) I = Wd; else { I.lo = l; I.hi = r; }
    const uint32_t pos = lane < 8 ? bound_task(ix, I.lo, I.hi, 0, lane & 1) : 0u;
    long long rows = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) rows += (long long)__shfl_sync(0xffffffffu, pos, 2 * q + 1) - (long long)__shfl_sync(0xffffffffu, pos, 2 * q);
    return rows;
}
__global__ void __launch_bounds__(CNT_WARPS * 32) k_est_anchor(const IdxView ix, const longlong4 *__restrict__ loops,
                                                                const long long *__restrict__ sel, int64_t K,
                                                                double *__restrict__ stats,
                                                                unsigned long long *__restrict__ scanned_total,
                                                                int32_t *__restrict__ biglist, int32_t *__restrict__ nbig) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * CNT_WARPS + w;
    if (i >= K) return;
    long long scanned = 0;
    const longlong4 lp = loops[sel[i]];
    if (anchor_rows(ix, lp, lane) > BIG_ROWS) {              // k_est_anchor_big's work
        if (lane == 0) biglist[atomicAdd(nbig, 1)] = (int32_t)i;
        return;
    }
    int64_t cx, wx, cy, wy;
    loop_anchors(ix, lp.x, lp.y, lp.z, lp.w, lane, coop_solo(), &cx, &wx, &cy, &wy, &scanned);
    if (lane == 0) {
        double *o = stats + CL2_NSTAT * i;
        o[9] = (double)cx; o[10] = (double)wx; o[11] = (double)cy; o[12] = (double)wy;
        if (scanned) atomicAdd(scanned_total, (unsigned long long)scanned);
    }
}
// the survivors with large anchor windows (the wide window of a 100 kb anchor of Hi-C-like data holds 10^5 rows), one
// block each, its warps sharing the scans
__global__ void __launch_bounds__(CNT_WARPS * 32) k_est_anchor_big(const IdxView ix, const longlong4 *__restrict__ loops,
                                                                    const long long *__restrict__ sel,
                                                                    double *__restrict__ stats,
                                                                    unsigned long long *__restrict__ scanned_total,
                                                                    const int32_t *__restrict__ biglist,
                                                                    const int32_t *__restrict__ nbig) {
    __shared__ int red[COOP_RED];
    const int lane = threadIdx.x & 31;
    const Coop C = {CNT_WARPS, (int)(threadIdx.x >> 5), red, false};
    const int n = *nbig;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const int64_t i = biglist[b];
        long long scanned = 0;
        const longlong4 lp = loops[sel[i]];
        int64_t cx, wx, cy, wy;
        loop_anchors(ix, lp.x, lp.y, lp.z, lp.w, lane, C, &cx, &wx, &cy, &wy, &scanned);
        if (threadIdx.x == 0) {
            double *o = stats + CL2_NSTAT * i;
            o[9] = (double)cx; o[10] = (double)wx; o[11] = (double)cy; o[12] = (double)wy;
            if (scanned) atomicAdd(scanned_total, (unsigned long long)scanned);
        }
    }
}

// the same survivors from rectangle counts: |P(I)| = cX(I) + cY(I) - R(I, I), eight range counts per survivor
__global__ void __launch_bounds__(CNT_WARPS * 32) k_est_anchor_wm(const IdxView ix, const longlong4 *__restrict__ loops,
                                                                   const long long *__restrict__ sel,
                                                                   double *__restrict__ stats,
                                                                   unsigned long long *__restrict__ scanned_total,
                                                                   const int32_t *__restrict__ biglist,
                                                                   const int32_t *__restrict__ nbig) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n = *nbig;
    for (int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < n; b += nwarps) {
        const int64_t i = biglist[b];
        const longlong4 lp = loops[sel[i]];
        // intervals 0..3: anchor x, its wide window, anchor y, its wide window; lane 4k + 2*list + end searches a bound
        const int k = (lane >> 2) & 3;
        const long long l = k < 2 ? lp.x : lp.z, r = k < 2 ? lp.y : lp.w;
        const long long m4 = 2 * (l + r), len = r - l;
        long long s4 = m4 - 20 * len;
        if (s4 < 0) s4 = 0;
        const Iv Wd = iv4(s4, m4 + 20 * len);
        Iv I;
        if (k & 1) I = Wd; else { I.lo = l; I.hi = r; }
        const uint32_t pos = lane < 16 ? bound_task(ix, I.lo, I.hi, (lane >> 1) & 1, lane & 1) : 0u;
        // lanes 0..7: interval lane >> 1, lower / upper threshold
        const int q = (lane >> 1) & 3;
        const IvPos P = gather_pos(pos, q);
        const long long ilo = __shfl_sync(0xffffffffu, I.lo, 4 * q), ihi = __shfl_sync(0xffffffffu, I.hi, 4 * q);
        int f = 0;
        if (lane < 8) f = ((lane & 1) ? 1 : -1) * wm_less(ix, P.xb, P.xe, (lane & 1) ? ihi + 1 : ilo);
        long long res[4], bytes = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int rr = warp_sum((lane < 8 && q == d) ? f : 0);
            const IvPos Pd = gather_pos(pos, d);
            res[d] = ((long long)Pd.xe - Pd.xb) + ((long long)Pd.ye - Pd.yb) - rr;
            if (d & 1) bytes += 16 * pos_max(Pd) + 2 * 32 * (long long)ix.lg;
        }
        if (lane == 0) {
            double *o = stats + CL2_NSTAT * i;
            o[9] = (double)res[0]; o[10] = (double)res[1]; o[11] = (double)res[2]; o[12] = (double)res[3];
            atomicAdd(scanned_total, (unsigned long long)bytes);
        }
    }
}

#define EST_GROUP 8      // threads per survivor in k_est_stats
__global__ void __launch_bounds__(256) k_est_stats(const longlong4 *__restrict__ loops, const long long *__restrict__ sel,
                                                    int64_t K, int64_t n, int mode, int hic, double peakPcut, double slack,
                                                    double rpb,
                                                    const long long *__restrict__ core, double *__restrict__ hyp,
                                                    double *__restrict__ stats, int32_t *__restrict__ flag2) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / EST_GROUP;
    const int s = (int)(t % EST_GROUP);
    const bool live = i < K;
    const bool cis = mode == CL2_MODE_CIS;
    double v = 0.0, es = 0.0;
    if (live && s < 5) {
        const double ra = (double)core[4 * i], rb = (double)core[4 * i + 1], rab = (double)core[4 * i + 2];
        double *o = stats + CL2_NSTAT * i;
        if (s == 0) {
            v = cl2_poisson_sf_ge(rab, o[0]);                                             // :329
        } else if (s == 1) {
            double nn = cis ? ra * rb - rab : ra * rb;                                    // :332 / callTransLoops.py:181
            v = cl2_binom_sf_ge(rab, nn, o[1]);
        } else if (s == 4) {
            v = cl2_hypergeom_sf_ge(rab, (double)n, ra, rb);                              // :310
        } else {
            // estAnchorSig (:226-247): c = max(local mean of the +-5x window, global rate * length, 1)
            const longlong4 lp = loops[sel[i]];
            double count = o[s == 2 ? 9 : 11], wide = o[s == 2 ? 10 : 12];
            double len = (double)(s == 2 ? lp.y - lp.x : lp.w - lp.z);
            double r = (wide - count) / 5.0 / 2.0;
            double c = fmax(fmax(r, rpb * len), 1.0);
            v = cl2_poisson_sf_ge(count, c);
            es = count / c;
        }
    }
    // the group's five results meet in its first lane for the cuts (:311, :347)
    const int base = (threadIdx.x & 31) & ~(EST_GROUP - 1);
    const double px = __shfl_sync(0xffffffffu, v, base + 2), py = __shfl_sync(0xffffffffu, v, base + 3);
    const double h = __shfl_sync(0xffffffffu, v, base + 4);
    if (!live) return;
    double *o = stats + CL2_NSTAT * i;
    if (s == 0) o[3] = v;
    else if (s == 1) o[4] = v;
    else if (s == 2) { o[5] = v; o[6] = es; }
    else if (s == 3) { o[7] = v; o[8] = es; }
    if (s == 0) {
        bool keep = true;
        // slack = 1 + cl2_set_pvalue_margin: the cuts are widened by that factor for callers that take the final
        // decision with another library's tail probabilities
        if (cis && fmax(1e-300, h) > 1e-2 * slack) keep = false;                          // :311
        if (cis && !hic && !(fmax(1e-300, px) < peakPcut * slack && fmax(1e-300, py) < peakPcut * slack)) keep = false;   // :347
        hyp[i] = h;
        flag2[i] = keep ? 1 : 0;
    }
}

// gather the survivors of kernel A (flag set; pos = exclusive scan of the flags) into dense arrays, order preserved
__global__ void __launch_bounds__(256) k_est_compact(const int32_t *__restrict__ pos, int64_t L,
                                                      const long long *__restrict__ core, const double *__restrict__ stats,
                                                      long long *__restrict__ sel, long long *__restrict__ core_c,
                                                      double *__restrict__ stats_c) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    int p = pos[k];
    bool kept = (k + 1 < L) ? (pos[k + 1] != p) : (sel[-1] != (long long)p);   // sel[-1] holds the total
    if (!kept) return;
    sel[p] = k;
    for (int j = 0; j < 4; j++) core_c[4 * (int64_t)p + j] = core[4 * k + j];
    for (int j = 0; j < 3; j++) stats_c[CL2_NSTAT * (int64_t)p + j] = stats[CL2_NSTAT * k + j];
}

extern "C" int cl2_loop_counts(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t win, int32_t mode,
                               int64_t *core, int64_t *na, int64_t *nb, int64_t *nab, int64_t *anchors) {
    if (!ctx || !idx || L < 0 || (L > 0 && !loops)) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_counts: bad argument");
    if (win < 1 || 2 * win > CNT_MAXW) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_counts: win must be in 1..8");
    if (mode != CL2_MODE_CIS && mode != CL2_MODE_TRANS) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_counts: bad mode");
    if (L == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = 2 * win;
    Scratch sc(ctx);
    longlong4 *d_loops;
    long long *d_core = nullptr, *d_na = nullptr, *d_nb = nullptr, *d_nab = nullptr, *d_anc = nullptr;
    CL2_TRY(sc.get(&d_loops, (size_t)L));
    if (core) CL2_TRY(sc.get(&d_core, (size_t)L * 4));
    if (na) CL2_TRY(sc.get(&d_na, (size_t)L * W));
    if (nb) CL2_TRY(sc.get(&d_nb, (size_t)L * W));
    if (nab) CL2_TRY(sc.get(&d_nab, (size_t)L * W * W));
    if (anchors) CL2_TRY(sc.get(&d_anc, (size_t)L * 4));
    CL2_CUDA(ctx, cudaMemcpyAsync(d_loops, loops, (size_t)L * 32, cudaMemcpyHostToDevice, ctx->stream));
    {
        FamTimer t(ctx, FAM_COUNT, 0.0);
        k_loop_counts<<<cl2_grid(L, CNT_WARPS), CNT_WARPS * 32, 0, ctx->stream>>>(cl2_view(idx), d_loops, L, win,
                                                                                  mode, d_core, d_na, d_nb, d_nab, d_anc);
    }
    if (core) CL2_CUDA(ctx, cudaMemcpyAsync(core, d_core, (size_t)L * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (na) CL2_CUDA(ctx, cudaMemcpyAsync(na, d_na, (size_t)L * W * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nb) CL2_CUDA(ctx, cudaMemcpyAsync(nb, d_nb, (size_t)L * W * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nab) CL2_CUDA(ctx, cudaMemcpyAsync(nab, d_nab, (size_t)L * W * W * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (anchors) CL2_CUDA(ctx, cudaMemcpyAsync(anchors, d_anc, (size_t)L * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->h2d_bytes += L * 32;
    ctx->d2h_bytes += L * 8 * ((core ? 4 : 0) + (na ? W : 0) + (nb ? W : 0) + (nab ? W * W : 0) + (anchors ? 4 : 0));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CL2_CUDA(ctx, cudaGetLastError());
    return CL2_OK;
}

extern "C" int cl2_loop_core(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t mode, int64_t *core,
                             double *hyp) {
    if (!ctx || !idx || L < 0 || (L > 0 && (!loops || !core || !hyp))) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_core: bad argument");
    if (mode != CL2_MODE_CIS && mode != CL2_MODE_TRANS) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_core: bad mode");
    if (L == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    longlong4 *d_loops;
    long long *d_core;
    double *d_hyp;
    CL2_TRY(sc.get(&d_loops, (size_t)L));
    CL2_TRY(sc.get(&d_core, (size_t)L * 4));
    CL2_TRY(sc.get(&d_hyp, (size_t)L));
    CL2_CUDA(ctx, cudaMemcpyAsync(d_loops, loops, (size_t)L * 32, cudaMemcpyHostToDevice, ctx->stream));
    {
        FamTimer t(ctx, FAM_COUNT, 0.0);
        k_loop_core<<<cl2_grid(L, CNT_WARPS), CNT_WARPS * 32, 0, ctx->stream>>>(cl2_view(idx), d_loops, L, mode,
                                                                                d_core, d_hyp);
    }
    CL2_CUDA(ctx, cudaMemcpyAsync(core, d_core, (size_t)L * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaMemcpyAsync(hyp, d_hyp, (size_t)L * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->h2d_bytes += L * 32;
    ctx->d2h_bytes += L * 40;
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CL2_CUDA(ctx, cudaGetLastError());
    return CL2_OK;
}

extern "C" int cl2_loop_tests(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, const int64_t *core, int64_t L,
                              int32_t win, int32_t mode, double rpb, double *stats) {
    if (!ctx || !idx || L < 0 || (L > 0 && (!loops || !core || !stats))) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_tests: bad argument");
    if (win < 1 || 2 * win > CNT_MAXW) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_tests: win must be in 1..8");
    if (mode != CL2_MODE_CIS && mode != CL2_MODE_TRANS) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_loop_tests: bad mode");
    if (L == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    longlong4 *d_loops;
    long long *d_core;
    double *d_stats;
    CL2_TRY(sc.get(&d_loops, (size_t)L));
    CL2_TRY(sc.get(&d_core, (size_t)L * 4));
    CL2_TRY(sc.get(&d_stats, (size_t)L * CL2_NSTAT));
    CL2_CUDA(ctx, cudaMemcpyAsync(d_loops, loops, (size_t)L * 32, cudaMemcpyHostToDevice, ctx->stream));
    CL2_CUDA(ctx, cudaMemcpyAsync(d_core, core, (size_t)L * 32, cudaMemcpyHostToDevice, ctx->stream));
    {
        FamTimer t(ctx, FAM_COUNT, 0.0);
        k_loop_tests<<<cl2_grid(L, CNT_WARPS), CNT_WARPS * 32, 0, ctx->stream>>>(cl2_view(idx), d_loops, d_core, L,
                                                                                 win, mode, rpb, d_stats);
    }
    CL2_CUDA(ctx, cudaMemcpyAsync(stats, d_stats, (size_t)L * CL2_NSTAT * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->h2d_bytes += L * 64;
    ctx->d2h_bytes += L * CL2_NSTAT * 8;
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CL2_CUDA(ctx, cudaGetLastError());
    return CL2_OK;
}

__global__ void __launch_bounds__(256) k_est_gather(const long long *__restrict__ sel, int64_t K, const longlong4 *__restrict__ loops,
                                                     const int32_t *__restrict__ pid, const uint64_t *__restrict__ key,
                                                     longlong4 *__restrict__ oc, int32_t *__restrict__ op,
                                                     uint64_t *__restrict__ ok) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const long long k = sel[i];
    if (oc) oc[i] = loops[k];
    if (op) op[i] = pid[k];
    if (ok) ok[i] = key[k];
}

extern "C" int cl2_est_loops(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t win, int32_t mode,
                             int32_t minPts, int32_t hic, double countDiffCut, double pseudo, double peakPcut,
                             double rpb, int64_t *n_out, int64_t *sel, int64_t *core, double *hyp, double *stats) {
    if (!ctx || !idx || !n_out || L < 0 || (L > 0 && (!loops || !sel || !core || !hyp || !stats)))
        return cl2_fail(ctx, CL2_ERR_ARG, "cl2_est_loops: bad argument");
    *n_out = 0;
    if (L == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    longlong4 *d_loops;
    CL2_TRY(sc.get(&d_loops, (size_t)L));
    CL2_CUDA(ctx, cudaMemcpyAsync(d_loops, loops, (size_t)L * 32, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += L * 32;
    cl2_est_out out;
    CL2_TRY(cl2_est_loops_device(ctx, idx, d_loops, L, win, mode, minPts, hic, countDiffCut, pseudo, peakPcut, rpb, nullptr,
                                 nullptr, false, out));
    const size_t m = out.sel.size();
    if (m) {
        memcpy(sel, out.sel.data(), m * 8);
        memcpy(core, out.core.data(), m * 32);
        memcpy(hyp, out.hyp.data(), m * 8);
        memcpy(stats, out.stats.data(), m * CL2_NSTAT * 8);
    }
    *n_out = (int64_t)m;
    return CL2_OK;
}

int cl2_est_loops_device(cl2_ctx *ctx, cl2_index *idx, const longlong4 *d_loops, int64_t L, int32_t win, int32_t mode,
                         int32_t minPts, int32_t hic, double countDiffCut, double pseudo, double peakPcut, double rpb,
                         const int32_t *aux_pid, const uint64_t *aux_key, bool want_coords, cl2_est_out &out) {
    out.sel.clear(); out.core.clear(); out.coords.clear(); out.hyp.clear(); out.stats.clear(); out.pid.clear(); out.key.clear();
    if (!ctx || !idx || L < 0 || (L > 0 && !d_loops)) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_est_loops: bad argument");
    if (win < 1 || 2 * win > CNT_MAXW) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_est_loops: win must be in 1..8");
    if (mode != CL2_MODE_CIS && mode != CL2_MODE_TRANS) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_est_loops: bad mode");
    if (L == 0) return CL2_OK;
    if (L >= (int64_t)INT32_MAX) return cl2_fail(ctx, CL2_ERR_RANGE, "cl2_est_loops: too many loops in one call");
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    long long *d_core;
    double *d_stats;
    int32_t *d_flag;
    int64_t *d_total;
    CL2_TRY(sc.get(&d_core, (size_t)L * 4));
    CL2_TRY(sc.get(&d_stats, (size_t)L * CL2_NSTAT));
    CL2_TRY(sc.get(&d_flag, (size_t)L));
    int32_t *d_big;                                          // candidates for the cooperative kernel + their number
    CL2_TRY(sc.get(&d_big, (size_t)L + 1));
    CL2_CUDA(ctx, cudaMemsetAsync(d_big + L, 0, 4, ctx->stream));
    CL2_TRY(sc.get(&d_total, 2));
    CL2_CUDA(ctx, cudaMemsetAsync(d_total, 0, 16, ctx->stream));
    {
        FamTimer t(ctx, FAM_COUNT, 0.0);
        static const int occ = getenv("CL2_EST_OCC") ? atoi(getenv("CL2_EST_OCC")) : 4;
        auto kern = occ == 3 ? k_est_bg<3> : (occ == 2 ? k_est_bg<2> : k_est_bg<4>);
        kern<<<cl2_grid(L, CNT_WARPS), CNT_WARPS * 32, 0, ctx->stream>>>(cl2_view(idx), d_loops, L, win, mode, minPts, hic,
                                                                         countDiffCut, pseudo, d_core, d_stats, d_flag,
                                                                         (unsigned long long *)(d_total + 1), d_big, d_big + L);
    }
    // The candidates kernel A declined (regions of more than BIG_ROWS rows).  A few of them: a block each shares the
    // scans (k_est_bg_big).  Many of them (Hi-C like data, where nearly every candidate is one): range counts from the
    // wavelet matrix, built once per index, cost the same whatever the regions hold.  CL2_WM=0 / 1 forces either.
    int32_t nbig = 0;
    CL2_TRY(cl2_read_i32(ctx, d_big + L, &nbig));
    bool use_wm = false;
    if (nbig > 0) {
        const int wmcfg = getenv("CL2_WM") ? atoi(getenv("CL2_WM")) : -1;      // read per call: the tests switch it
        use_wm = 2 * win <= 10 && idx->n < (int64_t)INT32_MAX && idx->ymax < (1ll << 31) &&
                 (wmcfg == 1 || (wmcfg != 0 && (int64_t)nbig * BIG_ROWS >= idx->n));
        if (use_wm) CL2_TRY(cl2_index_wavelet(ctx, idx));
        FamTimer t(ctx, FAM_COUNT, 0.0);
        if (use_wm) {
            const int64_t cap = (int64_t)ctx->sm_count * 16, want = cl2_grid(nbig, WM_WARPS);
            k_est_bg_wm<<<(unsigned)(want < cap ? want : cap), WM_WARPS * 32, 0, ctx->stream>>>(
                cl2_view(idx), d_loops, win, mode, minPts, hic, countDiffCut, pseudo, d_core, d_stats, d_flag,
                (unsigned long long *)(d_total + 1), d_big, d_big + L);
        } else {
            const int64_t cap = (int64_t)ctx->sm_count * 8;
            k_est_bg_big<256, 3><<<(unsigned)(nbig < cap ? nbig : cap), 256, 0, ctx->stream>>>(
                cl2_view(idx), d_loops, win, mode, minPts, hic, countDiffCut, pseudo, d_core, d_stats, d_flag,
                (unsigned long long *)(d_total + 1), d_big, d_big + L);
        }
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, d_flag, L, d_total));
    int64_t nk = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &nk));
    std::vector<int32_t> f2;
    if (nk > 0) {
        // d_selbuf[0] carries the total so the compaction kernel can tell whether the last row survived
        long long *d_selbuf, *d_sel, *d_core_c;
        double *d_hyp_c, *d_stats_c;
        int32_t *d_flag2;
        CL2_TRY(sc.get(&d_selbuf, (size_t)nk + 1));
        d_sel = d_selbuf + 1;
        CL2_TRY(sc.get(&d_core_c, (size_t)nk * 4));
        CL2_TRY(sc.get(&d_hyp_c, (size_t)nk));
        CL2_TRY(sc.get(&d_stats_c, (size_t)nk * CL2_NSTAT));
        CL2_TRY(sc.get(&d_flag2, (size_t)nk));
        CL2_CUDA(ctx, cudaMemcpyAsync(d_selbuf, d_total, 8, cudaMemcpyDeviceToDevice, ctx->stream));
        {
            FamTimer t(ctx, FAM_MISC, 60.0 * (double)nk + 4.0 * (double)L);
            k_est_compact<<<cl2_grid(L, 256), 256, 0, ctx->stream>>>(d_flag, L, d_core, d_stats, d_sel, d_core_c, d_stats_c);
        }
        {
            FamTimer t(ctx, FAM_COUNT, 0.0, 2);
            CL2_CUDA(ctx, cudaMemsetAsync(d_big + L, 0, 4, ctx->stream));       // the list is reused for the anchor stage
            k_est_anchor<<<cl2_grid(nk, CNT_WARPS), CNT_WARPS * 32, 0, ctx->stream>>>(cl2_view(idx), d_loops, d_sel, nk, d_stats_c,
                                                                                      (unsigned long long *)(d_total + 1), d_big,
                                                                                      d_big + L);
            const int64_t acap = (int64_t)ctx->sm_count * 8;
            if (idx->wm)
                k_est_anchor_wm<<<(unsigned)(cl2_grid(nk, CNT_WARPS) < acap ? cl2_grid(nk, CNT_WARPS) : acap), CNT_WARPS * 32, 0, ctx->stream>>>(
                    cl2_view(idx), d_loops, d_sel, d_stats_c, (unsigned long long *)(d_total + 1), d_big, d_big + L);
            else
                k_est_anchor_big<<<(unsigned)(nk < acap ? nk : acap), CNT_WARPS * 32, 0, ctx->stream>>>(
                    cl2_view(idx), d_loops, d_sel, d_stats_c, (unsigned long long *)(d_total + 1), d_big, d_big + L);
            ctx->launches++;
            k_est_stats<<<cl2_grid(nk * EST_GROUP, 256), 256, 0, ctx->stream>>>(d_loops, d_sel, nk, idx->n, mode, hic, peakPcut,
                                                                                1.0 + ctx->pvalue_margin, rpb, d_core_c,
                                                                                d_hyp_c, d_stats_c, d_flag2);
        }
        f2.resize((size_t)nk);
        out.sel.resize((size_t)nk);
        out.core.resize((size_t)nk * 4);
        out.hyp.resize((size_t)nk);
        out.stats.resize((size_t)nk * CL2_NSTAT);
        CL2_CUDA(ctx, cudaMemcpyAsync(out.sel.data(), d_sel, (size_t)nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CL2_CUDA(ctx, cudaMemcpyAsync(out.core.data(), d_core_c, (size_t)nk * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CL2_CUDA(ctx, cudaMemcpyAsync(out.hyp.data(), d_hyp_c, (size_t)nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CL2_CUDA(ctx, cudaMemcpyAsync(out.stats.data(), d_stats_c, (size_t)nk * CL2_NSTAT * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CL2_CUDA(ctx, cudaMemcpyAsync(f2.data(), d_flag2, (size_t)nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += nk * (8 + 32 + 8 + CL2_NSTAT * 8 + 4);
        if (want_coords || aux_pid || aux_key) {
            // the survivors' own rows (candidates that live on the device only)
            longlong4 *g_c = nullptr;
            int32_t *g_p = nullptr;
            uint64_t *g_k = nullptr;
            if (want_coords) CL2_TRY(sc.get(&g_c, (size_t)nk));
            if (aux_pid) CL2_TRY(sc.get(&g_p, (size_t)nk));
            if (aux_key) CL2_TRY(sc.get(&g_k, (size_t)nk));
            k_est_gather<<<cl2_grid(nk, 256), 256, 0, ctx->stream>>>(d_sel, nk, d_loops, aux_pid, aux_key, g_c, g_p, g_k);
            ctx->launches++;
            if (want_coords) {
                out.coords.resize((size_t)nk * 4);
                CL2_CUDA(ctx, cudaMemcpyAsync(out.coords.data(), g_c, (size_t)nk * 32, cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (aux_pid) {
                out.pid.resize((size_t)nk);
                CL2_CUDA(ctx, cudaMemcpyAsync(out.pid.data(), g_p, (size_t)nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (aux_key) {
                out.key.resize((size_t)nk);
                CL2_CUDA(ctx, cudaMemcpyAsync(out.key.data(), g_k, (size_t)nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
            }
            ctx->d2h_bytes += nk * ((want_coords ? 32 : 0) + (aux_pid ? 4 : 0) + (aux_key ? 8 : 0));
        }
        CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->timing) {
        // algorithmic bytes of the counting launches, summed by the kernels per candidate (16 B per
        // member of every queried region + the binary searches) + the candidate rows themselves
        int64_t abytes = 0;
        CL2_TRY(cl2_read_i64(ctx, d_total + 1, &abytes));
        ctx->fam_bytes[FAM_COUNT] += (double)abytes + 32.0 * (double)L;
    }
    CL2_CUDA(ctx, cudaGetLastError());
    // rows that failed kernel B's cuts are dropped here (few rows: kernel A already removed ~95 % of the candidates)
    int64_t m = 0;
    for (int64_t i = 0; i < nk; i++) {
        if (!f2[i]) continue;
        if (m != i) {
            out.sel[m] = out.sel[i];
            memcpy(&out.core[4 * m], &out.core[4 * i], 32);
            out.hyp[m] = out.hyp[i];
            memcpy(&out.stats[CL2_NSTAT * m], &out.stats[CL2_NSTAT * i], CL2_NSTAT * 8);
            if (!out.coords.empty()) memcpy(&out.coords[4 * m], &out.coords[4 * i], 32);
            if (!out.pid.empty()) out.pid[m] = out.pid[i];
            if (!out.key.empty()) out.key[m] = out.key[i];
        }
        m++;
    }
    out.sel.resize((size_t)m);
    out.core.resize((size_t)m * 4);
    out.hyp.resize((size_t)m);
    out.stats.resize((size_t)m * CL2_NSTAT);
    if (!out.coords.empty()) out.coords.resize((size_t)m * 4);
    if (!out.pid.empty()) out.pid.resize((size_t)m);
    if (!out.key.empty()) out.key.resize((size_t)m);
    return CL2_OK;
}
