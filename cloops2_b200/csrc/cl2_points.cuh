// cl2_points.cuh -- device-resident form of one .ixy file and the cached eps-grid built on it.
#pragma once
#include "cl2_common.cuh"

// Binning of the points into eps-sized cells, sorted by packed cell key; depends on (eps, grid kind) only, so it is
// kept across the minPts values of a sweep (the reference loops eps-major, callCisLoops.py:541-551).
struct cl2_grid_cache {
    bool valid = false;
    int kind = 0;             // 0 = blockDBSCAN grid (min-anchored, axis aligned), 1 = cDBSCAN2 grid (rotated)
    int64_t eps = 0;
    int64_t n = 0;            // points in the grid: all of them, or the rows of the subset it was built from
    int32_t min_minpts = -1;  // -1: built from every point; else valid for minPts >= this (built from a cl2_subset)
    int sub_level = -1;       // which pts->subs[] entry it was built from (-1: every point)
    int32_t ncells = 0;
    int by = 0;               // bits of the low (gy) field of the packed key
    uint32_t *row = nullptr;  // [n]   original row id of sorted position i
    int2 *sxy = nullptr;      // [n]   (x,y) of sorted position i   (cDBSCAN2 grid: (u,v) rotated, int64 pairs in suv)
    int32_t *cid = nullptr;   // [n]   cell id of sorted position i
    int32_t *cstart = nullptr;  // [ncells+1]
    uint64_t *ckey = nullptr;   // [ncells]
    int32_t *nbr = nullptr;     // [ncells*8] neighbour cell id or -1, order of blockDBSCAN.py:54-55
    int32_t *s3 = nullptr;      // [ncells]  population of the 3x3 block
    uint32_t *first = nullptr;  // [ncells]  smallest file row of the cell = its dict position in the reference
};

// Rows that survive the noise-cell removal of earlier blockDBSCAN runs with larger eps (cl2_block.cu).  A point in a
// cell that run (E, m) removed is also removed by any run (e, m') with m' >= m and E >= 3e - gcd(e, E), and removing it
// beforehand changes nothing for the points that stay, so such runs bin and sort only these rows.
#define CL2_SUB_MAXLINKS 8
struct cl2_subset {
    int64_t n = 0;
    int2 *byy = nullptr;        // [n] (y, x) of the kept rows, sorted by y
    uint32_t *rowy = nullptr;   // [n] their file rows
    int nlinks = 0;             // the runs that produced it
    int64_t link_eps[CL2_SUB_MAXLINKS];
    int32_t link_minpts[CL2_SUB_MAXLINKS];
};

struct cl2_points {
    int64_t n = 0;
    int2 *xy = nullptr;      // [n] int32 coordinates, file order after the cut/mcut filter
    int64_t ext[4] = {0, 0, 0, 0};  // minX, maxX, minY, maxY
    // y-order of the points, built once (XY index, and the starting order of every blockDBSCAN grid sort: a stable
    // sort of the y-sorted rows by column index alone yields (column, row-of-grid) order in 3 passes instead of 5-6)
    int2 *byy = nullptr;        // [n] (y, x) sorted by y
    uint32_t *rowy = nullptr;   // [n] file row of y-sorted position
    // x-order of the points derived from a blockDBSCAN grid over every row (cl2_block.cu: the grid's columns are x
    // intervals in ascending order, so sorting each column locally by x yields the global order -- one pass instead of
    // a radix sort).  Built only when the caller announced an index build (want_xorder); cl2_index_build takes it over
    // if xorder_bad (device flag: some column was too long for the local sort) stayed 0.
    int2 *byx = nullptr;        // [n] (x, y) sorted by x
    int32_t *xorder_bad = nullptr;
    int want_xorder = 0;
    cl2_grid_cache grid;
    std::vector<cl2_subset> subs;   // nested subsets, subs[k+1] derived from subs[k]; only kept when sweep_hint is set
    int sweep_hint = 0;         // cl2_points_sweep_hint: 0 off, 1 remember rows and use them, 2 use only
    // result of the last clustering run
    int32_t ncl = 0;
    std::vector<int64_t> bbox;  // [ncl*4]
    std::vector<int64_t> size;  // [ncl]
    std::vector<uint64_t> ckeys;  // [ncl] rank key of every cluster in the reference's numbering (blockDBSCAN runs)
};

struct cl2_index {
    int64_t n = 0;
    int2 *byx = nullptr;  // [n] (x, y) sorted by x
    int2 *byy = nullptr;  // [n] (y, x) sorted by y
    // bucket tables over the sorted coordinate (cl2_count.cu): t[b] = first position with coordinate >= b << shift
    uint32_t *tx = nullptr, *ty = nullptr;
    int sx = 0, sy = 0, nbx = 0, nby = 0;
    int small = 0;        // every coordinate < 2^29
    // Wavelet matrix over the y values in x order (built on demand, cl2_count.cu): level l holds, for every 32 rows in
    // the order after l stable partitions, the bits at position wm_L-1-l of their y and the number of ones before them;
    // wm_z[l] = zeros of level l.  Answers "rows at x positions [p, q) with y < v" in wm_L steps.
    uint2 *wm = nullptr;
    int wm_L = 0;
    int64_t wm_nw = 0;
    uint32_t wm_z[32] = {0};
    int64_t ymax = 0;
};

// Device-resident candidate list of one cl2_call_loops: every clustering run appends the boxes that pass the per-run
// candidate filters (and, cis, filterLoopsByDis), each with the index of its run and the key that ranks it in the
// reference's cluster numbering.  The boxes never visit the host: estLoopSig reads them where they are and only the
// few survivors are copied back (cl2_driver.cu).
struct cl2_cand_sink {
    longlong4 *cand = nullptr;     // [cap] x_start, x_end, y_start, y_end
    uint64_t *key = nullptr;       // [cap]
    int32_t *pid = nullptr;        // [cap]
    int64_t cap = 0;
    int64_t bound = 0;             // host-side upper bound of the rows used (sum of the cluster counts emitted so far)
    int32_t *cursor = nullptr;     // device: rows used
    long long *pass_stat = nullptr;  // device [2 * npass]: candidates of the run (before the distance filter), PETs in them
    int32_t mode = 0;              // CL2_MODE_CIS | CL2_MODE_TRANS: which per-run filters apply
    int64_t cut = 0;               // filterLoopsByDis keeps distance > cut (cis)
    int32_t pass = 0;              // index of the run that is emitting
};
// boxes: ncomp records of 5 ints (xmin, xmax, ymin, ymax, size) on the device; keys may be null (key = record index)
int cl2_sink_emit(cl2_ctx *ctx, cl2_cand_sink *sink, const int *boxes, const uint64_t *keys, int32_t ncomp);   // cl2_driver.cu
void cl2_sink_free(cl2_ctx *ctx, cl2_cand_sink *sink);

void cl2_grid_free(cl2_ctx *ctx, cl2_grid_cache &g);
void cl2_subsets_free(cl2_ctx *ctx, cl2_points *pts, size_t keep_levels = 0);
// keygen -> radix sort -> gather -> cell table -> neighbours + 3x3 sums (cl2_block.cu); cached per (eps, kind).
// minPts >= 0 allows a grid over the smallest pts->subs[] entry valid for (eps, minPts); -1 asks for every point.
int cl2_grid_build(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int kind, int32_t minPts = -1);
int cl2_points_yorder(cl2_ctx *ctx, cl2_points *pts);   // cl2_count.cu
void cl2_points_drop_xorder(cl2_ctx *ctx, cl2_points *pts);   // cl2_api.cu
int cl2_block_dbscan_impl(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                          int32_t *n_clusters, bool ordered, cl2_cand_sink *sink = nullptr);    // cl2_block.cu
int cl2_cdbscan2_impl(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                      int32_t *n_clusters, cl2_cand_sink *sink);                                // cl2_cdbscan.cu
// estLoopSig on candidates that already live on the device (cl2_count.cu).  aux_pid / aux_key (device, may be null)
// are gathered for the survivors along with their coordinates.
struct cl2_est_out {
    std::vector<int64_t> sel, core, coords;
    std::vector<double> hyp, stats;
    std::vector<int32_t> pid;
    std::vector<uint64_t> key;
};
int cl2_est_loops_device(cl2_ctx *ctx, cl2_index *idx, const longlong4 *d_loops, int64_t L, int32_t win, int32_t mode,
                         int32_t minPts, int32_t hic, double countDiffCut, double pseudo, double peakPcut, double rpb,
                         const int32_t *aux_pid, const uint64_t *aux_key, bool want_coords, cl2_est_out &out);
int cl2_read_i64(cl2_ctx *ctx, const int64_t *d, int64_t *out);
int cl2_read_i32(cl2_ctx *ctx, const int32_t *d, int32_t *out);

static inline int bits_for(uint64_t v) {  // number of bits needed to represent values 0..v
    int b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b < 1 ? 1 : b;
}
