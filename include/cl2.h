/*
 * cl2.h -- C ABI of libcl2.so: the B200 (sm_100a) implementation of the cLoops2 `callLoops` hot path.
 *
 * The reference (cLoops2 0.0.5, pure Python) has no FFI; these entry points are what a ctypes binding inside
 * cLoops2 would bind for this path (INTEGRATION.md shows the stub).  Each entry cites the reference interface it
 * replaces (path:line under the cLoops2 checkout).
 *
 * Conventions: caller-owned host buffers, plain pointers and sizes; every call returns CL2_OK (0) or a negative
 * error code, with a message retrievable through cl2_last_error(ctx); no exceptions cross the ABI.  One cl2_ctx
 * per GPU; calls on one context must be serialised by the caller, independent contexts may be used from
 * different host threads.  Labels are int32, -1 = not clustered, cluster ids numbered exactly as the reference's
 * `.labels` dict values.
 */
#ifndef CL2_H
#define CL2_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CL2_OK 0
#define CL2_ERR_CUDA (-1)      /* a CUDA runtime call failed (message has the cudaError string) */
#define CL2_ERR_ARG (-2)       /* bad argument (null pointer, eps <= 0, n < 0, coordinates out of range) */
#define CL2_ERR_NOMEM (-3)     /* device or host allocation failed */
#define CL2_ERR_RANGE (-4)     /* coordinate span does not fit the on-device int32 layout */

#define CL2_MODE_CIS 0
#define CL2_MODE_TRANS 1

typedef struct cl2_ctx cl2_ctx;        /* one GPU: stream, memory pool, timers */
typedef struct cl2_points cl2_points;  /* one .ixy file resident in HBM (kept across the eps x minPts sweep) */
typedef struct cl2_index cl2_index;    /* device form of cLoops2/ds.py:141-198 (XY): X-sorted and Y-sorted arrays */
typedef struct cl2_bedpe cl2_bedpe;    /* BEDPE tokeniser state of `pre`: chromosome name <-> id table */

/* version string of the library, e.g. "cl2-b200 0.1 sm_100a" */
const char *cl2_version(void);

/* ---- context ------------------------------------------------------------------------------------------ */
int cl2_ctx_create(int device, cl2_ctx **out);
void cl2_ctx_destroy(cl2_ctx *ctx);
const char *cl2_last_error(cl2_ctx *ctx);
/* blocks until all work queued on the context's stream has finished */
int cl2_ctx_sync(cl2_ctx *ctx);

/* Page-locked host staging (replaces the numpy array joblib.load returns in cLoops2/io.py:274-291). */
int cl2_host_alloc(cl2_ctx *ctx, int64_t bytes, void **out);
void cl2_host_free(cl2_ctx *ctx, void *p);

/* ---- points ------------------------------------------------------------------------------------------- */
/*
 * Upload one .ixy matrix: xy = int64[n,2] row-major (x,y interleaved), the layout joblib stores
 * (cLoops2/io.py:55-67, 274-291).  Rows failing (y-x >= cut when cut>0) or (y-x <= mcut when mcut>0) are dropped
 * on the device, order preserved, exactly as parseIxy does (io.py:282-289); row ids are positions after the filter
 * (callCisLoops.py:67-72).  cut=0, mcut=-1 keeps everything.
 */
int cl2_points_upload(cl2_ctx *ctx, const int64_t *xy, int64_t n, int64_t cut, int64_t mcut, cl2_points **out);
/* number of rows kept after the cut/mcut filter */
int64_t cl2_points_count(const cl2_points *pts);
/* min X, max X, min Y, max Y of the kept rows (callCisLoops.py:231 uses max Y - min X) */
int cl2_points_extent(cl2_ctx *ctx, const cl2_points *pts, int64_t ext[4]);
/* kept rows back to the host as int64[n,2] (for -filter, callCisLoops.py:450-468) */
int cl2_points_download(cl2_ctx *ctx, const cl2_points *pts, int64_t *xy_out);
void cl2_points_free(cl2_ctx *ctx, cl2_points *pts);
/* Release the derived data cached on `pts` (the eps grid kept between the minPts values of one eps, and the
 * y-sorted order shared by the grid sorts and the XY index); the points themselves stay resident. */
void cl2_points_drop_cache(cl2_ctx *ctx, cl2_points *pts);
/* Announce (enable != 0) that several cl2_block_dbscan runs with decreasing eps follow on `pts`, as the eps loop of
 * callCisLoops.py:541-551 / callTransLoops.py:241-251 does.  Each run then remembers the rows outside the noise
 * cells it removed (blockDBSCAN.py:101-122); a later run (e, m') after a run (E, m) with m' >= m and
 * E >= 3e - gcd(e, E) would remove those rows too and is unaffected by their absence, so it bins and sorts only the
 * remembered rows.  enable = 2 keeps using what is remembered without remembering more (the runs at the smallest eps
 * of a sweep).  Results are identical with and without the hint; 0 disables it and frees the remembered rows. */
int cl2_points_sweep_hint(cl2_ctx *ctx, cl2_points *pts, int32_t enable);
/* Announce (enable != 0) that a cl2_index_build follows the next cl2_block_dbscan runs on `pts`, as estLoopSig
 * follows the clustering runs of a file (callCisLoops.py:561-572 after :541-551).  The first run whose grid holds
 * every row then also leaves the x-order of the points behind (the grid's columns are x intervals, so sorting each
 * column locally by x gives ds.py:147's `xs`), and cl2_index_build takes it over instead of sorting again.  The index
 * is the same either way; cl2_call_loops does this by itself. */
int cl2_points_index_hint(cl2_ctx *ctx, cl2_points *pts, int32_t enable);
/* Rows of `pts` without exact duplicates, ascending by (X, Y) -- the de-duplication of `cLoops2 pre` (getUniqueTxt,
 * cLoops2/io.py:40-52, which keeps a Python set of the text rows of one chromosome pair).  xy_out: int64[n,2] host
 * buffer with n = cl2_points_count(pts); *n_out rows are written. */
int cl2_points_unique(cl2_ctx *ctx, cl2_points *pts, int64_t *xy_out, int64_t *n_out);

/* ---- clustering --------------------------------------------------------------------------------------- */
/*
 * blockDBSCAN(mat, eps, minPts).labels -- cLoops2/blockDBSCAN.py:13-41 (call sites callCisLoops.py:87,
 * callTransLoops.py:53).  labels_out: int32[n] host buffer or NULL (skip the per-point label read-back).
 */
int cl2_block_dbscan(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                     int32_t *n_clusters);
/* cDBSCAN(mat, eps, minPts).labels -- cLoops2/cDBSCAN2.py:13-35 (callCisLoops.py:499-502, -hic). */
int cl2_cdbscan2(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                 int32_t *n_clusters);
/*
 * Per-cluster bounding box and size of the LAST clustering run on `pts` (replaces the pandas scan at
 * callCisLoops.py:92-105 / callTransLoops.py:58-76).  bbox_out: int64[n_clusters,4] = xmin,xmax,ymin,ymax;
 * size_out: int64[n_clusters].
 */
int cl2_cluster_bbox(cl2_ctx *ctx, cl2_points *pts, int64_t *bbox_out, int64_t *size_out);
/* blockDBSCAN's noise-cell removal alone (blockDBSCAN.py:69-122; also cLoops2/filter.py:223-278, `filterPETs -knn`):
 * keep_out[i] = 1 iff row i survives removeNoiseGrids.  cell_first_out (int32[n], may be NULL) receives for every row
 * the smallest row id of its grid cell = the cell's position in the reference's dict, which is the order
 * filter.py:273-275 writes the surviving rows in. */
int cl2_block_noise(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, uint8_t *keep_out,
                    int32_t *cell_first_out);

/* ---- counting ----------------------------------------------------------------------------------------- */
/* XY(xs, ys) -- cLoops2/ds.py:146-163, built by ixy2pet (io.py:294-300). */
int cl2_index_build(cl2_ctx *ctx, cl2_points *pts, cl2_index **out);
void cl2_index_free(cl2_ctx *ctx, cl2_index *idx);
/*
 * Raw counts behind estLoopSig (callCisLoops.py:250-354; trans: callTransLoops.py:106-193) for L candidate
 * loops, loops = int64[L,4] x_start,x_end,y_start,y_end.  Any output pointer may be NULL to skip that group.
 *   core    int64[L,4]   ra, rb, rab (queryLoop, ds.py:184-190) and the P2LL lower-left |A' & B'|
 *                        (cis window callCisLoops.py:302-305, trans window callTransLoops.py:142-144)
 *   na, nb  int64[L,2*win]          |Na_i|, |Nb_j| of the shifted windows (getPerRegions, callCisLoops.py:173-198)
 *   nab     int64[L,(2*win)^2]      |Na_i & Nb_j|, a-major (getLoopNearbyPETs, callCisLoops.py:207-223)
 *   anchors int64[L,4]   count_x, wide_x, count_y, wide_y (estAnchorSig, callCisLoops.py:226-247, ext = 5)
 */
int cl2_loop_counts(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t win, int32_t mode,
                    int64_t *core, int64_t *na, int64_t *nb, int64_t *nab, int64_t *anchors);

/*
 * Fused counting + statistics, phase 1 (every candidate): core[L,4] = ra, rb, rab, lower as in cl2_loop_counts and
 * hyp[L] = hypergeom.sf(rab - 1, N, ra, rb) with N = rows of the index (callCisLoops.py:310), NOT clamped to 1e-300.
 */
int cl2_loop_core(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t mode, int64_t *core,
                  double *hyp);
/*
 * Phase 2 (candidates that passed the first filters): permuted-window background and the remaining tests of
 * estLoopSig (callCisLoops.py:314-346; trans callTransLoops.py:146-181).  core = the [L,4] rows phase 1 returned for
 * these loops; rpb = PETs per bp, N / (max Y - min X) (callCisLoops.py:231).
 * stats[L,13] = mrabs (median; mean for trans), mbps, FDR, poisson.sf(rab-1, mrabs), binom.sf(rab-1, n, mbps) with
 * n = ra*rb-rab (trans: ra*rb), x_peak_poisson_p, x_peak_es, y_peak_poisson_p, y_peak_es, and the anchor counts
 * count_x, wide_x, count_y, wide_y of estAnchorSig.
 * Tail probabilities are NOT clamped to 1e-300 (the host applies max(1e-300, p) as the reference does).
 */
int cl2_loop_tests(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, const int64_t *core, int64_t L, int32_t win,
                   int32_t mode, double rpb, double *stats);

/*
 * estLoopSig in one call: counting, statistics AND the reference's filter cascade on the device (cis
 * callCisLoops.py:282-347 with `hic` selecting the -hic rules; trans callTransLoops.py:127-139, no early exits).
 * Only the candidates that survive are returned, in candidate order: *n_out rows, sel[i] = index of the i-th survivor
 * in `loops`, core / hyp / stats as in cl2_loop_core and cl2_loop_tests.  All output buffers must hold L rows.
 * countDiffCut, pseudo, peakPcut are estLoopSig's keyword arguments (20 / 1 / 1e-5 cis; 10 / 1 / - trans).
 */
int cl2_est_loops(cl2_ctx *ctx, cl2_index *idx, const int64_t *loops, int64_t L, int32_t win, int32_t mode,
                  int32_t minPts, int32_t hic, double countDiffCut, double pseudo, double peakPcut, double rpb,
                  int64_t *n_out, int64_t *sel, int64_t *core, double *hyp, double *stats);

/*
 * The device-side cuts on tail probabilities inside cl2_est_loops / cl2_call_loops (hypergeometric > 1e-2,
 * callCisLoops.py:311; anchor peaks >= peakPcut, :347) are widened by the factor 1 + margin, so that a caller who
 * re-evaluates the tails with another library (scipy, whose values differ from the fp64 series here in the 8th digit
 * at large populations) receives every row its own decision could keep and takes that decision itself.  0 (default)
 * = the cuts as the reference writes them.
 */
int cl2_set_pvalue_margin(cl2_ctx *ctx, double margin);

/* ---- the per-file work unit ----------------------------------------------------------------------------- */
/*
 * What the reference hands to one joblib worker, for ONE chromosome (pair) file resident in HBM, in one call:
 * runCisDBSCANLoops for every (eps[i], minPts[i]) of the sweep (callCisLoops.py:58-124 as driven by :529-551; trans:
 * runTransDBSCANLoops, callTransLoops.py:33-84, :224-236) with its per-run candidate filters, filterLoopsByDis
 * (distance > cut, callCisLoops.py:146-156, cis only), estLoopSig on the candidates (callCisLoops.py:250-354 /
 * callTransLoops.py:106-193; arguments as in cl2_est_loops, min_pts_test = the minPts handed to estLoopSig,
 * callCisLoops.py:557-568) and the de-duplication of combineLoops / filterDupLoops (geo.py:90-112,
 * callCisLoops.py:159-169; trans keeps duplicates that arise inside one run).  mode = CL2_MODE_CIS | CL2_MODE_TRANS;
 * hic != 0 clusters with cDBSCAN2 and applies the -hic rules (cis only).  The runs may be executed in any order
 * (results are reported in the order given); the derived data cached on `pts` is dropped at the end.
 * want_unique != 0 also counts the distinct candidate coordinates (the number logged at callCisLoops.py:564-566).
 * The survivors wait in *out until cl2_loops_free.
 */
typedef struct cl2_loops cl2_loops;
int cl2_call_loops(cl2_ctx *ctx, cl2_points *pts, int32_t mode, int32_t hic, int32_t npass, const int64_t *eps,
                   const int32_t *minPts, int64_t cut, int32_t min_pts_test, int32_t win, double countDiffCut,
                   double pseudo, double peakPcut, int32_t want_unique, cl2_loops **out);
/* surviving loops, in candidate order (run by run as given, ascending cluster id inside a run) */
int64_t cl2_loops_count(const cl2_loops *r);
/* pass_cand[npass], pass_pets[npass]: candidates of every run after its filters and the PETs they hold (the numbers
 * logged at callCisLoops.py:122); totals[8] = PETs of the file, candidates tested, distinct candidate coordinates
 * (-1 unless want_unique), survivors, min X, max X, min Y, max Y; seconds[6] = host wall time of the clustering
 * runs, the index build, the tests, then the calling thread's CPU time of the same three.  Any pointer may be NULL. */
int cl2_loops_info(const cl2_loops *r, int64_t *pass_cand, int64_t *pass_pets, int64_t *totals, double *seconds);
/* coords int64[K,4], pass_id int32[K] (index of the run a survivor came from), core / hyp / stats as cl2_est_loops */
int cl2_loops_read(const cl2_loops *r, int64_t *coords, int32_t *pass_id, int64_t *core, double *hyp, double *stats);
void cl2_loops_free(cl2_loops *r);

/* ---- multi-GPU: the path's only collective ------------------------------------------------------------------ */
/*
 * One process per GPU, chromosome (pair) files sharded over the ranks (the reference fans the same units out over
 * joblib workers, callCisLoops.py:134-137, 561-572); no data flows between units, so the only exchange is the sum of
 * the per-rank counters (PETs, candidates, tested, loops) that the reference's single process simply accumulates.
 * cl2_comm wraps an NCCL communicator (libnccl is loaded at run time): rank 0 calls cl2_comm_unique_id and hands the
 * CL2_COMM_ID_BYTES bytes to the other ranks over any host channel (a file, MPI, the launcher's store), then every rank
 * calls cl2_comm_init with the same id.  cl2_stats_allreduce sums vals[0..k) over the ranks in place (host buffer;
 * ncclAllReduce of int64 on the context's stream).  Per-file loop lists are gathered by the host program.
 */
#define CL2_COMM_ID_BYTES 128
typedef struct cl2_comm cl2_comm;
int cl2_comm_unique_id(uint8_t *id_out);
int cl2_comm_init(cl2_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world, cl2_comm **out);
int cl2_stats_allreduce(cl2_ctx *ctx, cl2_comm *comm, int64_t *vals, int32_t k);
void cl2_comm_destroy(cl2_comm *comm);

/* ---- host-side selection (native, no GPU) ----------------------------------------------------------------- */
/*
 * selSigLoops (callCisLoops.py:379-422) for the loops of ONE chromosome (pair), all already significant, in list
 * order: greedy single-link overlap groups in order, per group the first loop attaining the maximum ES, then the
 * "search again" pass.  coords = int64[L,4] x_start,x_end,y_start,y_end; es = double[L].
 * out_idx: int64[L] receives the indices of the selected loops in output order; *out_n their number.
 */
int cl2_sel_sig_loops(const int64_t *coords, const double *es, int64_t L, int64_t *out_idx, int64_t *out_n);

/*
 * Rows of loops2txt (cLoops2/io.py:517-561) for n loops of one chromosome pair, header excluded, every field as
 * Python's str() prints it (ints as ints, floats in repr() form -- shortest round-trip digits, fixed notation with at
 * least one decimal for 1e-4 <= |x| < 1e16, d.ddde-XX otherwise --, `True` / `False`, significant = 1).  Ids are
 * loop_<chrA>-<chrB>-<first_id + i>.  Returns the bytes written to out[0, cap) or a negative error code
 * (CL2_ERR_RANGE: cap too small; 640 bytes per row plus the names are enough).  cl2_format_floats writes repr(v[i])
 * followed by a newline for every value (the building block, exposed for tests).
 */
int64_t cl2_loops_text(const char *chrA, const char *chrB, int64_t first_id, int64_t n, int32_t cis, const int64_t *coords,
                       const int64_t *ra, const int64_t *rb, const int64_t *rab, const double *density, const double *es,
                       const double *p2ll, const double *fdr, const double *nbp, const double *hyp, const double *pop,
                       const double *px, const double *py, char *out, int64_t cap);
int64_t cl2_format_floats(const double *v, int64_t n, char *out, int64_t cap);

/*
 * Host copies of the fp64 tail probabilities the device evaluates (scipy argument convention, sf(k) = P(X > k)):
 * scipy.stats.poisson.sf / binom.sf / hypergeom.sf as called at callCisLoops.py:246,310,329,332.
 */
double cl2_stat_poisson_sf(double k, double mu);
double cl2_stat_binom_sf(double k, double n, double p);
double cl2_stat_hypergeom_sf(double k, double M, double n, double N);

/* ---- `cLoops2 pre`: BEDPE text -> records (host code, no GPU) -------------------------------------------- */
/*
 * The field handling of PET.__init__ (cLoops2/ds.py:42-57) for every whole line of buf[0, len): a line needs at least
 * 10 tab-separated fields and integer fields 1, 2, 4, 5, otherwise it is skipped as io.py:97-100 skips it.  For the
 * i-th accepted line: chrom[2i], chrom[2i+1] = ids of fields 0 and 3 (ids number the names in first-seen order over
 * the life of `p`), pos[4i..4i+3] = fields 1, 2, 4, 5, mapq[i] = field 7 or 255 when that is not an integer
 * (ds.py:54-57).  cap = rows the output arrays can hold.  Returns the number of accepted lines (>= 0) or an error
 * code; *n_lines (optional) = lines seen.
 */
int cl2_bedpe_create(cl2_bedpe **out);
void cl2_bedpe_destroy(cl2_bedpe *p);
int64_t cl2_bedpe_parse(cl2_bedpe *p, const char *buf, int64_t len, int64_t cap, int32_t *chrom, int64_t *pos,
                        int32_t *mapq, int64_t *n_lines);
/*
 * The same for .pairs text: the line handling of parsePairs / Pair.__init__ (cLoops2/io.py:189-207, ds.py:111-137).
 * A record has exactly 7 or 8 tab-separated fields, with 8 the pair type (field 7) must be UU, MR, MU, RU or UR,
 * fields 2 and 4 must be integers, '#' lines are headers.  chrom[2i], chrom[2i+1] = ids of fields 1 and 3,
 * pos[2i], pos[2i+1] = fields 2 and 4.
 */
int64_t cl2_pairs_parse(cl2_bedpe *p, const char *buf, int64_t len, int64_t cap, int32_t *chrom, int64_t *pos,
                        int64_t *n_lines);
/* Names seen so far, '\n'-joined in id order; valid until the next call on `p`. */
const char *cl2_bedpe_names(cl2_bedpe *p, int32_t *count);

/* How the library waits for its streams, process wide: 0 (default) spin in cudaStreamSynchronize, 1 sleep on a
 * blocking-sync event, 2 poll the stream and yield the core between polls -- for hosts with fewer cores than worker
 * threads (several ranks x several contexts). */
void cl2_set_sync_mode(int32_t mode);

/* ---- measurement -------------------------------------------------------------------------------------- */
/* Kernel launches issued by this context since creation (or since the last reset). */
int64_t cl2_launch_count(cl2_ctx *ctx);
void cl2_launch_count_reset(cl2_ctx *ctx);
/*
 * Per-kernel-family device time (CUDA events on the context's stream) accumulated while timing is enabled.
 * names_out: buffer of k_max * 32 chars (NUL-terminated names), ms_out / launches_out / bytes_out: k_max entries;
 * bytes_out is the ALGORITHMIC byte count of the launches (DESIGN.md states each kernel's per-unit figure).
 * Returns the number of families written.
 */
int cl2_timing_enable(cl2_ctx *ctx, int on);
/* Bytes this context has copied host->device and device->host since creation (or the last launch-count reset). */
void cl2_transfer_bytes(cl2_ctx *ctx, int64_t *h2d, int64_t *d2h);
int cl2_timing_read(cl2_ctx *ctx, int k_max, char *names_out, double *ms_out, int64_t *launches_out,
                    double *bytes_out);
void cl2_timing_reset(cl2_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
