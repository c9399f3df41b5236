// cl2_block.cu -- blockDBSCAN (cLoops2/blockDBSCAN.py:13-234) as an order-independent, data-parallel pipeline.
//
// The reference is a sequence of Python dict passes whose result depends on dict insertion order (= .ixy row order).
// The restatement used here (re-verified against the reference through oracle/) makes every
// order-dependent rule an explicit key:
//   cells      = segments of the points sorted by packed cell key (stable sort => first element of a segment carries
//                the smallest row id = the cell's dict position `first`)
//   noise      = cell removed iff its 3x3 population < minPts and the same holds for every existing neighbour
//   adjacency  = centroid L1 <= eps (fp64, exactly the reference's expression) or some cross pair with L1 <= eps
//   core cell  = cnt + sum(cnt of adjacent kept cells) >= minPts
//   clusters   = connected components of core cells (lock-free union-find), numbered by the rank of their first
//                cell in the stable sort by (-cnt, first)
//   border     = kept non-core cell: largest cluster whose START cell is adjacent, else smallest adjacent cluster
#include "cl2_points.cuh"
#include <algorithm>

#define NOT_FOUND (-1)
#define CL2_HOST_SORT_MAX 4096   // component counts up to this are ordered on the host (cheaper than ~15 launches)

__constant__ int c_dx[8] = {0, 0, -1, 1, -1, -1, 1, 1};
__constant__ int c_dy[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
// opposite direction index
__constant__ int c_opp[8] = {1, 0, 3, 2, 7, 6, 5, 4};

// ------------------------------------------------------------------------------------------------ K3 gather
__global__ void __launch_bounds__(256) k_gather_xy(const int2 *__restrict__ xy, const uint32_t *__restrict__ row,
                                                    int2 *__restrict__ sxy, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sxy[i] = __ldg(&xy[row[i]]);
}

// ------------------------------------------------------------------------------------------------ x-order from the grid
// The rows of a blockDBSCAN grid are stored column by column (column = (x - minX) / eps, ascending), so the x-order of
// all points -- the first array of the XY index, ds.py:147 -- is the grid order with every column sorted by x.  One
// thread per point ranks it inside its column by walking to both ends of the column:
//     rank = #{earlier rows of the column with x' <= x} + #{later rows with x' < x}
// (stable: equal x keep the grid's y order).  Columns of sparse data hold tens of rows and neighbouring threads walk
// the same rows, so this is one read and one write of the coordinates instead of a 4-pass radix sort plus a gather.
// A walk longer than XORDER_WALK_MAX gives up and raises *bad; cl2_index_build then sorts by radix.
#define XORDER_WALK_MAX 4096
#define XORDER_AVG_MAX 64      // attempted only when the average column holds at most this many rows
// One side of the walk.  q points at the x of row i, row i + DIR * k is q[2 * DIR * k]; w = rows on that side that may
// be visited, more = rows exist beyond them.  Adds to `rank` the visited rows of the column that sort before row i
// (left side: x' <= x, right side: x' < x), to `members` the visited rows of the column.  The column's rows are
// contiguous, so a row outside it ends the walk; four rows are tested per round (the kernel is bound by the
// instructions of these loops: ncu 83 % SM throughput at IPC 3.3, memory idle).  False: the walk limit was reached
// inside the column.
template <int DIR>
__device__ __forceinline__ bool xorder_walk(const int *__restrict__ q, int w, bool more, int lo, unsigned eps, int px,
                                            int &rank, int &members) {
    const int lim = DIR < 0 ? px : px - 1;                   // x' <= lim sorts before row i
    int k = 1;
    for (; k + 3 <= w; k += 4) {
        const int a = q[2 * DIR * k], b = q[2 * DIR * (k + 1)], c = q[2 * DIR * (k + 2)], d = q[2 * DIR * (k + 3)];
        const bool ia = (unsigned)(a - lo) < eps, ib = (unsigned)(b - lo) < eps, ic = (unsigned)(c - lo) < eps,
                   id = (unsigned)(d - lo) < eps;
        rank += (ia && a <= lim ? 1 : 0) + (ib && b <= lim ? 1 : 0) + (ic && c <= lim ? 1 : 0) + (id && d <= lim ? 1 : 0);
        members += (ia ? 1 : 0) + (ib ? 1 : 0) + (ic ? 1 : 0) + (id ? 1 : 0);
        if (!id) return true;
    }
    for (; k <= w; k++) {
        const int a = q[2 * DIR * k];
        if ((unsigned)(a - lo) >= eps) return true;
        rank += a <= lim ? 1 : 0;
        members++;
    }
    return !(more && (unsigned)(q[2 * DIR * (w + 1)] - lo) < eps);
}

__global__ void __launch_bounds__(256) k_xorder_from_grid(const int2 *__restrict__ sxy, int64_t n, int minx, unsigned eps,
                                                           int2 *__restrict__ out, int32_t *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*(volatile int32_t *)bad) return;                    // the result is void already
    const int2 p = sxy[i];
    const unsigned rel = (unsigned)(p.x - minx);
    const int lo = minx + (int)(rel - rel % eps);            // smallest x of the column; the column is [lo, lo + eps)
    const int *q = &sxy[i].x;
    const int64_t rest = n - 1 - i;
    const int wl = (int)(i < XORDER_WALK_MAX ? i : XORDER_WALK_MAX);
    const int wr = (int)(rest < XORDER_WALK_MAX ? rest : XORDER_WALK_MAX);
    int rank = 0, nleft = 0, nright = 0;
    if (!xorder_walk<-1>(q, wl, (int64_t)wl < i, lo, eps, p.x, rank, nleft) ||
        !xorder_walk<1>(q, wr, (int64_t)wr < rest, lo, eps, p.x, rank, nright)) {
        atomicExch(bad, 1);
        return;
    }
    out[i - nleft + rank] = p;                               // i - nleft = first row of the column
}

// ------------------------------------------------------------------------------------------------ K4 cells
// Cells are the runs of equal packed key in the sorted order; the key is recomputed from the sorted coordinates.
__global__ void __launch_bounds__(256) k_head_flags(const int2 *__restrict__ sxy, const cl2_keygen gen,
                                                     int32_t *__restrict__ flag, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || cl2_cell_key(gen, sxy[i]) != cl2_cell_key(gen, sxy[i - 1])) ? 1 : 0;
}

// cid holds the exclusive scan of the head flags on entry; on exit the cell id of every sorted position.
// first[c] = smallest file row of the cell (the reference's dict position of the cell, blockDBSCAN.py:83).
__global__ void __launch_bounds__(256) k_cell_table(const int2 *__restrict__ sxy, const cl2_keygen gen,
                                                     const uint32_t *__restrict__ row, int32_t *__restrict__ cid,
                                                     int32_t *__restrict__ cstart, uint64_t *__restrict__ ckey,
                                                     uint32_t *__restrict__ first, int64_t n, int32_t ncells) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = cl2_cell_key(gen, sxy[i]);
    bool head = (i == 0 || k != cl2_cell_key(gen, sxy[i - 1]));
    int32_t c = cid[i] + (head ? 1 : 0) - 1;
    cid[i] = c;
    if (head) {
        cstart[c] = (int32_t)i;
        ckey[c] = k;
    }
    atomicMin(&first[c], row[i]);
    if (i == n - 1) cstart[ncells] = (int32_t)n;
}

// ------------------------------------------------------------------------------------------------ K6 neighbours
// first index >= target, searching forward from c (exclusive) by galloping: rows of the grid are short, so the
// answer is almost always within a few cells of c.
__device__ __forceinline__ int lb_forward(const uint64_t *__restrict__ ckey, int ncells, int c, uint64_t target) {
    int lo = c + 1, hi = c + 1, step = 1;
    while (hi < ncells && ckey[hi] < target) {
        lo = hi + 1;
        hi += step;
        step <<= 1;
    }
    if (hi > ncells) hi = ncells;
    while (lo < hi) {
        int m = (lo + hi) >> 1;
        if (ckey[m] < target) lo = m + 1; else hi = m;
    }
    return lo;
}
// first index >= target, target < ckey[c], searching backward from c
__device__ __forceinline__ int lb_backward(const uint64_t *__restrict__ ckey, int c, uint64_t target) {
    int hi = c, lo = c - 1, step = 1;
    while (lo >= 0 && ckey[lo] >= target) {
        hi = lo;
        lo -= step;
        step <<= 1;
    }
    int l = (lo < 0 ? -1 : lo) + 1, h = hi;
    while (l < h) {
        int m = (l + h) >> 1;
        if (ckey[m] < target) l = m + 1; else h = m;
    }
    return l;
}

__global__ void __launch_bounds__(256) k_block_nbr(const uint64_t *__restrict__ ckey,
                                                    const int32_t *__restrict__ cstart, int32_t *__restrict__ nbr,
                                                    int32_t *__restrict__ s3, int32_t ncells, int by) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) {
        const uint64_t k = ckey[c];
        const uint64_t ymask = (1ull << by) - 1;
        const long long gx = (long long)(k >> by), gy = (long long)(k & ymask);
        int out[8];
#pragma unroll
        for (int d = 0; d < 8; d++) out[d] = NOT_FOUND;
        // same row
        if (gy > 0 && c > 0 && ckey[c - 1] == k - 1) out[0] = c - 1;
        if (gy < (long long)ymask && c + 1 < ncells && ckey[c + 1] == k + 1) out[1] = c + 1;
        // row gx+1: wanted keys (gx+1, gy-1..gy+1)
        {
            long long y0 = gy > 0 ? gy - 1 : 0;
            uint64_t t0 = ((uint64_t)(gx + 1) << by) | (uint64_t)y0;
            int p = lb_forward(ckey, ncells, c, t0);
            for (int q = p; q < ncells && q < p + 3; q++) {
                uint64_t kk = ckey[q];
                if ((long long)(kk >> by) != gx + 1) break;
                long long yy = (long long)(kk & ymask);
                if (yy == gy - 1) out[6] = q;
                else if (yy == gy) out[3] = q;
                else if (yy == gy + 1) out[7] = q;
                else break;
            }
        }
        if (gx > 0) {
            long long y0 = gy > 0 ? gy - 1 : 0;
            uint64_t t0 = ((uint64_t)(gx - 1) << by) | (uint64_t)y0;
            int p = lb_backward(ckey, c, t0);
            for (int q = p; q < c && q < p + 3; q++) {
                uint64_t kk = ckey[q];
                if ((long long)(kk >> by) != gx - 1) break;
                long long yy = (long long)(kk & ymask);
                if (yy == gy - 1) out[4] = q;
                else if (yy == gy) out[2] = q;
                else if (yy == gy + 1) out[5] = q;
                else break;
            }
        }
        int s = cstart[c + 1] - cstart[c];
#pragma unroll
        for (int d = 0; d < 8; d++) {
            if (out[d] >= 0) s += cstart[out[d] + 1] - cstart[out[d]];
        }
        int4 *o = reinterpret_cast<int4 *>(nbr + 8 * (int64_t)c);
        o[0] = make_int4(out[0], out[1], out[2], out[3]);
        o[1] = make_int4(out[4], out[5], out[6], out[7]);
        s3[c] = s;
    }
}

// ------------------------------------------------------------------------------------------------ K7 noise
// keep[c] = 0 iff s3[c] < minPts and every existing neighbour has s3 < minPts (blockDBSCAN.py:101-122)
__global__ void __launch_bounds__(256) k_block_noise(const int32_t *__restrict__ nbr, const int32_t *__restrict__ s3,
                                                      uint8_t *__restrict__ keep, int32_t ncells, int minPts,
                                                      int32_t *__restrict__ keep_i32 = nullptr) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    bool k = true;
    if (s3[c] < minPts) {
        k = false;
        const int4 *o = reinterpret_cast<const int4 *>(nbr + 8 * (int64_t)c);
        int4 a = o[0], b = o[1];
        int nb[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int d = 0; d < 8; d++)
            if (nb[d] >= 0 && s3[nb[d]] >= minPts) k = true;
    }
    keep[c] = k ? 1 : 0;
    if (keep_i32) keep_i32[c] = k ? 1 : 0;       // the same flag as scan input (compact list of kept cells)
}

__global__ void __launch_bounds__(256) k_keep_points(const int32_t *__restrict__ cid, const uint32_t *__restrict__ row,
                                                      const uint8_t *__restrict__ keep,
                                                      const uint32_t *__restrict__ first, uint8_t *__restrict__ out,
                                                      int32_t *__restrict__ first_out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cid[i];
    out[row[i]] = keep[c];
    if (first_out) first_out[row[i]] = (int32_t)first[c];
}

// ------------------------------------------------------------------------------------------------ K8 cell stats
struct CellStat {
    long long sx, sy;        // coordinate sums (centroid = sum / cnt in fp64, blockDBSCAN.py:131-137)
    int xmin, xmax, ymin, ymax;
};

// compact list of kept cells (so later kernels launch one unit per kept cell)
__global__ void __launch_bounds__(256) k_keep_list(const uint8_t *__restrict__ keep, const int32_t *__restrict__ pos,
                                                    int32_t *__restrict__ list, int32_t ncells) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && keep[c]) list[pos[c]] = c;
}
#define CELLSTAT_SERIAL_MAX 48
// kept cells with few points (the common case): one thread per cell, consecutive threads read consecutive segments;
// crowded cells are only listed here and summed by k_block_cellstat_list
__global__ void __launch_bounds__(256) k_block_cellstat_small(const int2 *__restrict__ sxy,
                                                               const int32_t *__restrict__ cstart,
                                                               const int32_t *__restrict__ list,
                                                               CellStat *__restrict__ st, int32_t nkept,
                                                               int32_t *__restrict__ crowd, int32_t *__restrict__ ncrowd) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    int b = cstart[c], e = cstart[c + 1];
    if (e - b > CELLSTAT_SERIAL_MAX) {
        crowd[atomicAdd(ncrowd, 1)] = c;
        return;
    }
    long long sx = 0, sy = 0;
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    for (int i = b; i < e; i++) {
        int2 p = sxy[i];
        sx += p.x; sy += p.y;
        xmin = min(xmin, p.x); xmax = max(xmax, p.x);
        ymin = min(ymin, p.y); ymax = max(ymax, p.y);
    }
    CellStat s;
    s.sx = sx; s.sy = sy; s.xmin = xmin; s.xmax = xmax; s.ymin = ymin; s.ymax = ymax;
    st[c] = s;
}
// crowded cells: one warp per listed cell, the warps of a fixed-size grid stride over the list (its length stays on
// the device)
__global__ void __launch_bounds__(256) k_block_cellstat_list(const int2 *__restrict__ sxy,
                                                              const int32_t *__restrict__ cstart,
                                                              const int32_t *__restrict__ crowd,
                                                              const int32_t *__restrict__ ncrowd,
                                                              CellStat *__restrict__ st) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int m = *ncrowd;
    for (int k = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); k < m; k += nwarps) {
        int c = crowd[k];
        int b = cstart[c], e = cstart[c + 1];
        long long sx = 0, sy = 0;
        int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
        for (int i = b + lane; i < e; i += 32) {
            int2 p = sxy[i];
            sx += p.x; sy += p.y;
            xmin = min(xmin, p.x); xmax = max(xmax, p.x);
            ymin = min(ymin, p.y); ymax = max(ymax, p.y);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        if (lane == 0) {
            CellStat s;
            s.sx = sx; s.sy = sy; s.xmin = xmin; s.xmax = xmax; s.ymin = ymin; s.ymax = ymax;
            st[c] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K9 adjacency
// Three tiers by the number of point pairs |a|*|b| of a cell pair: up to ADJ_LIGHT_MAX one thread tests them (the bulk
// of sparse data: cells of 1-5 points); up to ADJ_WARP_MAX one warp does (k_block_adj_warp: the cells of a loop's own
// cluster hold tens of points each, and a thread working through 1600 pairs alone keeps its whole warp waiting);
// beyond that a block with shared-memory tiles (k_block_adj_heavy: crowded cells of Hi-C-like data).
#define ADJ_LIGHT_MAX 48
#define ADJ_WARP_MAX 65536

__device__ __forceinline__ bool centroid_close(const CellStat &a, int na, const CellStat &b, int nb, double eps) {
    // blockDBSCAN.py:136-137 (x / len) and :46,:227 (abs(dx)+abs(dy) <= eps), fp64, no contraction possible
    double ax = (double)a.sx / (double)na, ay = (double)a.sy / (double)na;
    double bx = (double)b.sx / (double)nb, by = (double)b.sy / (double)nb;
    return fabs(ax - bx) + fabs(ay - by) <= eps;
}

// forward directions (neighbour index > own index in key order): (0,+1)=1, (+1,-1)=6, (+1,0)=3, (+1,+1)=7
// One thread per kept cell tests its four forward neighbours.  The kernel waits on chains of dependent loads (list ->
// neighbour ids -> the neighbours' tables -> their points; ncu: 72 % of the stall samples are long-scoreboard), so the
// loads of every level are issued for all four neighbours before the first is used.
__global__ void __launch_bounds__(128, 6) k_block_adj(const int2 *__restrict__ sxy, const int32_t *__restrict__ cstart,
                                                   const int32_t *__restrict__ nbr, const uint8_t *__restrict__ keep,
                                                   const CellStat *__restrict__ st, const int32_t *__restrict__ list,
                                                   int32_t nkept, unsigned *__restrict__ adj, int64_t eps,
                                                   int2 *__restrict__ heavy, int32_t *__restrict__ nheavy,
                                                   int2 *__restrict__ medium, int32_t *__restrict__ nmedium) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    const int fdir[4] = {1, 6, 3, 7};
    const int a = list[k];
    const int4 *nb4 = reinterpret_cast<const int4 *>(nbr + 8 * (int64_t)a);
    const int4 n0 = nb4[0], n1 = nb4[1];
    const int a0 = cstart[a], a1 = cstart[a + 1];
    const CellStat A = st[a];
    const int bq[4] = {n0.y, n1.z, n0.w, n1.w};              // directions 1, 6, 3, 7
    int b0[4], b1[4];
    uint8_t kb[4];
    CellStat B[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int bb = bq[q] < 0 ? a : bq[q];                // clamped so the loads are unconditional
        kb[q] = keep[bb];
        b0[q] = cstart[bb];
        b1[q] = cstart[bb + 1];
        B[q] = st[bb];
    }
    const int na = a1 - a0;
    unsigned mine = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int b = bq[q], d = fdir[q];
        if (b < 0 || !kb[q]) continue;
        const int nb = b1[q] - b0[q];
        bool close = centroid_close(A, na, B[q], nb, (double)eps);
        if (!close) {
            // bounding boxes further apart than eps in L1 can hold no close pair
            long long gx = max(0, max(A.xmin - B[q].xmax, B[q].xmin - A.xmax));
            long long gy = max(0, max(A.ymin - B[q].ymax, B[q].ymin - A.ymax));
            if (gx + gy > eps) continue;
            if ((long long)na * nb > ADJ_LIGHT_MAX) {
                if ((long long)na * nb > ADJ_WARP_MAX) heavy[atomicAdd(nheavy, 1)] = make_int2(a, d);
                else medium[atomicAdd(nmedium, 1)] = make_int2(a, d);
                continue;
            }
            for (int i = a0; i < a1 && !close; i++) {
                int2 p = sxy[i];
                for (int j = b0[q]; j < b1[q]; j++) {
                    int2 qq = sxy[j];
                    long long dist = (long long)abs(p.x - qq.x) + (long long)abs(p.y - qq.y);
                    if (dist <= eps) { close = true; break; }
                }
            }
        }
        if (close) {
            mine |= 1u << d;
            atomicOr(&adj[b], 1u << c_opp[d]);
        }
    }
    if (mine) atomicOr(&adj[a], mine);
}

// one warp per listed pair (the warps of a fixed-size grid stride over the list, whose length stays on the device): the
// lanes hold points of the larger cell, the points of the smaller one are read by all lanes together (one broadcast
// load each) and a ballot ends the pair on the first hit
__device__ __forceinline__ void adj_warp_pairs(const int2 *__restrict__ sxy, const int32_t *__restrict__ cstart,
                                               const int32_t *__restrict__ nbr, const int2 *__restrict__ medium,
                                               const int32_t *__restrict__ nmedium, unsigned *__restrict__ adj,
                                               int64_t eps) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n = *nmedium;
    for (int h = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); h < n; h += nwarps) {
        const int2 pr = medium[h];
        const int a = pr.x, d = pr.y;
        const int b = nbr[8 * (int64_t)a + d];
        int l0 = cstart[a], l1 = cstart[a + 1], s0 = cstart[b], s1 = cstart[b + 1];
        if (l1 - l0 < s1 - s0) { int t0 = l0, t1 = l1; l0 = s0; l1 = s1; s0 = t0; s1 = t1; }     // lanes take the larger cell
        bool found = false;
        for (int i0 = l0; i0 < l1 && !found; i0 += 32) {
            const bool valid = i0 + lane < l1;
            const int2 p = valid ? sxy[i0 + lane] : make_int2(0, 0);
            for (int j = s0; j < s1; j++) {
                const int2 q = sxy[j];
                const long long dist = (long long)abs(p.x - q.x) + (long long)abs(p.y - q.y);
                if (__any_sync(0xffffffffu, valid && dist <= eps)) { found = true; break; }
            }
        }
        if (found && lane == 0) {
            atomicOr(&adj[a], 1u << d);
            atomicOr(&adj[b], 1u << c_opp[d]);
        }
    }
}

// one block per heavy pair: tile b's points through shared memory, all threads scan a's points against the tile
#define HEAVY_THREADS 256
#define HEAVY_TILE 512
__global__ void __launch_bounds__(HEAVY_THREADS) k_block_adj_heavy(const int2 *__restrict__ sxy,
                                                                    const int32_t *__restrict__ cstart,
                                                                    const int32_t *__restrict__ nbr,
                                                                    const int2 *__restrict__ heavy,
                                                                    const int32_t *__restrict__ nheavy,
                                                                    const int2 *__restrict__ medium,
                                                                    const int32_t *__restrict__ nmedium,
                                                                    unsigned *__restrict__ adj, int64_t eps) {
    __shared__ int2 tile[HEAVY_TILE];
    __shared__ int found;
    adj_warp_pairs(sxy, cstart, nbr, medium, nmedium, adj, eps);          // the medium pairs, one warp each
    const int nh = *nheavy;
    for (int hi = blockIdx.x; hi < nh; hi += gridDim.x) {
    __syncthreads();                       // `found` and the tile of the previous pair are done with
    int2 h = heavy[hi];
    int a = h.x, d = h.y;
    int b = nbr[8 * (int64_t)a + d];
    int a0 = cstart[a], a1 = cstart[a + 1], b0 = cstart[b], b1 = cstart[b + 1];
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    for (int tb = b0; tb < b1; tb += HEAVY_TILE) {
        int tn = min(HEAVY_TILE, b1 - tb);
        for (int j = threadIdx.x; j < tn; j += HEAVY_THREADS) tile[j] = sxy[tb + j];
        __syncthreads();
        for (int i = a0 + threadIdx.x; i < a1; i += HEAVY_THREADS) {
            int2 p = sxy[i];
            bool hit = false;
            for (int j = 0; j < tn; j++) {
                int2 q = tile[j];
                long long dist = (long long)abs(p.x - q.x) + (long long)abs(p.y - q.y);
                if (dist <= eps) { hit = true; break; }
                // another thread already found a pair: only shortens this thread's scan (any hit decides the pair)
                if ((j & 7) == 7 && *(volatile int *)&found) break;
            }
            if (hit) { found = 1; break; }
            if (*(volatile int *)&found) break;
        }
        __syncthreads();
        int f = found;          // uniform: written only between the two barriers above
        if (f) break;
    }
    __syncthreads();
    if (threadIdx.x == 0 && found) {
        atomicOr(&adj[a], 1u << d);
        atomicOr(&adj[b], 1u << c_opp[d]);
    }
    }
}

// ------------------------------------------------------------------------------------------------ K10 core
// near = cnt + sum of adjacent kept cells' cnt; core iff near >= minPts (blockDBSCAN.py:179-189).
// okey orders cells like the stable sort by -cnt of blockDBSCAN.py:147 (ties: dict position = first row id).
__global__ void __launch_bounds__(256) k_block_core(const int32_t *__restrict__ cstart, const int32_t *__restrict__ nbr,
                                                     const unsigned *__restrict__ adj, const uint32_t *__restrict__ first,
                                                     const int32_t *__restrict__ list, int32_t nkept, int minPts,
                                                     uint8_t *__restrict__ core,
                                                     uint64_t *__restrict__ okey, int32_t *__restrict__ parent,
                                                     uint64_t *__restrict__ minkey) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    int cnt = cstart[c + 1] - cstart[c];
    long long near = cnt;
    unsigned m = adj[c];
    while (m) {
        int d = __ffs(m) - 1;
        m &= m - 1;
        int b = nbr[8 * (int64_t)c + d];
        near += cstart[b + 1] - cstart[b];
    }
    core[c] = near >= minPts ? 1 : 0;
    okey[c] = ((uint64_t)(uint32_t)(0x7fffffff - cnt) << 32) | (uint64_t)first[c];     // larger cells first, ties by dict position
    parent[c] = c;
    minkey[c] = ~0ull;
}

// ------------------------------------------------------------------------------------------------ K11 union-find
__device__ __forceinline__ int uf_find(int32_t *parent, int x) {
    int p = parent[x];
    while (p != x) {
        int gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving; benign race (only ever moves a node closer to its root)
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_unite(int32_t *parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        int old = atomicCAS(&parent[b], b, a);   // hook the larger root under the smaller one
        if (old == b) return;
    }
}

__global__ void __launch_bounds__(256) k_block_union(const int32_t *__restrict__ nbr, const unsigned *__restrict__ adj,
                                                      const uint8_t *__restrict__ core, const int32_t *__restrict__ list,
                                                      int32_t nkept, int32_t *__restrict__ parent) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    if (!core[c]) return;
    unsigned m = adj[c] & ((1u << 1) | (1u << 6) | (1u << 3) | (1u << 7));
    while (m) {
        int d = __ffs(m) - 1;
        m &= m - 1;
        int b = nbr[8 * (int64_t)c + d];
        if (core[b]) uf_unite(parent, c, b);
    }
}

// root of every core cell + per-root minimum order key (the component's start cell, blockDBSCAN.py:146-152)
__global__ void __launch_bounds__(256) k_block_roots(const uint8_t *__restrict__ core, const int32_t *__restrict__ list,
                                                      int32_t nkept, int32_t *__restrict__ parent,
                                                      const uint64_t *__restrict__ okey, uint64_t *__restrict__ minkey,
                                                      int32_t *__restrict__ root) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    if (!core[c]) { root[c] = -1; return; }
    int r = uf_find(parent, c);
    root[c] = r;
    atomicMin((unsigned long long *)&minkey[r], (unsigned long long)okey[c]);
}

__global__ void __launch_bounds__(256) k_block_rootflags(const uint8_t *__restrict__ core,
                                                          const int32_t *__restrict__ root,
                                                          const int32_t *__restrict__ list, int32_t nkept,
                                                          int32_t *__restrict__ flag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    flag[k] = (core[c] && root[c] == c) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_block_rootlist(const uint8_t *__restrict__ core,
                                                         const int32_t *__restrict__ root,
                                                         const int32_t *__restrict__ list, int32_t nkept,
                                                         const int32_t *__restrict__ pos,
                                                         const uint64_t *__restrict__ minkey,
                                                         uint64_t *__restrict__ ckeys, int32_t *__restrict__ compid) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    if (core[c] && root[c] == c) {
        ckeys[pos[k]] = minkey[c];      // order key of the component's start cell
        compid[c] = pos[k];             // provisional id (cell-key order); the final numbering is a host sort of ckeys
    }
}
// ------------------------------------------------------------------------------------------------ K14 labels
__global__ void __launch_bounds__(256) k_block_celllabel(const int32_t *__restrict__ nbr,
                                                          const unsigned *__restrict__ adj,
                                                          const uint8_t *__restrict__ core,
                                                          const int32_t *__restrict__ root,
                                                          const int32_t *__restrict__ compid,
                                                          const uint64_t *__restrict__ okey,
                                                          const uint64_t *__restrict__ minkey,
                                                          const int32_t *__restrict__ list, int32_t nkept,
                                                          int32_t *__restrict__ clabel) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkept) return;
    int c = list[k];
    int lab = -1;
    if (core[c]) {
        lab = compid[root[c]];
    } else {
        // components are compared by the order key of their start cell (= the reference's cluster numbering)
        unsigned long long maxstart = 0, minany = ~0ull;
        int rmax = -1, rmin = -1;
        unsigned m = adj[c];
        while (m) {
            int d = __ffs(m) - 1;
            m &= m - 1;
            int b = nbr[8 * (int64_t)c + d];
            if (!core[b]) continue;
            int r = root[b];
            unsigned long long key = minkey[r];
            if (okey[b] == key && (rmax < 0 || key > maxstart)) { maxstart = key; rmax = r; }   // unconditional overwrite, blockDBSCAN.py:184-185
            if (key < minany) { minany = key; rmin = r; }                                        // first claim wins, blockDBSCAN.py:193-195
        }
        lab = rmax >= 0 ? compid[rmax] : (rmin >= 0 ? compid[rmin] : -1);
    }
    clabel[c] = lab;
}

__global__ void __launch_bounds__(256) k_point_labels(const int32_t *__restrict__ cid, const uint32_t *__restrict__ row,
                                                       const int32_t *__restrict__ clabel,
                                                       const int32_t *__restrict__ final_of_prov,
                                                       int32_t *__restrict__ labels, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l = clabel[cid[i]];
    labels[row[i]] = l >= 0 ? final_of_prov[l] : -1;
}

// ------------------------------------------------------------------------------------------------ K16 bbox
struct ClusterBox {
    int xmin, xmax, ymin, ymax;
    unsigned size;
};
__global__ void __launch_bounds__(256) k_box_init(ClusterBox *__restrict__ box, int32_t ncl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncl) {
        ClusterBox b;
        b.xmin = INT_MAX; b.xmax = INT_MIN; b.ymin = INT_MAX; b.ymax = INT_MIN; b.size = 0;
        box[i] = b;
    }
}
// Cells are visited in key order, so the cells of a large cluster (a diagonal band of Hi-C-like data holds most of
// them) arrive warp after warp with the same label: the lanes of a warp that share a label combine their boxes with a
// masked warp reduction and their first lane issues the five atomics, instead of 32 atomics on one address.
__global__ void __launch_bounds__(256) k_block_bbox(const int32_t *__restrict__ cstart,
                                                     const CellStat *__restrict__ st,
                                                     const int32_t *__restrict__ clabel,
                                                     const int32_t *__restrict__ list, int32_t nkept,
                                                     ClusterBox *__restrict__ box) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int lab = -1;
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    unsigned size = 0;
    if (k < nkept) {
        int c = list[k];
        lab = clabel[c];
        if (lab >= 0) {
            const CellStat s = st[c];
            xmin = s.xmin; xmax = s.xmax; ymin = s.ymin; ymax = s.ymax;
            size = (unsigned)(cstart[c + 1] - cstart[c]);
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, lab);
    xmin = __reduce_min_sync(peers, xmin);
    xmax = __reduce_max_sync(peers, xmax);
    ymin = __reduce_min_sync(peers, ymin);
    ymax = __reduce_max_sync(peers, ymax);
    size = __reduce_add_sync(peers, size);
    if (lab < 0 || (int)(threadIdx.x & 31) != __ffs(peers) - 1) return;
    ClusterBox *b = &box[lab];
    atomicMin(&b->xmin, xmin);
    atomicMax(&b->xmax, xmax);
    atomicMin(&b->ymin, ymin);
    atomicMax(&b->ymax, ymax);
    atomicAdd(&b->size, size);
}

__global__ void __launch_bounds__(256) k_fill_i32(int32_t *__restrict__ p, int32_t v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ================================================================================================ host driver
int cl2_read_i64(cl2_ctx *ctx, const int64_t *d, int64_t *out) {
    CL2_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = *(int64_t *)ctx->h_scratch;
    return CL2_OK;
}
int cl2_read_i32(cl2_ctx *ctx, const int32_t *d, int32_t *out) {
    CL2_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = *(int32_t *)ctx->h_scratch;
    return CL2_OK;
}

// Build (or reuse) the eps grid: keygen -> radix sort -> gather -> cell table -> neighbours + 3x3 sums.
// kind 0: blockDBSCAN grid anchored at (minX, minY) (blockDBSCAN.py:69-86); kind 1: cDBSCAN2 rotated grid.
static int64_t gcd64(int64_t a, int64_t b) {
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}
// may a run (eps, minPts) work on the rows of `sub` alone?  (cl2_points.cuh: cl2_subset)
static bool subset_usable(const cl2_subset &sub, int64_t eps, int32_t minPts) {
    if (sub.n <= 0 || sub.nlinks == 0 || minPts < 0) return false;
    for (int i = 0; i < sub.nlinks; i++) {
        const int64_t E = sub.link_eps[i];
        if (eps >= E || minPts < sub.link_minpts[i]) return false;
        if (E < 3 * eps - gcd64(eps, E)) return false;
    }
    return true;
}

int cl2_grid_build(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int kind, int32_t minPts) {
    cl2_grid_cache &g = pts->grid;
    if (g.valid && g.kind == kind && g.eps == eps && (g.min_minpts < 0 || (minPts >= 0 && minPts >= g.min_minpts)))
        return CL2_OK;
    cl2_grid_free(ctx, g);
    static const bool no_subset = getenv("CL2_NO_SUBSET") != nullptr;
    int level = -1;                                           // deepest (smallest) usable subset
    if (kind == 0 && pts->sweep_hint && !no_subset)
        for (int k = (int)pts->subs.size() - 1; k >= 0 && level < 0; k--)
            if (subset_usable(pts->subs[k], eps, minPts)) level = k;
    const bool use_sub = level >= 0;
    const cl2_subset *sub = use_sub ? &pts->subs[level] : nullptr;
    const int64_t n = use_sub ? sub->n : pts->n;
    unsigned e32 = eps >= (int64_t)INT32_MAX ? (unsigned)INT32_MAX : (unsigned)eps;
    uint64_t gxmax, gymax;
    int rgx0 = 0;
    unsigned rgy0 = 0;
    if (kind == 0) {
        gxmax = (uint64_t)(pts->ext[1] - pts->ext[0]) / e32;
        gymax = (uint64_t)(pts->ext[3] - pts->ext[2]) / e32;
    } else {
        long long ulo = pts->ext[0] - pts->ext[3], uhi = pts->ext[1] - pts->ext[2];   // bounds of X-Y
        long long vlo = pts->ext[0] + pts->ext[2], vhi = pts->ext[1] + pts->ext[3];   // bounds of X+Y
        rgx0 = (int)(ulo / (long long)e32);
        rgy0 = (unsigned)(vlo / (long long)e32);
        gxmax = (uint64_t)(uhi / (long long)e32 - ulo / (long long)e32);
        gymax = (uint64_t)(vhi / (long long)e32 - vlo / (long long)e32);
    }
    int bx = bits_for(gxmax), by = bits_for(gymax);
    Scratch sc(ctx);
    uint64_t *k0, *k1, *ks;
    uint32_t *v0, *v1, *vs;
    CL2_TRY(sc.get(&k0, (size_t)n));
    CL2_TRY(sc.get(&k1, (size_t)n));
    CL2_TRY(sc.get_pool(&v0, (size_t)n));           // one of the two is kept as g.row
    CL2_TRY(sc.get_pool(&v1, (size_t)n));
    cl2_keygen gen;
    gen.kind = kind == 0 ? CL2_KEY_BLOCK : CL2_KEY_ROT;
    gen.xy = pts->xy;
    gen.minx = (int)pts->ext[0];
    gen.miny = (int)pts->ext[2];
    gen.eps = e32;
    gen.by = by;
    gen.gx0 = rgx0;
    gen.gy0 = rgy0;
    if (kind == 0) {
        // (column, row) order = stable sort by column of the rows already sorted by y: the row index (y-miny)/eps is
        // monotone in y, so only the bx column bits are sorted (3 passes for hg38 at eps >= 100 instead of 5-6)
        if (!use_sub) CL2_TRY(cl2_points_yorder(ctx, pts));
        cl2_keygen gcol = gen;
        gcol.kind = CL2_KEY_GX_YX;
        gcol.xy = use_sub ? sub->byy : pts->byy;
        CL2_TRY(cl2_radix_sort_gen(ctx, gcol, k0, k1, v0, v1, n, bx, &ks, &vs, use_sub ? sub->rowy : pts->rowy));
    } else {
        CL2_TRY(cl2_radix_sort_gen(ctx, gen, k0, k1, v0, v1, n, bx + by, &ks, &vs));
    }
    // keep the sorted row ids (take ownership of whichever buffer holds them)
    sc.release(vs);
    g.row = vs;
    CL2_TRY(cl2_dev_alloc(ctx, &g.sxy, (size_t)n));
    CL2_TRY(cl2_dev_alloc(ctx, &g.cid, (size_t)n));
    {
        FamTimer t(ctx, FAM_GATHER, 20.0 * (double)n);
        k_gather_xy<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(pts->xy, g.row, g.sxy, n);
    }
    // an index build was announced (cl2_points_index_hint, cl2_call_loops): leave the x-order behind while the grid
    // order of every row is at hand.  CL2_XGRID=0 switches this off (the index then sorts by radix as before).
    static const bool xgrid_on = !(getenv("CL2_XGRID") && atoi(getenv("CL2_XGRID")) == 0);
    if (kind == 0 && !use_sub && pts->want_xorder && !pts->byx && xgrid_on && n == pts->n &&
        (uint64_t)n <= (uint64_t)XORDER_AVG_MAX * (gxmax + 1)) {
        CL2_TRY(cl2_dev_alloc(ctx, &pts->byx, (size_t)n));
        CL2_TRY(cl2_dev_alloc(ctx, &pts->xorder_bad, 1));
        CL2_CUDA(ctx, cudaMemsetAsync(pts->xorder_bad, 0, 4, ctx->stream));
        FamTimer t(ctx, FAM_INDEX, 16.0 * (double)n);
        k_xorder_from_grid<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(g.sxy, n, (int)pts->ext[0], e32, pts->byx,
                                                                     pts->xorder_bad);
    }
    int64_t *d_total;
    CL2_TRY(sc.get(&d_total, 1));
    {
        FamTimer t(ctx, FAM_CELLS, 12.0 * (double)n);
        k_head_flags<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(g.sxy, gen, g.cid, n);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, g.cid, n, d_total));
    int64_t ncells64 = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &ncells64));
    int32_t ncells = (int32_t)ncells64;
    CL2_TRY(cl2_dev_alloc(ctx, &g.cstart, (size_t)ncells + 1));
    CL2_TRY(cl2_dev_alloc(ctx, &g.ckey, (size_t)ncells));
    CL2_TRY(cl2_dev_alloc(ctx, &g.nbr, (size_t)ncells * 8));
    CL2_TRY(cl2_dev_alloc(ctx, &g.s3, (size_t)ncells));
    CL2_TRY(cl2_dev_alloc(ctx, &g.first, (size_t)ncells));
    CL2_CUDA(ctx, cudaMemsetAsync(g.first, 0xff, (size_t)ncells * 4, ctx->stream));
    {
        FamTimer t(ctx, FAM_CELLS, 20.0 * (double)n);
        k_cell_table<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(g.sxy, gen, g.row, g.cid, g.cstart, g.ckey, g.first, n, ncells);
    }
    {
        FamTimer t(ctx, FAM_NBR, 52.0 * (double)ncells);
        k_block_nbr<<<cl2_grid(ncells, 256), 256, 0, ctx->stream>>>(g.ckey, g.cstart, g.nbr, g.s3, ncells, by);
    }
    CL2_CUDA(ctx, cudaGetLastError());
    g.valid = true;
    g.kind = kind;
    g.eps = eps;
    g.n = n;
    g.min_minpts = -1;
    g.sub_level = level;
    if (use_sub)
        for (int i = 0; i < sub->nlinks; i++) g.min_minpts = std::max(g.min_minpts, sub->link_minpts[i]);
    g.ncells = ncells;
    g.by = by;
    return CL2_OK;
}

// ---- rows outside the removed noise cells, in y order (cl2_subset) ---------------------------------------------
__global__ void __launch_bounds__(256) k_mark_rows(const int32_t *__restrict__ cid, const uint32_t *__restrict__ row,
                                                    const uint8_t *__restrict__ keep, uint8_t *__restrict__ rowflag,
                                                    int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rowflag[row[i]] = keep[cid[i]];
}
__global__ void __launch_bounds__(256) k_sub_flags(const uint32_t *__restrict__ rowy, const uint8_t *__restrict__ rowflag,
                                                    int32_t *__restrict__ f, int64_t n) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) f[j] = rowflag[rowy[j]];
}
__global__ void __launch_bounds__(256) k_sub_compact(const int2 *__restrict__ byy, const uint32_t *__restrict__ rowy,
                                                      const int32_t *__restrict__ pos, int64_t n, int64_t total,
                                                      int2 *__restrict__ oby, uint32_t *__restrict__ orow) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int32_t p = pos[j];
    const bool kept = (j + 1 < n) ? (pos[j + 1] != p) : ((int64_t)p != total);
    if (kept) {
        oby[p] = byy[j];
        orow[p] = rowy[j];
    }
}
// After the noise step of a run (eps, minPts) on grid g: remember the rows of the kept cells when that shrinks the
// working set enough to pay for itself.  keep: per cell of g.
static int subset_update(cl2_ctx *ctx, cl2_points *pts, const uint8_t *keep, int64_t eps, int32_t minPts) {
    cl2_grid_cache &g = pts->grid;
    const int level = g.sub_level;                           // the subset this run worked on (-1: every point)
    if (level < 0 && !pts->subs.empty()) return CL2_OK;      // a chain exists and this run is not part of it
    if (level >= (int)pts->subs.size()) return CL2_OK;
    cl2_subset link;                                         // links of the new subset = its parent's + this run
    if (level >= 0) {
        const cl2_subset &par = pts->subs[level];
        if (par.nlinks >= CL2_SUB_MAXLINKS) return CL2_OK;
        link.nlinks = par.nlinks;
        for (int i = 0; i < par.nlinks; i++) { link.link_eps[i] = par.link_eps[i]; link.link_minpts[i] = par.link_minpts[i]; }
        if ((int)pts->subs.size() > level + 1) {
            // a child of the same parent exists: keep the one made at the larger eps (valid for more later runs)
            const cl2_subset &ch = pts->subs[level + 1];
            if (ch.link_eps[ch.nlinks - 1] >= eps) return CL2_OK;
        }
    }
    link.link_eps[link.nlinks] = eps;
    link.link_minpts[link.nlinks] = minPts;
    link.nlinks++;
    const int64_t nsrc = g.n;
    const int2 *sby = level >= 0 ? pts->subs[level].byy : pts->byy;
    const uint32_t *srow = level >= 0 ? pts->subs[level].rowy : pts->rowy;
    if (!sby || !srow || nsrc >= (int64_t)INT32_MAX) return CL2_OK;
    Scratch sc(ctx);
    uint8_t *rowflag;
    int32_t *f;
    int64_t *d_total;
    CL2_TRY(sc.get(&rowflag, (size_t)pts->n));
    CL2_TRY(sc.get(&f, (size_t)nsrc));
    CL2_TRY(sc.get(&d_total, 1));
    {
        FamTimer t(ctx, FAM_NOISE, 18.0 * (double)nsrc, 2);
        k_mark_rows<<<cl2_grid(nsrc, 256), 256, 0, ctx->stream>>>(g.cid, g.row, keep, rowflag, nsrc);
        k_sub_flags<<<cl2_grid(nsrc, 256), 256, 0, ctx->stream>>>(srow, rowflag, f, nsrc);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, f, nsrc, d_total));
    int64_t total = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &total));
    if (total <= 0 || total > nsrc - nsrc / 8) return CL2_OK;   // nothing left, or less than 1/8 dropped: not worth it
    int2 *oby = nullptr;
    uint32_t *orow = nullptr;
    CL2_TRY(cl2_dev_alloc(ctx, &oby, (size_t)total));
    int rc = cl2_dev_alloc(ctx, &orow, (size_t)total);
    if (rc) { cl2_dev_free(ctx, oby); return rc; }
    {
        FamTimer t(ctx, FAM_NOISE, 16.0 * (double)nsrc + 12.0 * (double)total);
        k_sub_compact<<<cl2_grid(nsrc, 256), 256, 0, ctx->stream>>>(sby, srow, f, nsrc, total, oby, orow);
    }
    cl2_subsets_free(ctx, pts, (size_t)(level + 1));          // deeper levels hang off a sibling: gone with it
    link.byy = oby;
    link.rowy = orow;
    link.n = total;
    pts->subs.push_back(link);
    return CL2_OK;
}

static int block_check_args(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts) {
    if (!ctx || !pts) return CL2_ERR_ARG;
    if (eps <= 0) return cl2_fail(ctx, CL2_ERR_ARG, "eps must be > 0");
    (void)minPts;
    return CL2_OK;
}

extern "C" int cl2_block_noise(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, uint8_t *keep_out,
                               int32_t *cell_first_out) {
    CL2_TRY(block_check_args(ctx, pts, eps, minPts));
    if (pts->n == 0) return CL2_OK;
    if (!keep_out) return CL2_ERR_ARG;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    CL2_TRY(cl2_grid_build(ctx, pts, eps, 0));
    cl2_grid_cache &g = pts->grid;
    Scratch sc(ctx);
    uint8_t *keep, *pk;
    int32_t *pf = nullptr;
    CL2_TRY(sc.get(&keep, (size_t)g.ncells));
    CL2_TRY(sc.get(&pk, (size_t)pts->n));
    if (cell_first_out) CL2_TRY(sc.get(&pf, (size_t)pts->n));
    {
        FamTimer t(ctx, FAM_NOISE, 37.0 * (double)g.ncells);
        k_block_noise<<<cl2_grid(g.ncells, 256), 256, 0, ctx->stream>>>(g.nbr, g.s3, keep, g.ncells, minPts);
    }
    k_keep_points<<<cl2_grid(pts->n, 256), 256, 0, ctx->stream>>>(g.cid, g.row, keep, g.first, pk, pf, pts->n);
    ctx->launches++;
    CL2_CUDA(ctx, cudaMemcpyAsync(keep_out, pk, (size_t)pts->n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += pts->n;
    if (cell_first_out) {
        CL2_CUDA(ctx, cudaMemcpyAsync(cell_first_out, pf, (size_t)pts->n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += pts->n * 4;
    }
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CL2_OK;
}

extern "C" int cl2_block_dbscan(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                                int32_t *n_clusters) {
    return cl2_block_dbscan_impl(ctx, pts, eps, minPts, labels_out, n_clusters, true);
}

// ordered = false: the clusters are reported in cell-key order with pts->ckeys[i] the key that ranks cluster i in the
// reference's numbering (no labels then)
int cl2_block_dbscan_impl(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                          int32_t *n_clusters, bool ordered, cl2_cand_sink *sink) {
    CL2_TRY(block_check_args(ctx, pts, eps, minPts));
    if (labels_out) ordered = true;
    if (n_clusters) *n_clusters = 0;
    pts->ncl = 0;
    pts->bbox.clear();
    pts->size.clear();
    pts->ckeys.clear();
    if (pts->n == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    CL2_TRY(cl2_grid_build(ctx, pts, eps, 0, minPts));
    cl2_grid_cache &g = pts->grid;
    const int64_t n = g.n;                 // rows in the grid: every point, or the rows earlier runs did not rule out
    const int32_t ncells = g.ncells;
    cudaStream_t s = ctx->stream;
    Scratch sc(ctx);

    // noise pruning + compact list of kept cells
    uint8_t *keep;
    int32_t *kpos, *list;
    int64_t *d_total;
    CL2_TRY(sc.get(&keep, (size_t)ncells));
    CL2_TRY(sc.get(&kpos, (size_t)ncells));
    CL2_TRY(sc.get(&d_total, 1));
    {
        FamTimer t(ctx, FAM_NOISE, 41.0 * (double)ncells);
        k_block_noise<<<cl2_grid(ncells, 256), 256, 0, s>>>(g.nbr, g.s3, keep, ncells, minPts, kpos);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, kpos, ncells, d_total));
    int64_t nkept64 = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &nkept64));
    const int32_t nkept = (int32_t)nkept64;
    if (pts->sweep_hint == 1 && nkept > 0) CL2_TRY(subset_update(ctx, pts, keep, eps, minPts));

    int32_t *clabel;
    CL2_TRY(sc.get(&clabel, (size_t)ncells));
    int32_t ncomp = 0;
    ClusterBox *box = nullptr;
    uint64_t *ckeys = nullptr;

    if (nkept > 0) {
        CL2_TRY(sc.get(&list, (size_t)nkept));
        {
            FamTimer t(ctx, FAM_NOISE, 9.0 * (double)ncells);
            k_keep_list<<<cl2_grid(ncells, 256), 256, 0, s>>>(keep, kpos, list, ncells);
        }
        CellStat *st;
        unsigned *adj;
        uint8_t *core;
        uint64_t *okey, *minkey;
        int32_t *parent, *root, *compid, *nheavy, *crowd = nullptr;
        int2 *heavy, *medium;
        // cell-indexed arrays are sized by ncells but only touched at kept cells
        CL2_TRY(sc.get(&st, (size_t)ncells));
        CL2_TRY(sc.get(&adj, (size_t)ncells));
        CL2_TRY(sc.get(&core, (size_t)ncells));
        CL2_TRY(sc.get(&okey, (size_t)ncells));
        CL2_TRY(sc.get(&minkey, (size_t)ncells));
        CL2_TRY(sc.get(&parent, (size_t)ncells));
        CL2_TRY(sc.get(&root, (size_t)ncells));
        CL2_TRY(sc.get(&compid, (size_t)ncells));
        CL2_TRY(sc.get(&heavy, (size_t)nkept * 4));
        CL2_TRY(sc.get(&medium, (size_t)nkept * 4));
        CL2_TRY(sc.get(&nheavy, 4));                 // [0] heavy pairs, [1] crowded cells, [2] medium pairs
        CL2_TRY(sc.get(&crowd, (size_t)nkept));
        CL2_CUDA(ctx, cudaMemsetAsync(adj, 0, (size_t)ncells * 4, s));
        CL2_CUDA(ctx, cudaMemsetAsync(core, 0, (size_t)ncells, s));
        CL2_CUDA(ctx, cudaMemsetAsync(nheavy, 0, 16, s));
        {
            FamTimer t(ctx, FAM_CELLSTAT, 8.0 * (double)n, 2);
            k_block_cellstat_small<<<cl2_grid(nkept, 256), 256, 0, s>>>(g.sxy, g.cstart, list, st, nkept, crowd, nheavy + 1);
            {
                int64_t want = cl2_grid((int64_t)nkept * 32, 256), cap = (int64_t)ctx->sm_count * 8;
                k_block_cellstat_list<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(g.sxy, g.cstart, crowd, nheavy + 1, st);
            }
        }
        {
            FamTimer t(ctx, FAM_ADJ, 8.0 * (double)n);
            k_block_adj<<<cl2_grid(nkept, 128), 128, 0, s>>>(g.sxy, g.cstart, g.nbr, keep, st, list, nkept,
                                                                           adj, eps, heavy, nheavy, medium, nheavy + 2);
        }
        {
            // medium and heavy pairs (listed by k_block_adj, counts on the device): the warps / blocks of one fixed-size
            // grid stride over the two lists
            FamTimer t(ctx, FAM_ADJ, 0.0);
            int64_t want = (int64_t)nkept * 4, cap = (int64_t)ctx->sm_count * 8;
            k_block_adj_heavy<<<(unsigned)(want < cap ? want : cap), HEAVY_THREADS, 0, s>>>(g.sxy, g.cstart, g.nbr, heavy, nheavy,
                                                                                           medium, nheavy + 2, adj, eps);
        }
        {
            FamTimer t(ctx, FAM_CORE, 64.0 * (double)nkept);
            k_block_core<<<cl2_grid(nkept, 256), 256, 0, s>>>(g.cstart, g.nbr, adj, g.first, list, nkept, minPts, core, okey,
                                                               parent, minkey);
        }
        {
            FamTimer t(ctx, FAM_UNION, 48.0 * (double)nkept, 2);
            k_block_union<<<cl2_grid(nkept, 256), 256, 0, s>>>(g.nbr, adj, core, list, nkept, parent);
            k_block_roots<<<cl2_grid(nkept, 256), 256, 0, s>>>(core, list, nkept, parent, okey, minkey, root);
        }
        // component list -> sort by start-cell order key -> cluster ids
        int32_t *rflag;
        CL2_TRY(sc.get(&rflag, (size_t)nkept));
        {
            FamTimer t(ctx, FAM_COMP, 16.0 * (double)nkept);
            k_block_rootflags<<<cl2_grid(nkept, 256), 256, 0, s>>>(core, root, list, nkept, rflag);
        }
        CL2_TRY(cl2_exclusive_scan_i32(ctx, rflag, nkept, d_total));
        int64_t ncomp64 = 0;
        CL2_TRY(cl2_read_i64(ctx, d_total, &ncomp64));
        ncomp = (int32_t)ncomp64;
        if (ncomp > 0) {
            CL2_TRY(sc.get(&ckeys, (size_t)ncomp));
            FamTimer t(ctx, FAM_COMP, 24.0 * (double)nkept);
            k_block_rootlist<<<cl2_grid(nkept, 256), 256, 0, s>>>(core, root, list, nkept, rflag, minkey, ckeys, compid);
        }
        {
            FamTimer t(ctx, FAM_BORDER, 64.0 * (double)nkept);
            k_fill_i32<<<cl2_grid(ncells, 256), 256, 0, s>>>(clabel, -1, ncells);
            k_block_celllabel<<<cl2_grid(nkept, 256), 256, 0, s>>>(g.nbr, adj, core, root, compid, okey, minkey, list,
                                                                   nkept, clabel);
            ctx->launches++;
        }
        if (ncomp > 0) {
            CL2_TRY(sc.get(&box, (size_t)ncomp));
            FamTimer t(ctx, FAM_BBOX, 40.0 * (double)nkept, 2);
            k_box_init<<<cl2_grid(ncomp, 256), 256, 0, s>>>(box, ncomp);
            k_block_bbox<<<cl2_grid(nkept, 256), 256, 0, s>>>(g.cstart, st, clabel, list, nkept, box);
        }
    } else {
        k_fill_i32<<<cl2_grid(ncells, 256), 256, 0, s>>>(clabel, -1, ncells);
        ctx->launches++;
    }

    if (sink && !labels_out) {
        // the boxes go straight into the device-resident candidate list of cl2_call_loops (filters applied there)
        static_assert(sizeof(ClusterBox) == 5 * sizeof(int), "box record");
        if (ncomp > 0) CL2_TRY(cl2_sink_emit(ctx, sink, reinterpret_cast<const int *>(box), ckeys, ncomp));
        CL2_CUDA(ctx, cudaGetLastError());
        pts->ncl = ncomp;
        if (n_clusters) *n_clusters = ncomp;
        return CL2_OK;
    }
    // Cluster numbering = rank of the components' start-cell order keys (blockDBSCAN.py:146-152).  The component
    // count is small next to n, so the keys and boxes go to the host and are ordered there.  Callers that only need the
    // boxes and sort a few survivors themselves (cl2_call_loops) take them unordered, each with its key.
    std::vector<ClusterBox> hb((size_t)ncomp);
    std::vector<uint64_t> hk((size_t)ncomp);
    if (ncomp > 0) {
        CL2_CUDA(ctx, cudaMemcpyAsync(hb.data(), box, (size_t)ncomp * sizeof(ClusterBox), cudaMemcpyDeviceToHost, s));
        CL2_CUDA(ctx, cudaMemcpyAsync(hk.data(), ckeys, (size_t)ncomp * 8, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += (int64_t)ncomp * ((int64_t)sizeof(ClusterBox) + 8);
    }
    std::vector<int32_t> order((size_t)ncomp);
    CL2_CUDA(ctx, cudaStreamSynchronize(s));
    CL2_CUDA(ctx, cudaGetLastError());
    for (int32_t i = 0; i < ncomp; i++) order[i] = i;
    if (ordered) std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hk[a] < hk[b]; });
    pts->ckeys.resize((size_t)ncomp);
    for (int32_t i = 0; i < ncomp; i++) pts->ckeys[i] = hk[order[i]];
    pts->ncl = ncomp;
    pts->bbox.resize((size_t)ncomp * 4);
    pts->size.resize((size_t)ncomp);
    for (int32_t i = 0; i < ncomp; i++) {
        const ClusterBox &b = hb[order[i]];
        pts->bbox[4 * (size_t)i] = b.xmin;
        pts->bbox[4 * (size_t)i + 1] = b.xmax;
        pts->bbox[4 * (size_t)i + 2] = b.ymin;
        pts->bbox[4 * (size_t)i + 3] = b.ymax;
        pts->size[i] = b.size;
    }
    if (labels_out) {
        int32_t *dl, *dmap;
        const int64_t nall = pts->n;
        CL2_TRY(sc.get(&dl, (size_t)nall));
        CL2_TRY(sc.get(&dmap, (size_t)ncomp));
        if (n < nall) CL2_CUDA(ctx, cudaMemsetAsync(dl, 0xff, (size_t)nall * 4, s));   // rows not in the grid are noise
        std::vector<int32_t> final_of_prov((size_t)ncomp);
        for (int32_t i = 0; i < ncomp; i++) final_of_prov[order[i]] = i;
        if (ncomp > 0) CL2_CUDA(ctx, cudaMemcpyAsync(dmap, final_of_prov.data(), (size_t)ncomp * 4, cudaMemcpyHostToDevice, s));
        {
            FamTimer t(ctx, FAM_LABEL, 12.0 * (double)n);
            k_point_labels<<<cl2_grid(n, 256), 256, 0, s>>>(g.cid, g.row, clabel, dmap, dl, n);
        }
        CL2_CUDA(ctx, cudaMemcpyAsync(labels_out, dl, (size_t)nall * 4, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += nall * 4;
        ctx->h2d_bytes += (int64_t)ncomp * 4;
        CL2_CUDA(ctx, cudaStreamSynchronize(s));
        CL2_CUDA(ctx, cudaGetLastError());
    }
    if (n_clusters) *n_clusters = ncomp;
    return CL2_OK;
}

extern "C" int cl2_cluster_bbox(cl2_ctx *ctx, cl2_points *pts, int64_t *bbox_out, int64_t *size_out) {
    if (!ctx || !pts) return CL2_ERR_ARG;
    if (pts->ncl > 0) {
        if (bbox_out) memcpy(bbox_out, pts->bbox.data(), (size_t)pts->ncl * 4 * sizeof(int64_t));
        if (size_out) memcpy(size_out, pts->size.data(), (size_t)pts->ncl * sizeof(int64_t));
    }
    return CL2_OK;
}
