// cl2_cdbscan.cu -- cDBSCAN2 (cLoops2/cDBSCAN2.py:13-383, used with -hic) as a data-parallel pipeline.
//
// Restatement (re-verified against the reference through oracle/):
//   grid       45-degree rotation u = X-Y, v = X+Y, cell = truncating division by eps (cDBSCAN2.py:67-70)
//   neighbours p ~ q inside the 3x3 cell block: same cell always; edge-adjacent cells compare ONLY the axis that
//              differs, corner cells both axes, one-sided as binSearchAdjPt does (:364-378)
//   core point deg(p) = #{q : p ~ q} >= minPts; every point of a cell with cnt >= minPts is core (:80-83)
//   clusters   components of core points.  All core points of one cell are mutual neighbours, so the union-find
//              runs over CELLS that hold core points; two such cells are joined iff some core-core pair is ~
//   order      components are visited in the dict order of their first cell (= smallest row id of that cell)
//   border     a non-core point goes to the first visited, non-released component that has a core neighbour of it
//   release    a component whose core points + claimed border points < minPts is released (:180-183); resolved
//              by a monotone lower/upper-bound iteration over the (few) components with < minPts core points
#include "cl2_points.cuh"
#include <algorithm>

#define ST_UNKNOWN 0
#define ST_ALIVE 1
#define ST_DEAD 2

__device__ __forceinline__ int dir_dx(int d) { return (d == 2 || d == 4 || d == 5) ? -1 : ((d == 3 || d == 6 || d == 7) ? 1 : 0); }
__device__ __forceinline__ int dir_dy(int d) { return (d == 0 || d == 4 || d == 6) ? -1 : ((d == 1 || d == 5 || d == 7) ? 1 : 0); }
__device__ __forceinline__ int dir_opp(int d) { return d ^ 1 ^ ((d >= 4) ? 2 : 0); }   // 0<->1, 2<->3, 4<->7, 5<->6

struct UV {
    long long u, v;
};
__device__ __forceinline__ UV rot(int2 p) {
    UV r;
    r.u = (long long)p.x - (long long)p.y;
    r.v = (long long)p.x + (long long)p.y;
    return r;
}
// q lies in the cell at offset (dx,dy) of p's cell
__device__ __forceinline__ bool near_off(int dx, int dy, const UV &p, const UV &q, long long eps) {
    if (dx == 1 && !(q.u <= p.u + eps)) return false;
    if (dx == -1 && !(q.u >= p.u - eps)) return false;
    if (dy == 1 && !(q.v <= p.v + eps)) return false;
    if (dy == -1 && !(q.v >= p.v - eps)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------ core points
__global__ void __launch_bounds__(256) k_c2_core(const int2 *__restrict__ sxy, const int32_t *__restrict__ cid,
                                                  const int32_t *__restrict__ cstart, const int32_t *__restrict__ nbr,
                                                  const int32_t *__restrict__ s3, uint8_t *__restrict__ corep, int64_t n,
                                                  int minPts, long long eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cid[i];
    int cnt = cstart[c + 1] - cstart[c];
    bool core;
    if (cnt >= minPts) core = true;
    else if (s3[c] < minPts) core = false;
    else {
        UV p = rot(sxy[i]);
        int deg = cnt;
        for (int d = 0; d < 8 && deg < minPts; d++) {
            int q = nbr[8 * (int64_t)c + d];
            if (q < 0) continue;
            int dx = dir_dx(d), dy = dir_dy(d);
            int e = cstart[q + 1];
            for (int j = cstart[q]; j < e && deg < minPts; j++)
                if (near_off(dx, dy, p, rot(sxy[j]), eps)) deg++;
        }
        core = deg >= minPts;
    }
    corep[i] = core ? 1 : 0;
}

// per cell: number of core points and their extent in (u,v); one warp per cell
struct CoreStat {
    long long umin, umax, vmin, vmax;
};
__global__ void __launch_bounds__(256) k_c2_cellinit(const int32_t *__restrict__ s3, int32_t *__restrict__ ncore,
                                                      int32_t *__restrict__ parent, unsigned *__restrict__ minkey,
                                                      int32_t *__restrict__ aflag, int32_t ncells, int minPts) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    ncore[c] = 0;
    parent[c] = c;
    minkey[c] = 0xffffffffu;
    aflag[c] = s3[c] >= minPts ? 1 : 0;   // only these cells can hold core points (cDBSCAN2.py:86-88)
}
__global__ void __launch_bounds__(256) k_c2_activelist(const int32_t *__restrict__ s3, const int32_t *__restrict__ pos,
                                                        int32_t *__restrict__ list, int32_t ncells, int minPts) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && s3[c] >= minPts) list[pos[c]] = c;
}
__global__ void __launch_bounds__(256) k_c2_cellstat(const int2 *__restrict__ sxy, const int32_t *__restrict__ cstart,
                                                      const uint8_t *__restrict__ corep, const int32_t *__restrict__ list,
                                                      int32_t nactive, int32_t *__restrict__ ncore,
                                                      CoreStat *__restrict__ cs) {
    int k = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (k >= nactive) return;
    int c = list[k];
    int nc = 0;
    long long umin = LLONG_MAX, umax = LLONG_MIN, vmin = LLONG_MAX, vmax = LLONG_MIN;
    for (int i = cstart[c] + lane; i < cstart[c + 1]; i += 32) {
        if (!corep[i]) continue;
        UV p = rot(sxy[i]);
        nc++;
        umin = min(umin, p.u); umax = max(umax, p.u);
        vmin = min(vmin, p.v); vmax = max(vmax, p.v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        nc += __shfl_xor_sync(0xffffffffu, nc, o);
        umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o));
        umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if (lane == 0) {
        ncore[c] = nc;
        CoreStat s;
        s.umin = umin; s.umax = umax; s.vmin = vmin; s.vmax = vmax;
        cs[c] = s;
    }
}

// ------------------------------------------------------------------------------------------------ cell edges
#define C2_LIGHT_MAX 4096

// forward directions (1: dy=+1, 6: (+1,-1), 3: dx=+1, 7: (+1,+1)); edge existence between cells holding core points
__global__ void __launch_bounds__(128) k_c2_edges(const int2 *__restrict__ sxy, const int32_t *__restrict__ cstart,
                                                   const int32_t *__restrict__ nbr, const uint8_t *__restrict__ corep,
                                                   const int32_t *__restrict__ ncore, const CoreStat *__restrict__ cs,
                                                   const int32_t *__restrict__ list, int32_t nactive, long long eps,
                                                   unsigned *__restrict__ adj, int2 *__restrict__ heavy,
                                                   int32_t *__restrict__ nheavy) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((t >> 2) >= nactive) return;
    int a = list[t >> 2];
    if (ncore[a] == 0) return;
    const int fdir[4] = {1, 6, 3, 7};
    int d = fdir[t & 3];
    int b = nbr[8 * (int64_t)a + d];
    if (b < 0 || ncore[b] == 0) return;
    const CoreStat A = cs[a], B = cs[b];
    bool hit = false;
    if (d == 1) hit = B.vmin <= A.vmax + eps;
    else if (d == 3) hit = B.umin <= A.umax + eps;
    else {
        bool possible = B.umin <= A.umax + eps && (d == 7 ? B.vmin <= A.vmax + eps : B.vmax >= A.vmin - eps);
        if (!possible) return;
        int a0 = cstart[a], a1 = cstart[a + 1], b0 = cstart[b], b1 = cstart[b + 1];
        if ((long long)(a1 - a0) * (b1 - b0) > C2_LIGHT_MAX) {
            int slot = atomicAdd(nheavy, 1);
            heavy[slot] = make_int2(a, d);
            return;
        }
        int dy = d == 7 ? 1 : -1;
        for (int i = a0; i < a1 && !hit; i++) {
            if (!corep[i]) continue;
            UV p = rot(sxy[i]);
            for (int j = b0; j < b1; j++) {
                if (!corep[j]) continue;
                if (near_off(1, dy, p, rot(sxy[j]), eps)) { hit = true; break; }
            }
        }
    }
    if (hit) {
        atomicOr(&adj[a], 1u << d);
        atomicOr(&adj[b], 1u << dir_opp(d));
    }
}

#define C2H_THREADS 256
#define C2H_TILE 1024
__global__ void __launch_bounds__(C2H_THREADS) k_c2_edges_heavy(const int2 *__restrict__ sxy,
                                                                 const int32_t *__restrict__ cstart,
                                                                 const int32_t *__restrict__ nbr,
                                                                 const uint8_t *__restrict__ corep,
                                                                 const int2 *__restrict__ heavy, long long eps,
                                                                 unsigned *__restrict__ adj) {
    __shared__ long long tu[C2H_TILE], tv[C2H_TILE];
    __shared__ int tn_s, found;
    int2 h = heavy[blockIdx.x];
    int a = h.x, d = h.y;
    int b = nbr[8 * (int64_t)a + d];
    int dy = d == 7 ? 1 : -1;
    int a0 = cstart[a], a1 = cstart[a + 1], b0 = cstart[b], b1 = cstart[b + 1];
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    for (int tb = b0; tb < b1; tb += C2H_TILE) {
        if (threadIdx.x == 0) tn_s = 0;
        __syncthreads();
        int lim = min(C2H_TILE, b1 - tb);
        for (int j = threadIdx.x; j < lim; j += C2H_THREADS) {
            if (corep[tb + j]) {
                UV q = rot(sxy[tb + j]);
                int s = atomicAdd(&tn_s, 1);
                tu[s] = q.u; tv[s] = q.v;
            }
        }
        __syncthreads();
        int tn = tn_s;
        for (int i = a0 + threadIdx.x; i < a1; i += C2H_THREADS) {
            if (!corep[i]) continue;
            UV p = rot(sxy[i]);
            bool hit = false;
            for (int j = 0; j < tn; j++) {
                UV q; q.u = tu[j]; q.v = tv[j];
                if (near_off(1, dy, p, q, eps)) { hit = true; break; }
            }
            if (hit) { found = 1; break; }
            if (found) break;
        }
        __syncthreads();
        int f = found;
        if (f) break;
    }
    if (threadIdx.x == 0 && found) {
        atomicOr(&adj[a], 1u << d);
        atomicOr(&adj[b], 1u << dir_opp(d));
    }
}

// ------------------------------------------------------------------------------------------------ union-find on cells
__device__ __forceinline__ int c2_find(int32_t *parent, int x) {
    int p = parent[x];
    while (p != x) {
        int gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void c2_unite(int32_t *parent, int a, int b) {
    while (true) {
        a = c2_find(parent, a);
        b = c2_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        if (atomicCAS(&parent[b], b, a) == b) return;
    }
}
__global__ void __launch_bounds__(256) k_c2_union(const int32_t *__restrict__ nbr, const unsigned *__restrict__ adj,
                                                   const int32_t *__restrict__ ncore,
                                                   const int32_t *__restrict__ list, int32_t nactive,
                                                   int32_t *__restrict__ parent) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nactive) return;
    int c = list[k];
    if (ncore[c] == 0) return;
    unsigned m = adj[c] & ((1u << 1) | (1u << 6) | (1u << 3) | (1u << 7));
    while (m) {
        int d = __ffs(m) - 1;
        m &= m - 1;
        c2_unite(parent, c, nbr[8 * (int64_t)c + d]);
    }
}
// root, component order key (smallest first-row over its cells, cDBSCAN2.py:117-140) and core-point total
__global__ void __launch_bounds__(256) k_c2_roots(const int32_t *__restrict__ ncore, const int32_t *__restrict__ cstart,
                                                   const uint32_t *__restrict__ first, int32_t ncells,
                                                   int32_t *__restrict__ parent, int32_t *__restrict__ root,
                                                   unsigned *__restrict__ minkey, int32_t *__restrict__ compcore) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    if (ncore[c] == 0) { root[c] = -1; return; }
    int r = c2_find(parent, c);
    root[c] = r;
    atomicMin(&minkey[r], first[c]);
    atomicAdd(&compcore[r], ncore[c]);
}
__global__ void __launch_bounds__(256) k_c2_status_init(const int32_t *__restrict__ root,
                                                         const int32_t *__restrict__ compcore, int32_t ncells,
                                                         int minPts, uint8_t *__restrict__ status,
                                                         int32_t *__restrict__ nunknown) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    uint8_t s = ST_DEAD;
    if (root[c] == c) {
        s = compcore[c] >= minPts ? ST_ALIVE : ST_UNKNOWN;
        if (s == ST_UNKNOWN) atomicAdd(nunknown, 1);
    }
    status[c] = s;
}

// ------------------------------------------------------------------------------------------------ border points
// bit t (0..7) : the neighbour cell in direction t holds a core neighbour of this point; bit 8: own cell holds core
__global__ void __launch_bounds__(256) k_c2_border_mask(const int2 *__restrict__ sxy, const int32_t *__restrict__ cid,
                                                         const int32_t *__restrict__ cstart,
                                                         const int32_t *__restrict__ nbr,
                                                         const uint8_t *__restrict__ corep,
                                                         const int32_t *__restrict__ ncore,
                                                         uint16_t *__restrict__ bmask, int64_t n, long long eps) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned m = 0;
    if (!corep[i]) {
        int c = cid[i];
        if (ncore[c] > 0) m |= 1u << 8;
        UV p = rot(sxy[i]);
        for (int d = 0; d < 8; d++) {
            int q = nbr[8 * (int64_t)c + d];
            if (q < 0 || ncore[q] == 0) continue;
            int dx = dir_dx(d), dy = dir_dy(d);
            int e = cstart[q + 1];
            for (int j = cstart[q]; j < e; j++) {
                if (corep[j] && near_off(dx, dy, p, rot(sxy[j]), eps)) { m |= 1u << d; break; }
            }
        }
    }
    bmask[i] = (uint16_t)m;
}

// candidate components of a border point, ordered by visiting order; returns count
__device__ __forceinline__ int c2_candidates(unsigned m, int c, const int32_t *__restrict__ nbr,
                                             const int32_t *__restrict__ root, const unsigned *__restrict__ minkey,
                                             int *rr, unsigned *kk) {
    int k = 0;
    while (m) {
        int t = __ffs(m) - 1;
        m &= m - 1;
        int cell = t == 8 ? c : nbr[8 * (int64_t)c + t];
        int r = root[cell];
        bool dup = false;
        for (int z = 0; z < k; z++) dup |= rr[z] == r;
        if (dup) continue;
        unsigned key = minkey[r];
        int z = k++;
        while (z > 0 && kk[z - 1] > key) { rr[z] = rr[z - 1]; kk[z] = kk[z - 1]; z--; }
        rr[z] = r; kk[z] = key;
    }
    return k;
}

// one round of the release iteration: bounds on the border points each undecided component would claim
__global__ void __launch_bounds__(256) k_c2_round(const int32_t *__restrict__ cid, const int32_t *__restrict__ nbr,
                                                   const uint16_t *__restrict__ bmask, const int32_t *__restrict__ root,
                                                   const unsigned *__restrict__ minkey, const uint8_t *__restrict__ status,
                                                   int32_t *__restrict__ lb, int32_t *__restrict__ ub, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned m = bmask[i];
    if (!m) return;
    int rr[9];
    unsigned kk[9];
    int k = c2_candidates(m, cid[i], nbr, root, minkey, rr, kk);
    bool all_dead = true, none_alive = true;
    for (int z = 0; z < k && none_alive; z++) {
        uint8_t s = status[rr[z]];
        if (s == ST_UNKNOWN) {
            if (all_dead) atomicAdd(&lb[rr[z]], 1);
            atomicAdd(&ub[rr[z]], 1);
            all_dead = false;
        } else if (s == ST_ALIVE) {
            none_alive = false;
        }
    }
}
__global__ void __launch_bounds__(256) k_c2_decide(const int32_t *__restrict__ root, const int32_t *__restrict__ compcore,
                                                    int32_t ncells, int minPts, uint8_t *__restrict__ status,
                                                    int32_t *__restrict__ lb, int32_t *__restrict__ ub,
                                                    int32_t *__restrict__ nunknown) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells || root[c] != c || status[c] != ST_UNKNOWN) return;
    int lo = compcore[c] + lb[c], hi = compcore[c] + ub[c];
    if (lo >= minPts) status[c] = ST_ALIVE;
    else if (hi < minPts) status[c] = ST_DEAD;
    else atomicAdd(nunknown, 1);
    lb[c] = 0;
    ub[c] = 0;
}

// ------------------------------------------------------------------------------------------------ numbering + labels
__global__ void __launch_bounds__(256) k_c2_aliveflags(const int32_t *__restrict__ root, const uint8_t *__restrict__ status,
                                                        int32_t ncells, int32_t *__restrict__ flag) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) flag[c] = (root[c] == c && status[c] == ST_ALIVE) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_c2_alivelist(const int32_t *__restrict__ root, const uint8_t *__restrict__ status,
                                                       const int32_t *__restrict__ pos, const unsigned *__restrict__ minkey,
                                                       int32_t ncells, unsigned *__restrict__ keys,
                                                       int32_t *__restrict__ compid) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && root[c] == c && status[c] == ST_ALIVE) {
        keys[pos[c]] = minkey[c];
        compid[c] = pos[c];            // provisional id; final numbering = host sort of the keys
    }
}
__global__ void __launch_bounds__(256) k_c2_widen(const unsigned *__restrict__ in, uint64_t *__restrict__ out, int32_t n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_c2_remap(const int32_t *__restrict__ root, const uint8_t *__restrict__ status,
                                                   const int32_t *__restrict__ final_of_prov, int32_t ncells,
                                                   int32_t *__restrict__ compid) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && root[c] == c && status[c] == ST_ALIVE) compid[c] = final_of_prov[compid[c]];
}

struct C2Box {
    int xmin, xmax, ymin, ymax;
    unsigned size;
};
__global__ void __launch_bounds__(256) k_c2_box_init(C2Box *__restrict__ box, int32_t ncl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncl) {
        C2Box b;
        b.xmin = INT_MAX; b.xmax = INT_MIN; b.ymin = INT_MAX; b.ymax = INT_MIN; b.size = 0;
        box[i] = b;
    }
}

// label of every point (sorted order) -> optional scatter to file order, and cluster boxes with warp aggregation
__global__ void __launch_bounds__(256) k_c2_labels(const int2 *__restrict__ sxy, const int32_t *__restrict__ cid,
                                                    const uint32_t *__restrict__ row, const int32_t *__restrict__ nbr,
                                                    const uint8_t *__restrict__ corep, const uint16_t *__restrict__ bmask,
                                                    const int32_t *__restrict__ root, const unsigned *__restrict__ minkey,
                                                    const uint8_t *__restrict__ status, const int32_t *__restrict__ compid,
                                                    int32_t *__restrict__ labels, C2Box *__restrict__ box, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lab = -1;
    int2 p = make_int2(0, 0);
    if (i < n) {
        int c = cid[i];
        p = sxy[i];
        if (corep[i]) {
            int r = root[c];
            if (status[r] == ST_ALIVE) lab = compid[r];
        } else {
            unsigned m = bmask[i];
            if (m) {
                int rr[9];
                unsigned kk[9];
                int k = c2_candidates(m, c, nbr, root, minkey, rr, kk);
                for (int z = 0; z < k; z++)
                    if (status[rr[z]] == ST_ALIVE) { lab = compid[rr[z]]; break; }
            }
        }
        if (labels) labels[row[i]] = lab;
    }
    // neighbouring threads are neighbouring points of the same cell: aggregate per label inside the warp
    unsigned peers = __match_any_sync(0xffffffffu, lab);
    if (lab >= 0) {
        int xmin = __reduce_min_sync(peers, p.x), xmax = __reduce_max_sync(peers, p.x);
        int ymin = __reduce_min_sync(peers, p.y), ymax = __reduce_max_sync(peers, p.y);
        if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) {
            C2Box *b = &box[lab];
            atomicMin(&b->xmin, xmin);
            atomicMax(&b->xmax, xmax);
            atomicMin(&b->ymin, ymin);
            atomicMax(&b->ymax, ymax);
            atomicAdd(&b->size, (unsigned)__popc(peers));
        }
    }
}

// ================================================================================================ host driver
extern "C" int cl2_cdbscan2(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                            int32_t *n_clusters) {
    return cl2_cdbscan2_impl(ctx, pts, eps, minPts, labels_out, n_clusters, nullptr);
}

int cl2_cdbscan2_impl(cl2_ctx *ctx, cl2_points *pts, int64_t eps, int32_t minPts, int32_t *labels_out,
                      int32_t *n_clusters, cl2_cand_sink *sink) {
    if (!ctx || !pts) return CL2_ERR_ARG;
    if (eps <= 0) return cl2_fail(ctx, CL2_ERR_ARG, "eps must be > 0");
    if (n_clusters) *n_clusters = 0;
    pts->ncl = 0;
    pts->bbox.clear();
    pts->size.clear();
    const int64_t n = pts->n;
    if (n == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    CL2_TRY(cl2_grid_build(ctx, pts, eps, 1));
    cl2_grid_cache &g = pts->grid;
    const int32_t ncells = g.ncells;
    cudaStream_t s = ctx->stream;
    Scratch sc(ctx);

    uint8_t *corep, *status;
    uint16_t *bmask;
    int32_t *ncore, *parent, *root, *compcore, *lb, *ub, *compid, *nheavy, *nunknown, *aflag;
    unsigned *minkey, *adj;
    CoreStat *cs;
    int2 *heavy;
    int64_t *d_total;
    CL2_TRY(sc.get(&corep, (size_t)n));
    CL2_TRY(sc.get(&bmask, (size_t)n));
    CL2_TRY(sc.get(&ncore, (size_t)ncells));
    CL2_TRY(sc.get(&cs, (size_t)ncells));
    CL2_TRY(sc.get(&parent, (size_t)ncells));
    CL2_TRY(sc.get(&root, (size_t)ncells));
    CL2_TRY(sc.get(&minkey, (size_t)ncells));
    CL2_TRY(sc.get(&adj, (size_t)ncells));
    CL2_TRY(sc.get(&compcore, (size_t)ncells));
    CL2_TRY(sc.get(&lb, (size_t)ncells));
    CL2_TRY(sc.get(&ub, (size_t)ncells));
    CL2_TRY(sc.get(&compid, (size_t)ncells));
    CL2_TRY(sc.get(&status, (size_t)ncells));
    CL2_TRY(sc.get(&aflag, (size_t)ncells));
    CL2_TRY(sc.get(&heavy, (size_t)ncells * 2));   // at most two corner directions per cell
    CL2_TRY(sc.get(&nheavy, 1));
    CL2_TRY(sc.get(&nunknown, 1));
    CL2_TRY(sc.get(&d_total, 1));
    CL2_CUDA(ctx, cudaMemsetAsync(adj, 0, (size_t)ncells * 4, s));
    CL2_CUDA(ctx, cudaMemsetAsync(compcore, 0, (size_t)ncells * 4, s));
    CL2_CUDA(ctx, cudaMemsetAsync(lb, 0, (size_t)ncells * 4, s));
    CL2_CUDA(ctx, cudaMemsetAsync(ub, 0, (size_t)ncells * 4, s));
    CL2_CUDA(ctx, cudaMemsetAsync(nheavy, 0, 4, s));
    CL2_CUDA(ctx, cudaMemsetAsync(nunknown, 0, 4, s));
    {
        FamTimer t(ctx, FAM_CORE, 13.0 * (double)n);
        k_c2_core<<<cl2_grid(n, 256), 256, 0, s>>>(g.sxy, g.cid, g.cstart, g.nbr, g.s3, corep, n, minPts, (long long)eps);
    }
    {
        FamTimer t(ctx, FAM_CELLSTAT, 20.0 * (double)ncells);
        k_c2_cellinit<<<cl2_grid(ncells, 256), 256, 0, s>>>(g.s3, ncore, parent, minkey, aflag, ncells, minPts);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, aflag, ncells, d_total));
    int64_t nactive64 = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &nactive64));
    const int32_t nactive = (int32_t)nactive64;
    int32_t *alist = nullptr;
    if (nactive > 0) {
        CL2_TRY(sc.get(&alist, (size_t)nactive));
        FamTimer t(ctx, FAM_CELLSTAT, 9.0 * (double)n, 2);
        k_c2_activelist<<<cl2_grid(ncells, 256), 256, 0, s>>>(g.s3, aflag, alist, ncells, minPts);
        k_c2_cellstat<<<cl2_grid((int64_t)nactive * 32, 256), 256, 0, s>>>(g.sxy, g.cstart, corep, alist, nactive, ncore, cs);
    }
    if (nactive > 0) {
        FamTimer t(ctx, FAM_ADJ, 9.0 * (double)n);
        k_c2_edges<<<cl2_grid((int64_t)nactive * 4, 128), 128, 0, s>>>(g.sxy, g.cstart, g.nbr, corep, ncore, cs, alist,
                                                                       nactive, (long long)eps, adj, heavy, nheavy);
    }
    int32_t nh = 0;
    CL2_TRY(cl2_read_i32(ctx, nheavy, &nh));
    if (nh > 0) {
        FamTimer t(ctx, FAM_ADJ, 0.0);
        k_c2_edges_heavy<<<nh, C2H_THREADS, 0, s>>>(g.sxy, g.cstart, g.nbr, corep, heavy, (long long)eps, adj);
    }
    {
        FamTimer t(ctx, FAM_UNION, 56.0 * (double)ncells, 3);
        if (nactive > 0) k_c2_union<<<cl2_grid(nactive, 256), 256, 0, s>>>(g.nbr, adj, ncore, alist, nactive, parent);
        k_c2_roots<<<cl2_grid(ncells, 256), 256, 0, s>>>(ncore, g.cstart, g.first, ncells, parent, root, minkey, compcore);
        k_c2_status_init<<<cl2_grid(ncells, 256), 256, 0, s>>>(root, compcore, ncells, minPts, status, nunknown);
    }
    {
        FamTimer t(ctx, FAM_BORDER, 15.0 * (double)n);
        k_c2_border_mask<<<cl2_grid(n, 256), 256, 0, s>>>(g.sxy, g.cid, g.cstart, g.nbr, corep, ncore, bmask, n,
                                                          (long long)eps);
    }
    // release iteration (cDBSCAN2.py:180-183): at least the earliest undecided component is decided per round
    int32_t nunk = 0;
    CL2_TRY(cl2_read_i32(ctx, nunknown, &nunk));
    int rounds = 0;
    while (nunk > 0) {
        CL2_CUDA(ctx, cudaMemsetAsync(nunknown, 0, 4, s));
        {
            FamTimer t(ctx, FAM_BORDER, 6.0 * (double)n, 2);
            k_c2_round<<<cl2_grid(n, 256), 256, 0, s>>>(g.cid, g.nbr, bmask, root, minkey, status, lb, ub, n);
            k_c2_decide<<<cl2_grid(ncells, 256), 256, 0, s>>>(root, compcore, ncells, minPts, status, lb, ub, nunknown);
        }
        int32_t prev = nunk;
        CL2_TRY(cl2_read_i32(ctx, nunknown, &nunk));
        if (++rounds > 1000000 || nunk > prev) return cl2_fail(ctx, CL2_ERR_CUDA, "cl2_cdbscan2: release iteration did not converge");
    }
    // number the surviving components in visiting order
    {
        FamTimer t(ctx, FAM_COMP, 9.0 * (double)ncells);
        k_c2_aliveflags<<<cl2_grid(ncells, 256), 256, 0, s>>>(root, status, ncells, aflag);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, aflag, ncells, d_total));
    int64_t ncomp64 = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &ncomp64));
    const int32_t ncomp = (int32_t)ncomp64;
    C2Box *box = nullptr;
    if (ncomp > 0) {
        unsigned *dkeys;
        int32_t *dmap;
        CL2_TRY(sc.get(&dkeys, (size_t)ncomp));
        CL2_TRY(sc.get(&dmap, (size_t)ncomp));
        CL2_TRY(sc.get(&box, (size_t)ncomp));
        {
            FamTimer t(ctx, FAM_COMP, 21.0 * (double)ncells);
            k_c2_alivelist<<<cl2_grid(ncells, 256), 256, 0, s>>>(root, status, aflag, minkey, ncells, dkeys, compid);
        }
        // visiting order of the components = ascending first-row key (cDBSCAN2.py:117-140); few components, host sort
        std::vector<unsigned> hk((size_t)ncomp);
        std::vector<int32_t> order((size_t)ncomp), fin((size_t)ncomp);
        if (ncomp > 4096) {
            // many components: device radix sort of (key, position), read back the permutation
            uint64_t *k0, *k1, *cks;
            uint32_t *o0, *o1, *os;
            CL2_TRY(sc.get(&k0, (size_t)ncomp));
            CL2_TRY(sc.get(&k1, (size_t)ncomp));
            CL2_TRY(sc.get(&o0, (size_t)ncomp));
            CL2_TRY(sc.get(&o1, (size_t)ncomp));
            k_c2_widen<<<cl2_grid(ncomp, 256), 256, 0, s>>>(dkeys, k0, ncomp);
            ctx->launches++;
            ctx->sort_family = FAM_COMP;   // a few thousand keys: part of the component ordering, not of the PET sorts
            const int src = cl2_radix_sort_pairs(ctx, k0, k1, o0, o1, ncomp, bits_for((uint64_t)n), &cks, &os);
            ctx->sort_family = -1;
            CL2_TRY(src);
            (void)cks;
            CL2_CUDA(ctx, cudaMemcpyAsync(order.data(), os, (size_t)ncomp * 4, cudaMemcpyDeviceToHost, s));
            CL2_CUDA(ctx, cudaStreamSynchronize(s));
        } else {
            CL2_CUDA(ctx, cudaMemcpyAsync(hk.data(), dkeys, (size_t)ncomp * 4, cudaMemcpyDeviceToHost, s));
            CL2_CUDA(ctx, cudaStreamSynchronize(s));
            for (int32_t i = 0; i < ncomp; i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hk[a] < hk[b]; });
        }
        for (int32_t i = 0; i < ncomp; i++) fin[order[i]] = i;
        CL2_CUDA(ctx, cudaMemcpyAsync(dmap, fin.data(), (size_t)ncomp * 4, cudaMemcpyHostToDevice, s));
        ctx->d2h_bytes += (int64_t)ncomp * 4;
        ctx->h2d_bytes += (int64_t)ncomp * 4;
        {
            FamTimer t(ctx, FAM_COMP, 12.0 * (double)ncells, 2);
            k_c2_remap<<<cl2_grid(ncells, 256), 256, 0, s>>>(root, status, dmap, ncells, compid);
            k_c2_box_init<<<cl2_grid(ncomp, 256), 256, 0, s>>>(box, ncomp);
        }
        CL2_CUDA(ctx, cudaStreamSynchronize(s));   // fin[] must outlive the copy
    }
    int32_t *dl = nullptr;
    if (labels_out) CL2_TRY(sc.get(&dl, (size_t)n));
    if (ncomp > 0 || labels_out) {
        FamTimer t(ctx, FAM_LABEL, (labels_out ? 23.0 : 15.0) * (double)n);
        k_c2_labels<<<cl2_grid(n, 256), 256, 0, s>>>(g.sxy, g.cid, g.row, g.nbr, corep, bmask, root, minkey, status, compid,
                                                     dl, box, n);
    }
    if (labels_out) {
        CL2_CUDA(ctx, cudaMemcpyAsync(labels_out, dl, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += n * 4;
    }
    if (sink && !labels_out) {
        // clusters are in their final order here: the record index is the rank key
        static_assert(sizeof(C2Box) == 5 * sizeof(int), "box record");
        if (ncomp > 0) CL2_TRY(cl2_sink_emit(ctx, sink, reinterpret_cast<const int *>(box), nullptr, ncomp));
        CL2_CUDA(ctx, cudaStreamSynchronize(s));
        CL2_CUDA(ctx, cudaGetLastError());
        pts->ncl = ncomp;
        pts->bbox.clear();
        pts->size.clear();
        if (n_clusters) *n_clusters = ncomp;
        return CL2_OK;
    }
    std::vector<C2Box> hb((size_t)ncomp);
    if (ncomp > 0) CL2_CUDA(ctx, cudaMemcpyAsync(hb.data(), box, (size_t)ncomp * sizeof(C2Box), cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += (int64_t)ncomp * (int64_t)sizeof(C2Box);
    CL2_CUDA(ctx, cudaStreamSynchronize(s));
    CL2_CUDA(ctx, cudaGetLastError());
    pts->ncl = ncomp;
    pts->bbox.resize((size_t)ncomp * 4);
    pts->size.resize((size_t)ncomp);
    for (int32_t i = 0; i < ncomp; i++) {
        pts->bbox[4 * (size_t)i] = hb[i].xmin;
        pts->bbox[4 * (size_t)i + 1] = hb[i].xmax;
        pts->bbox[4 * (size_t)i + 2] = hb[i].ymin;
        pts->bbox[4 * (size_t)i + 3] = hb[i].ymax;
        pts->size[i] = hb[i].size;
    }
    if (n_clusters) *n_clusters = ncomp;
    return CL2_OK;
}
