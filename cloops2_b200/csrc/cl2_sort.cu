// cl2_sort.cu -- hand-written device primitives for the binning stage: a stable LSD radix sort of
// (cell key, uint32 row id) pairs and an int32 exclusive scan.  sm_100a.
//
// The sort in use is the one-kernel-per-pass "onesweep" variant further down (k_os_hist / k_os_pass), on 32-bit keys
// whenever the key fits (every grid, x-order and y-order sort of the path), on 64-bit keys otherwise.  Only
// ceil(nbits/8) passes are run, nbits being the number of significant bits of the key for this chromosome and eps.
// The three-kernel pass right below -- upsweep (per-tile digit histogram) + scan of the [256][ntiles] table +
// downsweep (stable ranking with __match_any_sync, shared-memory reorder so each digit run leaves the SM as a
// contiguous store) -- remains as the fallback for n >= 2^30 rows and as CL2_SORT_3KERNEL=1.
#include "cl2_common.cuh"
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------ scan
#define SCAN_THREADS 256
#define SCAN_IPT 8
#define SCAN_TILE (SCAN_THREADS * SCAN_IPT)

__device__ __forceinline__ int warp_incl_scan(int v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int *total) {
    __shared__ int wsum[THREADS / 32];
    __shared__ int tot;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < THREADS / 32 ? wsum[lane] : 0;
        int si = warp_incl_scan(s);
        if (lane < THREADS / 32) wsum[lane] = si - s;
        if (lane == THREADS / 32 - 1) tot = si;
    }
    __syncthreads();
    int r = inc - v + wsum[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t *__restrict__ d, int64_t n,
                                                               int32_t *__restrict__ sums) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
    int s = 0;
    if (base + SCAN_IPT <= n) {
        const int4 *p = reinterpret_cast<const int4 *>(d + base);
        int4 a = p[0], b = p[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < SCAN_IPT; k++)
            if (base + k < n) s += d[base + k];
    }
    int tot;
    block_excl_scan<SCAN_THREADS>(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int32_t *__restrict__ d, int64_t n,
                                                              const int32_t *__restrict__ sums,
                                                              int64_t *__restrict__ total) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
    int v[SCAN_IPT];
    bool full = base + SCAN_IPT <= n;
    if (full) {
        const int4 *p = reinterpret_cast<const int4 *>(d + base);
        int4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++) v[k] = (base + k < n) ? d[base + k] : 0;
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) s += v[k];
    int tot;
    int ex = block_excl_scan<SCAN_THREADS>(s, &tot);
    int off = (sums ? sums[blockIdx.x] : 0) + ex;
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = (int64_t)(sums ? sums[blockIdx.x] : 0) + tot;
    int o[SCAN_IPT];
#pragma unroll
    for (int k = 0; k < SCAN_IPT; k++) {
        o[k] = off;
        off += v[k];
    }
    if (full) {
        int4 *p = reinterpret_cast<int4 *>(d + base);
        p[0] = make_int4(o[0], o[1], o[2], o[3]);
        p[1] = make_int4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_IPT; k++)
            if (base + k < n) d[base + k] = o[k];
    }
}

static int scan_rec(cl2_ctx *ctx, Scratch &sc, int32_t *d, int64_t n, int64_t *d_total) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb <= 1) {
        k_scan_apply<<<1, SCAN_THREADS, 0, ctx->stream>>>(d, n, nullptr, d_total);
        ctx->launches++;
        return CL2_OK;
    }
    int32_t *sums;
    CL2_TRY(sc.get(&sums, (size_t)nb));
    k_scan_reduce<<<(int)nb, SCAN_THREADS, 0, ctx->stream>>>(d, n, sums);
    CL2_TRY(scan_rec(ctx, sc, sums, nb, nullptr));
    k_scan_apply<<<(int)nb, SCAN_THREADS, 0, ctx->stream>>>(d, n, sums, d_total);
    ctx->launches += 2;
    return CL2_OK;
}

int cl2_exclusive_scan_i32(cl2_ctx *ctx, int32_t *d, int64_t n, int64_t *d_total) {
    if (n <= 0) {
        if (d_total) CL2_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int64_t), ctx->stream));
        return CL2_OK;
    }
    Scratch sc(ctx);
    FamTimer t(ctx, FAM_SCAN, 8.0 * (double)n, 0);
    CL2_TRY(scan_rec(ctx, sc, d, n, d_total));
    CL2_CUDA(ctx, cudaGetLastError());
    return CL2_OK;
}

// ------------------------------------------------------------------------------------------------ radix sort
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_IPT 16
#define RS_TILE (RS_THREADS * RS_IPT)

__global__ void __launch_bounds__(RS_THREADS) k_rs_upsweep(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                            int ntiles, int32_t *__restrict__ tilehist) {
    __shared__ int hist[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    int w = threadIdx.x >> 5;
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_IPT; r++) {
        int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) {
            int d = (int)((keys[i] >> shift) & 255);
            atomicAdd(&hist[w][d], 1);
        }
    }
    __syncthreads();
    int d = threadIdx.x;
    int s = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) s += hist[k][d];
    tilehist[(int64_t)d * ntiles + blockIdx.x] = s;
}

struct RsSmem {
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    int warp_cnt[RS_WARPS][256];
    int dprefix[256];
    int gbase[256];
};

__global__ void __launch_bounds__(RS_THREADS) k_rs_downsweep(const uint64_t *__restrict__ kin,
                                                              const uint32_t *__restrict__ vin,
                                                              uint64_t *__restrict__ kout, uint32_t *__restrict__ vout,
                                                              int64_t n, int shift, int ntiles,
                                                              const int32_t *__restrict__ tilebase, int iota) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem &S = *reinterpret_cast<RsSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&S.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
    const int64_t wbase = tbase + (int64_t)w * (32 * RS_IPT);
    uint64_t key[RS_IPT];
    unsigned short rank[RS_IPT];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        int64_t i = wbase + r * 32 + lane;
        key[r] = (i < n) ? kin[i] : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        int64_t i = wbase + r * 32 + lane;
        bool valid = i < n;
        unsigned d = valid ? (unsigned)((key[r] >> shift) & 255) : (0x10000u | lane);
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int before = 0;
        if (valid) before = S.warp_cnt[w][d];
        __syncwarp();
        rank[r] = (unsigned short)(before + __popc(peers & lt));
        if (valid && (peers & lt) == 0) S.warp_cnt[w][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        // per digit: exclusive prefix over warps, then block scan over digits
        int d = threadIdx.x;
        int off = 0;
#pragma unroll
        for (int k = 0; k < RS_WARPS; k++) {
            int c = S.warp_cnt[k][d];
            S.warp_cnt[k][d] = off;
            off += c;
        }
        int tot;
        int ex = block_excl_scan<RS_THREADS>(off, &tot);
        S.dprefix[d] = ex;
        S.gbase[d] = tilebase[(int64_t)d * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            int d = (int)((key[r] >> shift) & 255);
            int pos = S.dprefix[d] + S.warp_cnt[w][d] + rank[r];
            S.keys[pos] = key[r];
            S.vals[pos] = iota ? (uint32_t)i : vin[i];
        }
    }
    __syncthreads();
    int cnt = (int)((n - tbase) < RS_TILE ? (n - tbase) : RS_TILE);
#pragma unroll 4
    for (int j = threadIdx.x; j < cnt; j += RS_THREADS) {
        uint64_t k = S.keys[j];
        int d = (int)((k >> shift) & 255);
        int64_t dst = (int64_t)S.gbase[d] + (j - S.dprefix[d]);
        kout[dst] = k;
        vout[dst] = S.vals[j];
    }
}

// ------------------------------------------------------------------------------------------------ onesweep variant
// One kernel per pass (16 B/PET with 32-bit keys, 24 with 64-bit): digit histograms of ALL passes are taken once up front (they do not depend on the
// element order), tiles are claimed through an atomic ticket and obtain their per-digit base by decoupled look-back
// over the tile descriptors (count | AGGREGATE, later inclusive | INCLUSIVE), so no per-pass upsweep read and no
// table scan.  Tiles are ticket-ordered, so every predecessor a tile waits on is already running or finished.
__device__ __forceinline__ uint64_t os_key(const cl2_keygen &g, const uint64_t *__restrict__ keys, int64_t i) {
    if (g.kind == CL2_KEY_ARRAY) return keys[i];
    int2 p = g.xy[i];
    if (g.kind == CL2_KEY_BLOCK || g.kind == CL2_KEY_ROT) return cl2_cell_key(g, p);
    if (g.kind == CL2_KEY_GX_YX) return (uint64_t)((unsigned)(p.y - g.minx) / g.eps);   // g.xy holds (y, x) pairs
    return (uint64_t)(uint32_t)(g.kind == CL2_KEY_X ? p.x : p.y);
}

// descriptor load / store at gpu scope (coherent in L2, never served from L1)
__device__ __forceinline__ unsigned os_ld(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void os_st(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#define OS_AGG 0x40000000u
#define OS_INC 0x80000000u
#define OS_VAL 0x3FFFFFFFu
#define OS_MAXPASS 8

template <typename K>
__global__ void __launch_bounds__(256) k_os_hist(const K *__restrict__ keys, const cl2_keygen gen, int64_t n,
                                                  int passes, unsigned *__restrict__ ghist) {
    __shared__ unsigned h[OS_MAXPASS][256];
    for (int i = threadIdx.x; i < OS_MAXPASS * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t nround = ((n + 31) / 32) * 32;      // whole warps stay in the loop together
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        bool valid = i < n;
        uint64_t k = valid ? (gen.kind == CL2_KEY_ARRAY ? (uint64_t)keys[i] : os_key(gen, nullptr, i)) : 0;
        unsigned act = __ballot_sync(0xffffffffu, valid);
        for (int p = 0; p < passes; p++) {
            unsigned d = (unsigned)((k >> (8 * p)) & 255);
            int same;
            __match_all_sync(0xffffffffu, valid ? d : 0x100u + (threadIdx.x & 31), &same);
            if (same) {                      // high digits are often constant: one add per warp instead of 32 conflicts
                if ((threadIdx.x & 31) == 0) atomicAdd(&h[p][d], (unsigned)__popc(act));
            } else if (valid) {
                atomicAdd(&h[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; p++) {
        unsigned v = h[p][threadIdx.x];
        if (v) atomicAdd(&ghist[p * 256 + threadIdx.x], v);
    }
}
// exclusive prefix of each pass's 256-bin histogram, in place; one block per pass
__global__ void __launch_bounds__(256) k_os_prefix(unsigned *__restrict__ ghist) {
    int tot;
    unsigned v = ghist[blockIdx.x * 256 + threadIdx.x];
    int ex = block_excl_scan<256>((int)v, &tot);
    ghist[blockIdx.x * 256 + threadIdx.x] = (unsigned)ex;
}

// One tile of one pass.  NT threads x IPT items; FULL tiles (all but the last) carry no bounds checks, GEN = the keys of
// this pass are generated from the coordinates (first pass of a sort) instead of read from the key buffer.
template <typename K, int NT, int IPT>
struct OsSmem {
    K keys[NT * IPT];
    uint32_t vals[NT * IPT];
    int warp_cnt[NT / 32][256];
    int hist[256];
    int gbase[256];
};

template <typename K, int NT, int IPT, bool GEN, bool FULL>
__device__ __forceinline__ void os_tile(OsSmem<K, NT, IPT> &S, const K *__restrict__ kin, const cl2_keygen &gen,
                                        const uint32_t *__restrict__ vin, K *__restrict__ kout, uint32_t *__restrict__ vout,
                                        const uint32_t n, const int shift, const unsigned *__restrict__ gprefix,
                                        unsigned *__restrict__ desc, const unsigned tile, const bool iota, const bool put_keys) {
    constexpr int TILE = NT * IPT, NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t tbase = tile * (uint32_t)TILE;
    const uint32_t i0 = tbase + (uint32_t)w * (32 * IPT) + lane;
    K key[IPT];
    uint32_t val[IPT];
    unsigned short rank[IPT];
    unsigned peers_of[IPT];
    const unsigned lt = (1u << lane) - 1u;
    // all global loads of the tile are issued up front (keys and row ids) so their latency overlaps the ranking
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        const uint32_t i = i0 + r * 32;
        if (FULL || i < n) key[r] = GEN ? (K)os_key(gen, nullptr, i) : kin[i];
        else key[r] = (K) ~(K)0;
    }
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        const uint32_t i = i0 + r * 32;
        val[r] = iota ? i : ((FULL || i < n) ? vin[i] : 0u);
    }
    // tile histogram first, so the tile's digit counts are published (AGGREGATE) before the longer ranking phase: by the
    // time this tile looks back, its predecessors have usually published their inclusive prefixes (one shared atomic
    // per group of equal digits in the warp)
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        const bool valid = FULL || (i0 + r * 32 < n);
        const unsigned d = valid ? (unsigned)((key[r] >> shift) & 255) : (0x10000u | lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        peers_of[r] = peers;
        if (valid && (peers & lt) == 0) atomicAdd(&S.hist[d], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 256) os_st(&desc[(size_t)tile * 256 + threadIdx.x], (unsigned)S.hist[threadIdx.x] | OS_AGG);
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        const bool valid = FULL || (i0 + r * 32 < n);
        const unsigned d = (unsigned)((key[r] >> shift) & 255);
        const unsigned peers = peers_of[r];
        int before = 0;
        if (valid) before = S.warp_cnt[w][d];
        __syncwarp();
        rank[r] = (unsigned short)(before + __popc(peers & lt));
        if (valid && (peers & lt) == 0) S.warp_cnt[w][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: the warps' counts (kept in registers until the block scan has placed the digit inside the tile), the
    // look-back for the digit's base over earlier tiles, and the two offsets the later phases add to an element's rank:
    // warp_cnt[w][d] := position inside the tile of warp w's first element with digit d; gbase[d] := global position
    // of the tile's first element with digit d minus its position inside the tile
    int off = 0;
    int wc[NW];
    unsigned excl = 0;
    if (NT == 256 || threadIdx.x < 256) {
        const int d = threadIdx.x;
#pragma unroll
        for (int k = 0; k < NW; k++) {
            wc[k] = S.warp_cnt[k][d];
            off += wc[k];
        }
        // look back for the exclusive prefix over earlier tiles, four descriptors per round trip
        int t = (int)tile - 1;
        while (t >= 0) {
            unsigned v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (t - j >= 0) ? os_ld(&desc[(size_t)(t - j) * 256 + d]) : OS_INC;
            bool done = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (done) break;
                if (v[j] & OS_INC) { excl += v[j] & OS_VAL; done = true; t = -1; }
                else if (v[j] & OS_AGG) { excl += v[j] & OS_VAL; t--; }
                else done = true;      // not published yet: re-read from this tile
            }
        }
        os_st(&desc[(size_t)tile * 256 + d], (excl + (unsigned)off) | OS_INC);
    }
    {
        int tot;
        const int ex = block_excl_scan<NT>(off, &tot);
        if (NT == 256 || threadIdx.x < 256) {
            const int d = threadIdx.x;
            int run = ex;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                S.warp_cnt[k][d] = run;
                run += wc[k];
            }
            S.gbase[d] = (int)(gprefix[d] + excl) - ex;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; r++) {
        if (FULL || i0 + r * 32 < n) {
            const int d = (int)((key[r] >> shift) & 255);
            const int pos = S.warp_cnt[w][d] + rank[r];
            S.keys[pos] = key[r];
            S.vals[pos] = val[r];
        }
    }
    __syncthreads();
    const int cnt = FULL ? TILE : (int)(n - tbase);
#pragma unroll 4
    for (int j = threadIdx.x; j < cnt; j += NT) {
        const K k = S.keys[j];
        const int d = (int)((k >> shift) & 255);
        const uint32_t dst = (uint32_t)(S.gbase[d] + j);
        if (put_keys) kout[dst] = k;
        vout[dst] = S.vals[j];
    }
}

template <typename K, int NT, int IPT, int MINB, bool GEN>
__global__ void __launch_bounds__(NT, MINB) k_os_pass(const K *__restrict__ kin, const cl2_keygen gen,
                                                   const uint32_t *__restrict__ vin, K *__restrict__ kout,
                                                   uint32_t *__restrict__ vout, uint32_t n, int shift,
                                                   const unsigned *__restrict__ gprefix, unsigned *__restrict__ desc,
                                                   unsigned *__restrict__ ticket, int flags) {
    const bool iota = flags & 1;
    const bool put_keys = !(flags & 2);   // the last pass of a sort whose caller only wants the row order
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OsSmem<K, NT, IPT> &S = *reinterpret_cast<OsSmem<K, NT, IPT> *>(smem_raw);
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < (NT / 32) * 256 + 256; i += NT) (&S.warp_cnt[0][0])[i] = 0;   // + hist
    __syncthreads();
    const unsigned tile = s_tile;
    if ((tile + 1u) * (uint32_t)(NT * IPT) <= n)
        os_tile<K, NT, IPT, GEN, true>(S, kin, gen, vin, kout, vout, n, shift, gprefix, desc, tile, iota, put_keys);
    else
        os_tile<K, NT, IPT, GEN, false>(S, kin, gen, vin, kout, vout, n, shift, gprefix, desc, tile, iota, put_keys);
}

template <typename K, int NT, int IPT, int MINB>
static int onesweep_passes(cl2_ctx *ctx, const cl2_keygen &gen, K *keys0, K *keys1, uint32_t *vals0, uint32_t *vals1, int64_t n,
                           int passes, K **kout, uint32_t **vout, const uint32_t *vals_first, unsigned *ghist,
                           unsigned *desc, unsigned *ticket, int ntiles) {
    static bool attr_set = false;      // one flag per instantiation
    if (!attr_set) {
        CL2_CUDA(ctx, cudaFuncSetAttribute(k_os_pass<K, NT, IPT, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(OsSmem<K, NT, IPT>)));
        CL2_CUDA(ctx, cudaFuncSetAttribute(k_os_pass<K, NT, IPT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(OsSmem<K, NT, IPT>)));
        attr_set = true;
    }
    const bool want_keys = kout != nullptr;
    K *kin = keys0, *ko = keys1;
    uint32_t *vi = vals0, *vo = vals1;
    cl2_keygen none;
    none.kind = CL2_KEY_ARRAY;
    for (int p = 0; p < passes; p++) {
        {
            FamTimer t(ctx, ctx->sort_family >= 0 ? ctx->sort_family : FAM_SORT_SCATTER,
                       ((p == passes - 1 && !want_keys ? 1.0 : 2.0) * sizeof(K) + 8.0) * (double)n);
            const uint32_t *vin = (p == 0 && vals_first) ? vals_first : vi;
            const int flags = ((p == 0 && !vals_first) ? 1 : 0) | ((p == passes - 1 && !want_keys) ? 2 : 0);
            const bool generated = p == 0 && gen.kind != CL2_KEY_ARRAY;
            auto kern = generated ? k_os_pass<K, NT, IPT, MINB, true> : k_os_pass<K, NT, IPT, MINB, false>;
            kern<<<ntiles, NT, sizeof(OsSmem<K, NT, IPT>), ctx->stream>>>(kin, generated ? gen : none, vin, ko, vo, (uint32_t)n, p * 8,
                                                                          ghist + p * 256, desc + (size_t)p * ntiles * 256,
                                                                          ticket + p, flags);
        }
        K *tk = kin; kin = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    CL2_CUDA(ctx, cudaGetLastError());
    if (kout) *kout = kin;
    *vout = vi;
    return CL2_OK;
}

template <typename K>
static int radix_sort_onesweep(cl2_ctx *ctx, const cl2_keygen &gen, K *keys0, K *keys1, uint32_t *vals0,
                               uint32_t *vals1, int64_t n, int passes, K **kout, uint32_t **vout,
                               const uint32_t *vals_first) {
    // tile shape: 512 threads x 8 items (2 resident blocks per SM) by default; CL2_SORT_CFG=1: 256 x 8 (4 blocks),
    // 2: 256 x 16 (2 blocks)
    static const int cfg = getenv("CL2_SORT_CFG") ? atoi(getenv("CL2_SORT_CFG")) : 0;
    const int tile = cfg == 1 ? 256 * 8 : 512 * 8;
    int ntiles = (int)((n + tile - 1) / tile);
    Scratch sc(ctx);
    unsigned *ghist, *desc, *ticket;
    CL2_TRY(sc.get(&ghist, (size_t)OS_MAXPASS * 256 + OS_MAXPASS));
    CL2_TRY(sc.get(&desc, (size_t)passes * ntiles * 256));
    ticket = ghist + OS_MAXPASS * 256;
    CL2_CUDA(ctx, cudaMemsetAsync(ghist, 0, ((size_t)OS_MAXPASS * 256 + OS_MAXPASS) * 4, ctx->stream));
    CL2_CUDA(ctx, cudaMemsetAsync(desc, 0, (size_t)passes * ntiles * 256 * 4, ctx->stream));
    {
        FamTimer t(ctx, ctx->sort_family >= 0 ? ctx->sort_family : FAM_SORT_HIST, 8.0 * (double)n, 2);
        int grid = ctx->sm_count * 8;
        int need = cl2_grid(n, 256);
        k_os_hist<K><<<need < grid ? need : grid, 256, 0, ctx->stream>>>(keys0, gen, n, passes, ghist);
        k_os_prefix<<<passes, 256, 0, ctx->stream>>>(ghist);
    }
    if (cfg == 1)
        return onesweep_passes<K, 256, 8, 4>(ctx, gen, keys0, keys1, vals0, vals1, n, passes, kout, vout, vals_first, ghist, desc,
                                             ticket, ntiles);
    if (cfg == 2)
        return onesweep_passes<K, 256, 16, 2>(ctx, gen, keys0, keys1, vals0, vals1, n, passes, kout, vout, vals_first, ghist, desc,
                                              ticket, ntiles);
    return onesweep_passes<K, 512, 8, 2>(ctx, gen, keys0, keys1, vals0, vals1, n, passes, kout, vout, vals_first, ghist, desc,
                                         ticket, ntiles);
}

__global__ void __launch_bounds__(256) k_os_materialise(const cl2_keygen gen, uint64_t *__restrict__ keys, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = os_key(gen, keys, i);
}

int cl2_radix_sort_gen(cl2_ctx *ctx, const cl2_keygen &gen, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0,
                       uint32_t *vals1, int64_t n, int nbits, uint64_t **kout, uint32_t **vout,
                       const uint32_t *vals_first) {
    *kout = keys0;
    *vout = vals0;
    if (n <= 0) return CL2_OK;
    if (nbits < 1) nbits = 1;
    int passes = (nbits + 7) / 8;
    if (n < (int64_t)OS_VAL && passes <= OS_MAXPASS && !getenv("CL2_SORT_3KERNEL")) {
        if (nbits <= 32 && gen.kind != CL2_KEY_ARRAY && !getenv("CL2_SORT_U64")) {
            // every generated key of this path fits 32 bits: sort (u32 key, u32 row) pairs, 16 B per PET per pass.
            // The caller's 8-byte key buffers are simply used as 4-byte ones; sorted keys are not handed back.
            *kout = nullptr;
            return radix_sort_onesweep<uint32_t>(ctx, gen, (uint32_t *)keys0, (uint32_t *)keys1, vals0, vals1, n, passes,
                                                 nullptr, vout, vals_first);
        }
        return radix_sort_onesweep<uint64_t>(ctx, gen, keys0, keys1, vals0, vals1, n, passes, kout, vout, vals_first);
    }
    if (vals_first) return cl2_fail(ctx, CL2_ERR_RANGE, "radix sort: more than 2^30 rows with a value input");
    if (gen.kind != CL2_KEY_ARRAY) {
        FamTimer t(ctx, FAM_KEYGEN, 16.0 * (double)n);
        k_os_materialise<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(gen, keys0, n);
    }
    return cl2_radix_sort_pairs(ctx, keys0, keys1, vals0, vals1, n, nbits, kout, vout);
}

int cl2_radix_sort_pairs(cl2_ctx *ctx, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0, uint32_t *vals1, int64_t n,
                         int nbits, uint64_t **kout, uint32_t **vout) {
    *kout = keys0;
    *vout = vals0;
    if (n <= 0) return CL2_OK;
    if (nbits < 1) nbits = 1;
    int passes = (nbits + 7) / 8;
    if (n < (int64_t)OS_VAL && passes <= OS_MAXPASS && !getenv("CL2_SORT_3KERNEL")) {
        cl2_keygen none;
        none.kind = CL2_KEY_ARRAY;
        return radix_sort_onesweep<uint64_t>(ctx, none, keys0, keys1, vals0, vals1, n, passes, kout, vout, nullptr);
    }
    int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    Scratch sc(ctx);
    int32_t *tilehist;
    CL2_TRY(sc.get(&tilehist, (size_t)256 * ntiles));
    static bool attr_set = false;
    if (!attr_set) {
        CL2_CUDA(ctx, cudaFuncSetAttribute(k_rs_downsweep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(RsSmem)));
        attr_set = true;
    }
    uint64_t *kin = keys0, *ko = keys1;
    uint32_t *vi = vals0, *vo = vals1;
    for (int p = 0; p < passes; p++) {
        int shift = p * 8;
        {
            FamTimer t(ctx, FAM_SORT_HIST, 8.0 * (double)n);
            k_rs_upsweep<<<ntiles, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, ntiles, tilehist);
        }
        CL2_TRY(cl2_exclusive_scan_i32(ctx, tilehist, (int64_t)256 * ntiles, nullptr));
        {
            FamTimer t(ctx, FAM_SORT_SCATTER, 24.0 * (double)n);
            k_rs_downsweep<<<ntiles, RS_THREADS, sizeof(RsSmem), ctx->stream>>>(kin, vi, ko, vo, n, shift, ntiles,
                                                                                 tilehist, p == 0 ? 1 : 0);
        }
        uint64_t *tk = kin; kin = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    CL2_CUDA(ctx, cudaGetLastError());
    *kout = kin;
    *vout = vi;
    return CL2_OK;
}
