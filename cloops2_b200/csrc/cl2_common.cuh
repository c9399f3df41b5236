// cl2_common.cuh -- context, stream-ordered scratch memory, error handling, per-kernel timing.
#pragma once
#include <cuda_runtime.h>

// Every wait on a stream goes through here.  Mode 0: cudaStreamSynchronize (the runtime spins: lowest latency, one
// busy host core per waiting thread).  Mode 1: record a blocking-sync event and sleep on it.  Mode 2: poll the stream
// and give the core away between polls (sched_yield): as quick as spinning while cores are idle, but a waiting thread
// no longer keeps a runnable one -- of this rank or another -- off the core when there are more workers than cores
// (8 ranks x 8 workers on a 32-core box).
#include <sched.h>
extern int g_cl2_sync_mode;
cudaError_t cl2_sync_stream(cudaStream_t s);               // cl2_api.cu
void cl2_sync_forget(cudaStream_t s);                       // the stream is about to be destroyed
#define cudaStreamSynchronize(s) cl2_sync_stream(s)
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <string>
#include <map>
#include <set>
#include "../../include/cl2.h"

#define CL2_NFAM 32

enum cl2_family {
    FAM_CONVERT = 0,   // int64 -> int32 narrowing + extent reduction + cut/mcut compaction
    FAM_KEYGEN,        // cell key generation
    FAM_SORT_HIST,     // radix sort: per-tile digit histogram (upsweep)
    FAM_SORT_SCATTER,  // radix sort: stable scatter (downsweep)
    FAM_SCAN,          // exclusive scans
    FAM_GATHER,        // sorted-order coordinate gather
    FAM_CELLS,         // segment heads -> cell table
    FAM_NBR,           // neighbour cell lookup + 3x3 sums
    FAM_NOISE,         // noise-cell pruning
    FAM_CELLSTAT,      // per-cell sums / bbox (centroids)
    FAM_ADJ,           // cell adjacency (centroid test + min-pair L1)
    FAM_CORE,          // near sums / core flags / neighbour counts
    FAM_UNION,         // lock-free union-find + pointer jumping
    FAM_COMP,          // component ordering / numbering
    FAM_BORDER,        // border assignment
    FAM_LABEL,         // per-point label scatter
    FAM_BBOX,          // cluster bounding boxes
    FAM_INDEX,         // XY index build (two sorts + gathers)
    FAM_COUNT,         // rectangle / window counting
    FAM_MISC,
    FAM_COUNT_
};

static const char *const CL2_FAM_NAMES[] = {
    "convert", "keygen", "sort_hist", "sort_scatter", "scan", "gather", "cells", "nbr", "noise", "cellstat",
    "adj", "core", "union", "comp", "border", "label", "bbox", "index", "count", "misc"};

struct cl2_evpair {
    cudaEvent_t a, b;
    int fam;
};

// Scope-bound scratch of one context comes from an arena kept by the context instead of one cudaMallocAsync /
// cudaFreeAsync pair per buffer: a clustering run asks for ~35 short-lived arrays, and with several worker threads
// driving one device the runtime calls -- not the kernels -- are what a rank with few host cores runs out of.  All work
// of a context is ordered on its one stream, so a block handed back by a scope may be handed out again at once (the
// same guarantee the stream-ordered pool gives).  Blocks are carved first-fit from chunks taken from that pool; requests
// above CL2_ARENA_MAX_REQ, and everything when CL2_ARENA=0 or CL2_NO_POOL=1, go to the pool / cudaMalloc directly.
#define CL2_ARENA_ALIGN ((size_t)512)
#define CL2_ARENA_MAX_REQ ((size_t)128 << 20)
#define CL2_ARENA_CHUNK ((size_t)256 << 20)
struct cl2_arena {
    std::map<char *, size_t> free_blocks;      // address -> bytes, neighbours inside one chunk coalesced
    std::vector<void *> chunks;
    std::set<char *> chunk_base;               // a block never spans two chunks: cudaMemsetAsync / cudaMemcpyAsync
                                               // reject a range that is not inside ONE allocation, adjacent or not
    size_t held = 0;
    size_t chunk_bytes = CL2_ARENA_CHUNK;      // CL2_ARENA_CHUNK_MB (tests use small chunks to get many of them)
    void add_chunk(char *c, size_t bytes) {
        chunks.push_back(c);
        chunk_base.insert(c);
        held += bytes;
        put(c, bytes);
    }
    void put(char *p, size_t bytes) {
        auto nx = free_blocks.lower_bound(p);
        if (nx != free_blocks.end() && p + bytes == nx->first && !chunk_base.count(nx->first)) {
            bytes += nx->second;
            nx = free_blocks.erase(nx);
        }
        if (nx != free_blocks.begin() && !chunk_base.count(p)) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == p) {
                pv->second += bytes;
                return;
            }
        }
        free_blocks.emplace(p, bytes);
    }
    char *take(size_t bytes) {
        for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it) {
            if (it->second < bytes) continue;
            char *p = it->first;
            const size_t rest = it->second - bytes;
            free_blocks.erase(it);
            if (rest) free_blocks.emplace(p + bytes, rest);
            return p;
        }
        return nullptr;
    }
};

struct cl2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    char err[512] = {0};
    int64_t launches = 0;
    int sort_family = -1;      // >= 0: the radix sort books its launches under this family (small auxiliary sorts)
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int timing = 0;
    double fam_ms[CL2_NFAM] = {0};
    int64_t fam_launch[CL2_NFAM] = {0};
    double fam_bytes[CL2_NFAM] = {0};
    std::vector<cl2_evpair> pending;
    std::vector<cudaEvent_t> evpool;
    int sm_count = 148;
    int no_pool = 0;            // CL2_NO_POOL=1: plain cudaMalloc/cudaFree (used when profiling under ncu)
    void *h_scratch = nullptr;  // small pinned buffer for scalar read-backs
    double pvalue_margin = 0.0; // cl2_set_pvalue_margin
    cl2_arena arena;
    int use_arena = 1;          // CL2_ARENA=0: every scratch buffer straight from the pool
};

// a block of the context's arena (null: none to be had -- the caller then asks the pool)
inline void *cl2_arena_alloc(cl2_ctx *ctx, size_t bytes) {
    bytes = (bytes + CL2_ARENA_ALIGN - 1) / CL2_ARENA_ALIGN * CL2_ARENA_ALIGN;
    char *p = ctx->arena.take(bytes);
    if (p) return p;
    const size_t chunk = bytes > ctx->arena.chunk_bytes ? bytes : ctx->arena.chunk_bytes;
    void *c = nullptr;
    if (cudaMallocAsync(&c, chunk, ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    ctx->arena.add_chunk((char *)c, chunk);
    return ctx->arena.take(bytes);
}
inline void cl2_arena_free(cl2_ctx *ctx, void *p, size_t bytes) {
    bytes = (bytes + CL2_ARENA_ALIGN - 1) / CL2_ARENA_ALIGN * CL2_ARENA_ALIGN;
    ctx->arena.put((char *)p, bytes);
}
inline void cl2_arena_release(cl2_ctx *ctx) {          // context teardown
    for (void *c : ctx->arena.chunks) cudaFreeAsync(c, ctx->stream);
    ctx->arena.chunks.clear();
    ctx->arena.chunk_base.clear();
    ctx->arena.free_blocks.clear();
    ctx->arena.held = 0;
}

inline int cl2_fail(cl2_ctx *ctx, int code, const char *fmt, const char *a = "", const char *b = "") {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b);
    return code;
}

#define CL2_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s failed: %s (%s:%d)", #call,        \
                              cudaGetErrorString(e_), __FILE__, __LINE__);                           \
            return CL2_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define CL2_TRY(call)            \
    do {                         \
        int rc_ = (call);        \
        if (rc_ != CL2_OK) return rc_; \
    } while (0)

// Scratch allocations released when the scope ends: blocks of the context's arena (no runtime call), or -- large
// requests, get_pool(), CL2_ARENA=0 -- stream-ordered allocations on the context's pool (cudaMallocAsync; the pool
// keeps freed blocks, so steady-state calls do not touch the driver allocator).
struct Scratch {
    cl2_ctx *ctx;
    std::vector<void *> ptrs;                          // pool allocations
    struct Blk { void *p; size_t bytes; };
    std::vector<Blk> blks;                             // arena blocks
    explicit Scratch(cl2_ctx *c) : ctx(c) {}
    ~Scratch() {
        for (const Blk &b : blks) cl2_arena_free(ctx, b.p, b.bytes);
        if (ctx->no_pool && !ptrs.empty()) cudaStreamSynchronize(ctx->stream);
        for (void *p : ptrs) {
            if (ctx->no_pool) cudaFree(p); else cudaFreeAsync(p, ctx->stream);
        }
    }
    template <typename T>
    int get(T **out, size_t count) {
        const size_t bytes = (count ? count : 1) * sizeof(T);
        if (ctx->use_arena && !ctx->no_pool && bytes <= CL2_ARENA_MAX_REQ) {
            void *p = cl2_arena_alloc(ctx, bytes);
            if (p) {
                blks.push_back({p, bytes});
                *out = (T *)p;
                return CL2_OK;
            }
        }
        return get_pool(out, count);
    }
    // always a pool allocation: buffers whose ownership may be handed on with release()
    template <typename T>
    int get_pool(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = (count ? count : 1) * sizeof(T);
        cudaError_t e = ctx->no_pool ? cudaMalloc(&p, bytes) : cudaMallocAsync(&p, bytes, ctx->stream);
        if (e != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            *out = nullptr;
            return CL2_ERR_NOMEM;
        }
        ptrs.push_back(p);
        *out = (T *)p;
        return CL2_OK;
    }
    // hand ownership to the caller (long-lived buffers; must come from get_pool)
    void release(void *p) {
        for (size_t i = 0; i < ptrs.size(); i++)
            if (ptrs[i] == p) {
                ptrs.erase(ptrs.begin() + i);
                return;
            }
    }
};

template <typename T>
inline int cl2_dev_alloc(cl2_ctx *ctx, T **out, size_t count) {
    void *p = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t e = ctx->no_pool ? cudaMalloc(&p, bytes) : cudaMallocAsync(&p, bytes, ctx->stream);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        *out = nullptr;
        return CL2_ERR_NOMEM;
    }
    *out = (T *)p;
    return CL2_OK;
}
inline void cl2_dev_free(cl2_ctx *ctx, void *p) {
    if (!p) return;
    if (ctx->no_pool) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(p);
    } else {
        cudaFreeAsync(p, ctx->stream);
    }
}

inline void cl2_timing_drain(cl2_ctx *ctx);

// RAII timing scope: records an event pair around the launches of one kernel family when timing is on.
struct FamTimer {
    cl2_ctx *ctx;
    int fam;
    cudaEvent_t a = nullptr, b = nullptr;
    FamTimer(cl2_ctx *c, int f, double bytes, int nlaunch = 1) : ctx(c), fam(f) {
        ctx->launches += nlaunch;
        if (!ctx->timing) return;
        if (ctx->pending.size() >= 2048) cl2_timing_drain(ctx);   // recycle events (creation is ~10 us each)
        ctx->fam_launch[f] += nlaunch;
        ctx->fam_bytes[f] += bytes;
        a = take();
        b = take();
        cudaEventRecord(a, ctx->stream);
    }
    ~FamTimer() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->pending.push_back({a, b, fam});
    }
    cudaEvent_t take() {
        cudaEvent_t e;
        if (!ctx->evpool.empty()) {
            e = ctx->evpool.back();
            ctx->evpool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
};

inline void cl2_timing_drain(cl2_ctx *ctx) {
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        ctx->fam_ms[p.fam] += ms;
        ctx->evpool.push_back(p.a);
        ctx->evpool.push_back(p.b);
    }
    ctx->pending.clear();
}

static inline int cl2_grid(int64_t n, int block) { return (int)((n + block - 1) / block); }

// ---- shared device-side primitives (cl2_sort.cu) ----
// Where the sort's first pass takes its keys from: a materialised array, or generated on the fly from the resident
// coordinates (saves the separate key-generation pass: 8 B read + 8 B written per PET).
#define CL2_KEY_ARRAY 0
#define CL2_KEY_BLOCK 1   // blockDBSCAN grid: ((x-minx)/eps << by) | (y-miny)/eps
#define CL2_KEY_ROT 2     // cDBSCAN2 grid: truncating division of x-y and x+y, shifted by gx0 / gy0
#define CL2_KEY_X 3       // x coordinate (XY index)
#define CL2_KEY_Y 4       // y coordinate
#define CL2_KEY_GX_YX 5   // blockDBSCAN column index (x-minx)/eps from the y-sorted (y,x) array
struct cl2_keygen {
    int kind = CL2_KEY_ARRAY;
    const int2 *xy = nullptr;
    int minx = 0, miny = 0;
    unsigned eps = 1;
    int by = 0;
    int gx0 = 0;
    unsigned gy0 = 0;
};
// vals_first (optional): values of the unsorted input (read by the first pass only, never written); when null the
// first pass generates 0..n-1.
int cl2_radix_sort_gen(cl2_ctx *ctx, const cl2_keygen &gen, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0,
                       uint32_t *vals1, int64_t n, int nbits, uint64_t **kout, uint32_t **vout,
                       const uint32_t *vals_first = nullptr);

#ifdef __CUDACC__
// packed cell key of a point for the two grids (blockDBSCAN.py:81-82 on non-negative ints; cDBSCAN2.py:67-70 with
// truncation toward zero, shifted so both fields are non-negative)
__device__ __forceinline__ uint64_t cl2_cell_key(const cl2_keygen &g, int2 p) {
    if (g.kind == CL2_KEY_ROT) {
        int u = p.x - p.y;
        unsigned v = (unsigned)p.x + (unsigned)p.y;
        unsigned gx = (unsigned)(u / (int)g.eps - g.gx0), gy = v / g.eps - g.gy0;
        return ((uint64_t)gx << g.by) | (uint64_t)gy;
    }
    unsigned gx = (unsigned)(p.x - g.minx) / g.eps, gy = (unsigned)(p.y - g.miny) / g.eps;
    return ((uint64_t)gx << g.by) | (uint64_t)gy;
}
#endif
// Stable LSD radix sort of (uint64 key, uint32 val) pairs on bits [0, nbits).  keys/vals are double buffers;
// on return *kout / *vout point at the buffer holding the sorted data.
int cl2_radix_sort_pairs(cl2_ctx *ctx, uint64_t *keys0, uint64_t *keys1, uint32_t *vals0, uint32_t *vals1, int64_t n,
                         int nbits, uint64_t **kout, uint32_t **vout);
// Exclusive prefix sum of int32 in place (n up to 2^31-1); total written to *d_total (device int64) when non-null.
int cl2_exclusive_scan_i32(cl2_ctx *ctx, int32_t *d, int64_t n, int64_t *d_total);
