// cl2_api.cu -- context, pinned staging, .ixy upload (int64 -> int32 narrowing + cut/mcut compaction + extents),
// timing read-out.  C ABI declared in include/cl2.h.
#include "cl2_points.cuh"
#include <new>
#include <stdlib.h>

#ifndef CL2_SRC_HASH
#define CL2_SRC_HASH "unknown"
#endif
// the hash of the sources this library was compiled from (cloops2_b200/_lib.py: source_hash)
extern "C" const char *cl2_version(void) { return "cl2-b200 0.2 sm_100a src:" CL2_SRC_HASH; }

extern "C" int cl2_ctx_create(int device, cl2_ctx **out) {
    if (!out) return CL2_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) return CL2_ERR_CUDA;
    cl2_ctx *ctx = new (std::nothrow) cl2_ctx();
    if (!ctx) return CL2_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CL2_ERR_CUDA;
    }
    const char *np = getenv("CL2_NO_POOL");
    ctx->no_pool = (np && np[0] == '1') ? 1 : 0;
    const char *ar = getenv("CL2_ARENA");
    ctx->use_arena = (ar && ar[0] == '0') ? 0 : 1;
    const char *ac = getenv("CL2_ARENA_CHUNK_MB");
    if (ac && atoi(ac) > 0) ctx->arena.chunk_bytes = (size_t)atoi(ac) << 20;
    cudaDeviceGetDefaultMemPool(&ctx->pool, device);
    if (ctx->pool) {
        uint64_t thr = UINT64_MAX;  // keep freed scratch in the pool: steady-state calls never reach the driver allocator
        cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
        // Reserve pool memory up front (once per device and process): growing the pool maps physical memory, which
        // costs milliseconds per call and made step times depend on the order the worker threads picked their files.
        // CL2_POOL_RESERVE_MB overrides the default of 1/8 of the free memory (capped at 16 GB); 0 disables.
        static bool reserved[64] = {false};
        if (!ctx->no_pool && device < 64 && !reserved[device]) {
            reserved[device] = true;
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            size_t want = free_b / 8;
            if (want > ((size_t)16 << 30)) want = (size_t)16 << 30;
            const char *rs = getenv("CL2_POOL_RESERVE_MB");
            if (rs) want = (size_t)strtoull(rs, nullptr, 10) << 20;
            if (want > 0 && want < free_b) {
                void *p = nullptr;
                if (cudaMallocAsync(&p, want, ctx->stream) == cudaSuccess) {
                    cudaFreeAsync(p, ctx->stream);
                    cudaStreamSynchronize(ctx->stream);
                } else {
                    cudaGetLastError();
                }
            }
        }
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaMallocHost(&ctx->h_scratch, 4096) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return CL2_ERR_NOMEM;
    }
    *out = ctx;
    return CL2_OK;
}

extern "C" void cl2_ctx_destroy(cl2_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : ctx->evpool) cudaEventDestroy(e);
    cl2_arena_release(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    cl2_sync_forget(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *cl2_last_error(cl2_ctx *ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int cl2_set_pvalue_margin(cl2_ctx *ctx, double margin) {
    if (!ctx || !(margin >= 0.0) || margin > 1.0) return CL2_ERR_ARG;
    ctx->pvalue_margin = margin;
    return CL2_OK;
}

extern "C" int cl2_ctx_sync(cl2_ctx *ctx) {
    if (!ctx) return CL2_ERR_ARG;
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CL2_OK;
}

extern "C" int cl2_host_alloc(cl2_ctx *ctx, int64_t bytes, void **out) {
    if (!ctx || !out || bytes < 0) return CL2_ERR_ARG;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    CL2_CUDA(ctx, cudaMallocHost(out, (size_t)(bytes ? bytes : 1)));
    return CL2_OK;
}
extern "C" void cl2_host_free(cl2_ctx *ctx, void *p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ upload
struct ExtAcc {
    long long minx, maxx, miny, maxy;
};

__device__ __forceinline__ long long warp_min(long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ long long warp_max(long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// keep flag per row (parseIxy, io.py:282-289)
__device__ __forceinline__ bool keep_row(long long x, long long y, long long cut, long long mcut) {
    long long d = y - x;
    if (cut > 0 && d < cut) return false;
    if (mcut > 0 && d > mcut) return false;
    return true;
}

// int64[n,2] -> int2[n] (no filter), two PETs (32 B) per thread per iteration; extent reduction fused.
__global__ void __launch_bounds__(256) k_convert(const longlong2 *__restrict__ in, int2 *__restrict__ out, int64_t n,
                                                  ExtAcc *__restrict__ ext) {
    long long mnx = LLONG_MAX, mxx = LLONG_MIN, mny = LLONG_MAX, mxy = LLONG_MIN;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        longlong2 p = in[i];
        out[i] = make_int2((int)p.x, (int)p.y);
        mnx = p.x < mnx ? p.x : mnx;
        mxx = p.x > mxx ? p.x : mxx;
        mny = p.y < mny ? p.y : mny;
        mxy = p.y > mxy ? p.y : mxy;
    }
    mnx = warp_min(mnx); mxx = warp_max(mxx); mny = warp_min(mny); mxy = warp_max(mxy);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&ext->minx, mnx);
        atomicMax(&ext->maxx, mxx);
        atomicMin(&ext->miny, mny);
        atomicMax(&ext->maxy, mxy);
    }
}

__global__ void __launch_bounds__(256) k_keepflags(const longlong2 *__restrict__ in, int32_t *__restrict__ flag,
                                                    int64_t n, long long cut, long long mcut) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    longlong2 p = in[i];
    flag[i] = keep_row(p.x, p.y, cut, mcut) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_convert_compact(const longlong2 *__restrict__ in,
                                                          const int32_t *__restrict__ pos, int2 *__restrict__ out,
                                                          int64_t n, long long cut, long long mcut,
                                                          ExtAcc *__restrict__ ext) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long mnx = LLONG_MAX, mxx = LLONG_MIN, mny = LLONG_MAX, mxy = LLONG_MIN;
    if (i < n) {
        longlong2 p = in[i];
        if (keep_row(p.x, p.y, cut, mcut)) {
            out[pos[i]] = make_int2((int)p.x, (int)p.y);
            mnx = mxx = p.x;
            mny = mxy = p.y;
        }
    }
    mnx = warp_min(mnx); mxx = warp_max(mxx); mny = warp_min(mny); mxy = warp_max(mxy);
    if ((threadIdx.x & 31) == 0 && mnx != LLONG_MAX) {
        atomicMin(&ext->minx, mnx);
        atomicMax(&ext->maxx, mxx);
        atomicMin(&ext->miny, mny);
        atomicMax(&ext->maxy, mxy);
    }
}

__global__ void k_ext_init(ExtAcc *e) {
    e->minx = LLONG_MAX; e->maxx = LLONG_MIN; e->miny = LLONG_MAX; e->maxy = LLONG_MIN;
}

__global__ void __launch_bounds__(256) k_download(const int2 *__restrict__ in, longlong2 *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int2 p = in[i];
        out[i] = make_longlong2(p.x, p.y);
    }
}

extern "C" int cl2_points_upload(cl2_ctx *ctx, const int64_t *xy, int64_t n, int64_t cut, int64_t mcut,
                                 cl2_points **out) {
    if (!ctx || !out || n < 0 || (n > 0 && !xy)) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_points_upload: bad argument");
    if (n >= (int64_t)INT32_MAX) return cl2_fail(ctx, CL2_ERR_RANGE, "cl2_points_upload: more than 2^31-2 rows in one file");
    *out = nullptr;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    cl2_points *pts = new (std::nothrow) cl2_points();
    if (!pts) return CL2_ERR_NOMEM;
    int rc = CL2_OK;
    {
        Scratch sc(ctx);
        longlong2 *stage = nullptr;
        ExtAcc *d_ext = nullptr;
        int32_t *flag = nullptr;
        int64_t *d_total = nullptr;
        int64_t kept = n;
        do {
            if ((rc = sc.get(&d_ext, 1)) != CL2_OK) break;
            if ((rc = sc.get(&d_total, 1)) != CL2_OK) break;
            k_ext_init<<<1, 1, 0, ctx->stream>>>(d_ext);
            if (n > 0) {
                if ((rc = sc.get(&stage, (size_t)n)) != CL2_OK) break;
                if (cudaMemcpyAsync(stage, xy, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
                    rc = cl2_fail(ctx, CL2_ERR_CUDA, "cl2_points_upload: host->device copy failed");
                    break;
                }
                ctx->h2d_bytes += n * 16;
                bool filt = cut > 0 || mcut > 0;
                if (!filt) {
                    if ((rc = cl2_dev_alloc(ctx, &pts->xy, (size_t)n)) != CL2_OK) break;
                    FamTimer t(ctx, FAM_CONVERT, 24.0 * (double)n);
                    int grid = ctx->sm_count * 8;
                    int need = cl2_grid(n, 256);
                    k_convert<<<need < grid ? need : grid, 256, 0, ctx->stream>>>(stage, pts->xy, n, d_ext);
                } else {
                    if ((rc = sc.get(&flag, (size_t)n)) != CL2_OK) break;
                    {
                        FamTimer t(ctx, FAM_CONVERT, 20.0 * (double)n);
                        k_keepflags<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(stage, flag, n, cut, mcut);
                    }
                    if ((rc = cl2_exclusive_scan_i32(ctx, flag, n, d_total)) != CL2_OK) break;
                    if (cudaMemcpyAsync(ctx->h_scratch, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                        rc = cl2_fail(ctx, CL2_ERR_CUDA, "cl2_points_upload: scan read-back failed");
                        break;
                    }
                    kept = *(int64_t *)ctx->h_scratch;
                    if ((rc = cl2_dev_alloc(ctx, &pts->xy, (size_t)kept)) != CL2_OK) break;
                    FamTimer t(ctx, FAM_CONVERT, 28.0 * (double)n);
                    k_convert_compact<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(stage, flag, pts->xy, n, cut, mcut, d_ext);
                }
            }
            if (cudaMemcpyAsync(ctx->h_scratch, d_ext, sizeof(ExtAcc), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                rc = cl2_fail(ctx, CL2_ERR_CUDA, "cl2_points_upload: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            ExtAcc h = *(ExtAcc *)ctx->h_scratch;
            pts->n = kept;
            if (kept > 0) {
                pts->ext[0] = h.minx; pts->ext[1] = h.maxx; pts->ext[2] = h.miny; pts->ext[3] = h.maxy;
                if (h.minx < 0 || h.miny < 0 || h.maxx >= (long long)INT32_MAX || h.maxy >= (long long)INT32_MAX) {
                    rc = cl2_fail(ctx, CL2_ERR_RANGE, "cl2_points_upload: coordinates outside [0, 2^31-1)");
                    break;
                }
            }
        } while (0);
    }
    if (rc != CL2_OK) {
        cl2_dev_free(ctx, pts->xy);
        delete pts;
        return rc;
    }
    *out = pts;
    return CL2_OK;
}

extern "C" int64_t cl2_points_count(const cl2_points *pts) { return pts ? pts->n : -1; }

extern "C" int cl2_points_extent(cl2_ctx *ctx, const cl2_points *pts, int64_t ext[4]) {
    if (!ctx || !pts || !ext) return CL2_ERR_ARG;
    for (int i = 0; i < 4; i++) ext[i] = pts->ext[i];
    return CL2_OK;
}

extern "C" int cl2_points_download(cl2_ctx *ctx, const cl2_points *pts, int64_t *xy_out) {
    if (!ctx || !pts || (pts->n > 0 && !xy_out)) return CL2_ERR_ARG;
    if (pts->n == 0) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    longlong2 *tmp;
    CL2_TRY(sc.get(&tmp, (size_t)pts->n));
    k_download<<<cl2_grid(pts->n, 256), 256, 0, ctx->stream>>>(pts->xy, tmp, pts->n);
    ctx->launches++;
    CL2_CUDA(ctx, cudaMemcpyAsync(xy_out, tmp, (size_t)pts->n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += pts->n * 16;
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CL2_OK;
}

void cl2_grid_free(cl2_ctx *ctx, cl2_grid_cache &g) {
    cl2_dev_free(ctx, g.row);
    cl2_dev_free(ctx, g.sxy);
    cl2_dev_free(ctx, g.cid);
    cl2_dev_free(ctx, g.cstart);
    cl2_dev_free(ctx, g.ckey);
    cl2_dev_free(ctx, g.nbr);
    cl2_dev_free(ctx, g.s3);
    cl2_dev_free(ctx, g.first);
    g = cl2_grid_cache();
}

void cl2_subsets_free(cl2_ctx *ctx, cl2_points *pts, size_t keep_levels) {
    while (pts->subs.size() > keep_levels) {
        cl2_dev_free(ctx, pts->subs.back().byy);
        cl2_dev_free(ctx, pts->subs.back().rowy);
        pts->subs.pop_back();
    }
}

extern "C" int cl2_points_sweep_hint(cl2_ctx *ctx, cl2_points *pts, int32_t enable) {
    if (!ctx || !pts) return CL2_ERR_ARG;
    pts->sweep_hint = enable == 0 ? 0 : (enable == 2 ? 2 : 1);
    if (!pts->sweep_hint) {
        cudaSetDevice(ctx->device);
        if (pts->grid.min_minpts >= 0) cl2_grid_free(ctx, pts->grid);
        cl2_subsets_free(ctx, pts);
    }
    return CL2_OK;
}

extern "C" int cl2_points_index_hint(cl2_ctx *ctx, cl2_points *pts, int32_t enable) {
    if (!ctx || !pts) return CL2_ERR_ARG;
    pts->want_xorder = enable ? 1 : 0;
    if (!enable) {
        cudaSetDevice(ctx->device);
        cl2_points_drop_xorder(ctx, pts);
    }
    return CL2_OK;
}

int g_cl2_sync_mode = 0;
extern "C" void cl2_set_sync_mode(int32_t mode) { g_cl2_sync_mode = (mode == 1 || mode == 2) ? mode : 0; }

// The blocking-sync event of every stream that was ever waited on in mode 1: one per stream for the life of the stream
// (worker threads come and go with every job; an event per thread would be created again and again).
#include <mutex>
#include <unordered_map>
#undef cudaStreamSynchronize
static std::mutex g_sync_mu;
static std::unordered_map<cudaStream_t, cudaEvent_t> g_sync_ev;
cudaError_t cl2_sync_stream(cudaStream_t s) {
    if (g_cl2_sync_mode == 0) return cudaStreamSynchronize(s);
    if (g_cl2_sync_mode == 2) {
        for (;;) {
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaErrorNotReady) return e;
            sched_yield();
        }
    }
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_sync_mu);
        auto it = g_sync_ev.find(s);
        if (it != g_sync_ev.end()) {
            ev = it->second;
        } else {
            cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            g_sync_ev.emplace(s, ev);
        }
    }
    cudaError_t e = cudaEventRecord(ev, s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev);
}
void cl2_sync_forget(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_sync_mu);
    auto it = g_sync_ev.find(s);
    if (it == g_sync_ev.end()) return;
    cudaEventDestroy(it->second);
    g_sync_ev.erase(it);
}
#define cudaStreamSynchronize(s) cl2_sync_stream(s)

extern "C" void cl2_points_drop_cache(cl2_ctx *ctx, cl2_points *pts) {
    if (!ctx || !pts) return;
    cudaSetDevice(ctx->device);
    cl2_grid_free(ctx, pts->grid);
    cl2_subsets_free(ctx, pts);
    cl2_dev_free(ctx, pts->byy);     // the y-order is derived data too: the next sweep rebuilds it
    cl2_dev_free(ctx, pts->rowy);
    pts->byy = nullptr;
    pts->rowy = nullptr;
    cl2_points_drop_xorder(ctx, pts);
}

void cl2_points_drop_xorder(cl2_ctx *ctx, cl2_points *pts) {
    cl2_dev_free(ctx, pts->byx);
    cl2_dev_free(ctx, pts->xorder_bad);
    pts->byx = nullptr;
    pts->xorder_bad = nullptr;
}

extern "C" void cl2_points_free(cl2_ctx *ctx, cl2_points *pts) {
    if (!ctx || !pts) return;
    cudaSetDevice(ctx->device);
    cl2_grid_free(ctx, pts->grid);
    cl2_subsets_free(ctx, pts);
    cl2_dev_free(ctx, pts->byy);
    cl2_dev_free(ctx, pts->rowy);
    cl2_points_drop_xorder(ctx, pts);
    cl2_dev_free(ctx, pts->xy);
    delete pts;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int64_t cl2_launch_count(cl2_ctx *ctx) { return ctx ? ctx->launches : -1; }
extern "C" void cl2_launch_count_reset(cl2_ctx *ctx) {
    if (ctx) {
        ctx->launches = 0;
        ctx->h2d_bytes = 0;
        ctx->d2h_bytes = 0;
    }
}
extern "C" void cl2_transfer_bytes(cl2_ctx *ctx, int64_t *h2d, int64_t *d2h) {
    if (!ctx) return;
    if (h2d) *h2d = ctx->h2d_bytes;
    if (d2h) *d2h = ctx->d2h_bytes;
}
extern "C" int cl2_timing_enable(cl2_ctx *ctx, int on) {
    if (!ctx) return CL2_ERR_ARG;
    if (!on) cl2_timing_drain(ctx);
    ctx->timing = on ? 1 : 0;
    return CL2_OK;
}
extern "C" void cl2_timing_reset(cl2_ctx *ctx) {
    if (!ctx) return;
    cl2_timing_drain(ctx);
    for (int i = 0; i < CL2_NFAM; i++) {
        ctx->fam_ms[i] = 0;
        ctx->fam_launch[i] = 0;
        ctx->fam_bytes[i] = 0;
    }
}
extern "C" int cl2_timing_read(cl2_ctx *ctx, int k_max, char *names_out, double *ms_out, int64_t *launches_out,
                               double *bytes_out) {
    if (!ctx) return CL2_ERR_ARG;
    cl2_timing_drain(ctx);
    int k = 0;
    for (int i = 0; i < FAM_COUNT_ && k < k_max; i++) {
        if (ctx->fam_launch[i] == 0) continue;
        if (names_out) {
            strncpy(names_out + 32 * k, CL2_FAM_NAMES[i], 31);
            names_out[32 * k + 31] = 0;
        }
        if (ms_out) ms_out[k] = ctx->fam_ms[i];
        if (launches_out) launches_out[k] = ctx->fam_launch[i];
        if (bytes_out) bytes_out[k] = ctx->fam_bytes[i];
        k++;
    }
    return k;
}
