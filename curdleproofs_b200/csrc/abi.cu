// C ABI (include/cdp_msm.h) over the sm_100a kernels.  Host-side glue only: buffer management, launch geometry,
// chunking of large MSMs.  No arithmetic happens on the host and there is no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#pragma GCC visibility push(default)
#include "../../include/cdp_msm.h"
#pragma GCC visibility pop
#include "launch.h"

using namespace cdp;

namespace {

struct scratch_t {
    void *ptr = nullptr;
    size_t cap = 0;
};

}  // namespace

struct cdp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err = "ok";
    uint64_t launches = 0;
    // grow-on-demand device scratch
    scratch_t d_pts, d_scalars, d_segs, d_win, d_jac, d_aux, d_out;
    scratch_t d_dig;  // digit rows of the small-MSM path
    scratch_t d_big;  // large-MSM workspace (keys, values, offsets, buckets, weights, sort temp)
    // pinned host staging
    scratch_t h_stage;
    int sm_count = 148;
    size_t big_ba_min = 0;   // cdp_set_big_ba_min: pairs from which the large path sums its buckets by batched affine rounds (0 = built-in)
    size_t big_msm_min = 0;  // cdp_set_big_msm_min: pairs from which one MSM takes the sort-based path (0 = the built-in threshold)
    // optional per-kernel profiling (cdp_profile_*): CUDA events around every launch on the context's stream
    bool profiling = false;
    struct prof_rec { int kind; cudaEvent_t e0, e1; uint64_t units; };
    std::vector<prof_rec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[CDP_PROFILE_KINDS] = {0};
    uint64_t prof_launches[CDP_PROFILE_KINDS] = {0};
    uint64_t prof_units[CDP_PROFILE_KINDS] = {0};
};

namespace {

int fail(cdp_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define CUDA_TRY(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, CDP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
    } while (0)

int ensure_dev(cdp_ctx *ctx, scratch_t &s, size_t bytes) {
    if (bytes <= s.cap) return CDP_OK;
    if (s.ptr) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaFree(s.ptr));
        s.ptr = nullptr;
        s.cap = 0;
    }
    size_t cap = std::max<size_t>(bytes, 1 << 16);
    cap = (cap + (cap >> 2) + 255) & ~size_t(255);
    CUDA_TRY(ctx, cudaMalloc(&s.ptr, cap));
    s.cap = cap;
    return CDP_OK;
}
int ensure_host(cdp_ctx *ctx, scratch_t &s, size_t bytes) {
    if (bytes <= s.cap) return CDP_OK;
    if (s.ptr) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaFreeHost(s.ptr));
        s.ptr = nullptr;
        s.cap = 0;
    }
    size_t cap = std::max<size_t>(bytes, 1 << 16);
    cap = (cap + (cap >> 2) + 255) & ~size_t(255);
    CUDA_TRY(ctx, cudaMallocHost(&s.ptr, cap));
    s.cap = cap;
    return CDP_OK;
}
// RAII bracket around one kernel launch: counts it and, when profiling is on, times it with CUDA events
struct launch_scope {
    cdp_ctx *ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int kind;
    uint64_t units;
    launch_scope(cdp_ctx *c, int k, uint64_t u) : ctx(c), kind(k), units(u) {
        ctx->launches++;
        if (!ctx->profiling) return;
        auto get = [&]() {
            cudaEvent_t e = nullptr;
            if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        e0 = get(); e1 = get();
        cudaEventRecord(e0, ctx->stream);
    }
    ~launch_scope() {
        if (!e0) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->prof_pending.push_back({kind, e0, e1, units});
    }
};
#define TRY(expr)                      \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != CDP_OK) return rc__; \
    } while (0)

// ---- MSM geometry -------------------------------------------------------------------------------------------
struct msm_cfg {
    int c, nwin;
};
// Window width by MSM size.  Cost model per MSM in mixed-addition equivalents (GLV doubles the point count):
//   accumulate  2n * nwin(c)     combine  2^(c-1) * (0.58 * 128 + 1.55 * nwin(c))   (Horner per bucket index, one reduction)
// c = 6 from n ~ 400, c = 5 from n ~ 50 (e.g. n = 256: 15.1k at c = 5, 14.7k at c = 6; n = 128: 8.5k at c = 5, 9.4k at c = 4).
// Override for tuning: CDP_MSM_THRESH="T6,T5,T4,T3" (smallest n that uses c = 6, 5, 4, 3).
msm_cfg pick_cfg(size_t max_n) {
    static size_t T[4] = {0, 0, 0, 0};
    if (T[3] == 0) {
        size_t d[4] = {384, 48, 12, 4};
        if (const char *e = getenv("CDP_MSM_THRESH")) sscanf(e, "%zu,%zu,%zu,%zu", &d[0], &d[1], &d[2], &d[3]);
        for (int i = 0; i < 4; i++) T[i] = d[i] ? d[i] : 1;
    }
    msm_cfg g;
    g.c = max_n >= T[0] ? 6 : max_n >= T[1] ? 5 : max_n >= T[2] ? 4 : max_n >= T[3] ? 3 : 2;
    g.nwin = msm_nwin_for(g.c);
    return g;
}
constexpr size_t SMALL_MSM_MAX_N = 2048;  // uint16 point ids with a sign bit (2n < 2^15); 4 warps x 12 KB of shared memory at c = 6

// bucket sums for `count` segments -> d_win[count][nwin][2^(c-1)]
int msm_buckets_dev(cdp_ctx *ctx, const msm_cfg &g, const uint8_t *d_pts, const uint8_t *d_scalars, const msm_seg_t *d_segs, size_t count,
                    size_t nmax, uint32_t *d_win, uint64_t pairs) {
    TRY(ensure_dev(ctx, ctx->d_dig, msm_dig_bytes(g.c, nmax, count)));
    launch_scope ls(ctx, CDP_PROFILE_MSM_BUCKETS, pairs);
    ctx->launches++;  // digit kernel + bucket kernel
    CUDA_TRY(ctx, launch_msm_buckets(ctx->stream, g.c, reinterpret_cast<const uint32_t *>(d_pts), reinterpret_cast<const uint32_t *>(d_scalars),
                                     d_segs, (uint32_t)count, (uint32_t)nmax, reinterpret_cast<int8_t *>(ctx->d_dig.ptr), d_win));
    return CDP_OK;
}

int combine_dev(cdp_ctx *ctx, const msm_cfg &g, const uint32_t *d_win, size_t count, uint32_t *d_out_jac) {
    // few MSMs in flight: the combine is a latency chain (130 dependent doublings); run it with a quad of lanes per (MSM, bucket index)
    // (k_msm_horner_quad, ~2.4x shorter), then the weighted reduction over the bucket indices as a combine with one window.
    // CDP_COMBINE_QUAD_MAX: largest launch (in MSMs) that takes this route (0 disables).
    static const size_t quad_max = [] { const char *e = getenv("CDP_COMBINE_QUAD_MAX"); return e ? (size_t)atoll(e) : (size_t)1024; }();
    if (count <= quad_max && g.nwin > 1) {
        const size_t nb = (size_t)1 << (g.c - 1);
        // S, then (for launches of very few MSMs) the group sums of the two-step Horner
        const bool groups = count <= horner_groups_max();
        TRY(ensure_dev(ctx, ctx->d_aux, count * nb * 144 + (groups ? horner_groups_scratch_bytes((uint32_t)count, g.c, g.nwin) : 0)));
        uint32_t *S = reinterpret_cast<uint32_t *>(ctx->d_aux.ptr);
        {
            launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, count);
            CUDA_TRY(ctx, launch_msm_horner_quad(ctx->stream, d_win, S, (uint32_t)count, g.c, g.nwin, groups ? S + 36 * count * nb : nullptr));
            if (groups) ctx->launches++;
        }
        launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, count);
        if (g.c >= 2) CUDA_TRY(ctx, launch_msm_reduce_quad(ctx->stream, S, d_out_jac, (uint32_t)count, g.c));
        else CUDA_TRY(ctx, launch_msm_combine(ctx->stream, S, d_out_jac, (uint32_t)count, g.c, 1));
        return CDP_OK;
    }
    launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, count);
    CUDA_TRY(ctx, launch_msm_combine(ctx->stream, d_win, d_out_jac, (uint32_t)count, g.c, g.nwin));
    return CDP_OK;
}

int normalize_dev(cdp_ctx *ctx, const uint8_t *d_jac, size_t n, uint8_t *d_aff, uint8_t *d_comp, const smul_job_t *jobs = nullptr,
                  uint32_t elems_per_job = 1) {
    if (n == 0) return CDP_OK;
    const uint32_t *j = reinterpret_cast<const uint32_t *>(d_jac);
    uint32_t *a = reinterpret_cast<uint32_t *>(d_aff);
    // enough threads to fill the machine first, then amortise the inversion over a chunk
    size_t fill = (size_t)ctx->sm_count * 256;
    int chunk = n >= 8 * fill ? 8 : n >= 2 * fill ? 2 : 1;
    launch_scope ls(ctx, CDP_PROFILE_NORMALIZE, n);
    CUDA_TRY(ctx, launch_normalize(ctx->stream, chunk, j, a, d_comp, (uint32_t)n, jobs, elems_per_job));
    return CDP_OK;
}

// d_pts[out] = affine(d_pts[add] + s * d_pts[src]) for every element of every job
int smul_jobs_dev(cdp_ctx *ctx, uint8_t *d_pts, const uint8_t *d_scalars, const smul_job_t *d_jobs, size_t n_jobs, size_t epj) {
    size_t total = n_jobs * epj;
    if (total == 0) return CDP_OK;
    if (total >= (size_t(1) << 31)) return fail(ctx, CDP_ERR_TOO_LARGE, "smul jobs: too many elements");
    TRY(ensure_dev(ctx, ctx->d_jac, total * CDP_JACOBIAN_BYTES));
    {
        launch_scope ls(ctx, CDP_PROFILE_SMUL, total);
        CUDA_TRY(ctx, launch_smul_jobs(ctx->stream, reinterpret_cast<const uint32_t *>(d_pts), reinterpret_cast<const uint32_t *>(d_scalars), d_jobs,
                                       (uint32_t)n_jobs, (uint32_t)epj, reinterpret_cast<uint32_t *>(ctx->d_jac.ptr)));
    }
    return normalize_dev(ctx, (const uint8_t *)ctx->d_jac.ptr, total, d_pts, nullptr, d_jobs, (uint32_t)epj);
}

const uint8_t INF_JAC_ZERO[CDP_JACOBIAN_BYTES] = {0};

// Host-to-device copy of a caller's (pageable) buffer, queued on the context's stream.  A cudaMemcpyAsync from pageable memory is staged by
// the driver at ~1-3 GB/s; large inputs of the host-buffer drop-ins (cdp_msm at 2^20+ pairs: 128 B per pair) go through two pinned
// 8 MiB buffers instead, filled by a few host threads while the previous chunk is on the wire.
int h2d_pipelined(cdp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    constexpr size_t CH = size_t(8) << 20;
    if (bytes <= (size_t(1) << 20)) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return CDP_OK;
    }
    TRY(ensure_host(ctx, ctx->h_stage, 2 * CH));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // earlier users of the staging buffer are done
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (auto &e : ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>(4, std::max(1u, hw / 2));
    uint8_t *stage = (uint8_t *)ctx->h_stage.ptr;
    const uint8_t *src = (const uint8_t *)h_src;
    int rc = CDP_OK;
    size_t k = 0;
    for (size_t off = 0; off < bytes && rc == CDP_OK; off += CH, k++) {
        const size_t b = k & 1, len = std::min(CH, bytes - off);
        if (k >= 2 && cudaEventSynchronize(ev[b]) != cudaSuccess) { rc = fail(ctx, CDP_ERR_CUDA, "h2d_pipelined: event"); break; }
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++)
            th.emplace_back([=] { memcpy(stage + b * CH + len * t / nt, src + off + len * t / nt, len * (t + 1) / nt - len * t / nt); });
        memcpy(stage + b * CH, src + off, len / nt);
        for (auto &x : th) x.join();
        if (cudaMemcpyAsync((uint8_t *)d_dst + off, stage + b * CH, len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaEventRecord(ev[b], ctx->stream) != cudaSuccess)
            rc = fail(ctx, CDP_ERR_CUDA, "h2d_pipelined: copy");
    }
    if (rc == CDP_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, CDP_ERR_CUDA, "h2d_pipelined: sync");  // the staging buffer is free again
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

}  // namespace

static void prof_drain(cdp_ctx *ctx);

// =================================================================================================== context
extern "C" int cdp_ctx_create(cdp_ctx **out, int device_id, void *stream) {
    if (!out) return CDP_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device_id < 0 || device_id >= ndev) return CDP_ERR_CUDA;
    if (cudaSetDevice(device_id) != cudaSuccess) return CDP_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return CDP_ERR_CUDA;
    if (prop.major != 10) return CDP_ERR_CUDA;  // sm_100a cubins only; fail loudly on anything else
    cdp_ctx *ctx = new cdp_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return CDP_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    *out = ctx;
    return CDP_OK;
}
extern "C" void cdp_ctx_destroy(cdp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (scratch_t *s : {&ctx->d_pts, &ctx->d_scalars, &ctx->d_segs, &ctx->d_win, &ctx->d_jac, &ctx->d_aux, &ctx->d_out, &ctx->d_big, &ctx->d_dig})
        if (s->ptr) cudaFree(s->ptr);
    if (ctx->h_stage.ptr) cudaFreeHost(ctx->h_stage.ptr);
    prof_drain(ctx);
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char *cdp_last_error(const cdp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" uint64_t cdp_launch_count(const cdp_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int cdp_ctx_device(const cdp_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" int cdp_sync(cdp_ctx *ctx) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}

// =================================================================================================== device-resident
extern "C" void *cdp_dev_alloc(cdp_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        ctx->err = "cudaMalloc failed";
        return nullptr;
    }
    return p;
}
extern "C" void cdp_dev_free(cdp_ctx *ctx, void *d_ptr) {
    if (!ctx || !d_ptr) return;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ptr);
}
extern "C" void *cdp_host_alloc(cdp_ctx *ctx, size_t bytes) {
    if (!ctx) return nullptr;
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        ctx->err = "cudaMallocHost failed";
        return nullptr;
    }
    return p;
}
extern "C" void cdp_host_free(cdp_ctx *ctx, void *h_ptr) {
    if (!ctx || !h_ptr) return;
    cudaFreeHost(h_ptr);
}
extern "C" int cdp_h2d(cdp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    if (bytes == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CDP_OK;
}
extern "C" int cdp_host_is_pinned(const void *h_ptr) {
    if (!h_ptr) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h_ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}
extern "C" int cdp_h2d_2d(cdp_ctx *ctx, void *d_dst, size_t dpitch, const void *h_src, size_t spitch, size_t width, size_t height) {
    if (!ctx || !d_dst || !h_src) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_h2d_2d: null argument");
    if (width == 0 || height == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(d_dst, dpitch, h_src, spitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    return CDP_OK;
}
extern "C" int cdp_d2h(cdp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    if (bytes == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return CDP_OK;
}

static int msm_batch_dev_c(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, const cdp_msm_seg *d_segs, size_t count,
                           size_t max_n, size_t total_pairs, uint8_t *d_out_jac, int force_c);
extern "C" int cdp_msm_batch_dev(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, const cdp_msm_seg *d_segs,
                                 size_t count, size_t max_n, size_t total_pairs, uint8_t *d_out_jac) {
    return msm_batch_dev_c(ctx, d_affine_pts, d_scalars, d_segs, count, max_n, total_pairs, d_out_jac, 0);
}
static int msm_batch_dev_c(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, const cdp_msm_seg *d_segs, size_t count,
                           size_t max_n, size_t total_pairs, uint8_t *d_out_jac, int force_c) {
    if (!ctx || !d_out_jac) return CDP_ERR_INVALID_ARG;
    if (count == 0) return CDP_OK;
    if (max_n > SMALL_MSM_MAX_N) return fail(ctx, CDP_ERR_TOO_LARGE, "cdp_msm_batch_dev: segment longer than 2048 points; split it");
    if (max_n == 0) max_n = 1;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    msm_cfg g = pick_cfg(max_n);
    if (force_c) { g.c = force_c; g.nwin = msm_nwin_for(force_c); }
    TRY(ensure_dev(ctx, ctx->d_win, msm_bucket_sums_bytes(g.c, count)));
    static_assert(sizeof(cdp_msm_seg) == sizeof(msm_seg_t), "segment layout");
    TRY(msm_buckets_dev(ctx, g, d_affine_pts, d_scalars, reinterpret_cast<const msm_seg_t *>(d_segs), count, max_n,
                        reinterpret_cast<uint32_t *>(ctx->d_win.ptr), total_pairs));
    return combine_dev(ctx, g, reinterpret_cast<const uint32_t *>(ctx->d_win.ptr), count, reinterpret_cast<uint32_t *>(d_out_jac));
}

extern "C" int cdp_smul_jobs_dev(cdp_ctx *ctx, uint8_t *d_pts, const uint8_t *d_scalars, const cdp_smul_job *d_jobs, size_t n_jobs,
                                 size_t elems_per_job) {
    if (!ctx || !d_pts || !d_scalars || !d_jobs) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_smul_jobs_dev: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cdp_smul_job) == sizeof(smul_job_t), "job layout");
    return smul_jobs_dev(ctx, d_pts, d_scalars, reinterpret_cast<const smul_job_t *>(d_jobs), n_jobs, elems_per_job);
}

extern "C" int cdp_gather_dev(cdp_ctx *ctx, uint8_t *d_pts, const uint8_t *d_src, const uint32_t *d_src_idx, const uint32_t *d_dst_idx, size_t n) {
    if (!ctx || (n && (!d_pts || !d_src || !d_src_idx || !d_dst_idx))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_gather_dev: null argument");
    if (n == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, n);
    CUDA_TRY(ctx, launch_gather_points(ctx->stream, reinterpret_cast<uint32_t *>(d_pts), reinterpret_cast<const uint32_t *>(d_src), d_src_idx,
                                       d_dst_idx, (uint32_t)n));
    return CDP_OK;
}

extern "C" int cdp_compress_affine_dev(cdp_ctx *ctx, const uint8_t *d_pts, const uint32_t *d_index, size_t n, uint8_t *d_out_compressed) {
    if (!ctx || (n && (!d_pts || !d_out_compressed))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_compress_affine_dev: null argument");
    if (n == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, n);
    CUDA_TRY(ctx, launch_compress_affine(ctx->stream, reinterpret_cast<const uint32_t *>(d_pts), d_index, d_out_compressed, (uint32_t)n));
    return CDP_OK;
}

extern "C" int cdp_transcript_open_dev(cdp_ctx *ctx, const uint8_t *d_comp_vecs, const uint8_t *d_comp_M, size_t ell, size_t batch, uint8_t *d_vec_a,
                                       uint8_t *d_state) {
    if (!ctx || (batch && (!d_comp_vecs || !d_comp_M || !d_vec_a || !d_state)) || ell == 0) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_transcript_open_dev: bad argument");
    if (batch == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_TRANSCRIPT, batch);
    CUDA_TRY(ctx, launch_transcript_open(ctx->stream, d_comp_vecs, d_comp_M, (uint32_t)ell, (uint32_t)batch, d_vec_a, reinterpret_cast<uint64_t *>(d_state)));
    return CDP_OK;
}

extern "C" int cdp_verify_transcript_a_dev(cdp_ctx *ctx, const uint8_t *d_proof_points, const uint8_t *d_proof_scalars, const uint8_t *d_comp_vecs,
                                          const uint8_t *d_comp_M, const uint8_t *d_vec_a, size_t ell, size_t batch, uint8_t *d_state,
                                          uint8_t *d_challenges, uint8_t *d_tmp, uint8_t *d_stage_scalars, uint8_t *d_flags) {
    if (!ctx || ell == 0 || (batch && (!d_proof_points || !d_proof_scalars || !d_comp_vecs || !d_comp_M || !d_vec_a || !d_state || !d_challenges || !d_tmp ||
                                       !d_stage_scalars || !d_flags)))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_transcript_a_dev: bad argument");
    if (batch == 0) return CDP_OK;
    size_t n = ell + 4, m = 0;
    while ((size_t(1) << m) < n) m++;
    if ((size_t(1) << m) != n || m > 16) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_transcript_a_dev: ell + 4 must be a power of two");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_TRANSCRIPT, batch);
    CUDA_TRY(ctx, launch_verify_transcript_a(ctx->stream, d_proof_points, d_proof_scalars, d_comp_M, d_vec_a, (uint32_t)ell, (uint32_t)(18 + 10 * m),
                                             (uint32_t)(27 + 4 * m), (uint32_t)batch, reinterpret_cast<uint64_t *>(d_state),
                                             reinterpret_cast<uint32_t *>(d_challenges), reinterpret_cast<uint32_t *>(d_tmp),
                                             reinterpret_cast<uint32_t *>(d_stage_scalars), d_comp_vecs, d_flags));
    return CDP_OK;
}
extern "C" int cdp_verify_transcript_b_dev(cdp_ctx *ctx, const uint8_t *d_proof_points, const uint8_t *d_proof_scalars, const uint8_t *d_comp_vecs,
                                          const uint8_t *d_comp_DA, const uint8_t *d_comp_H, size_t ell, size_t batch, uint8_t *d_state,
                                          uint8_t *d_challenges, const uint8_t *d_tmp) {
    if (!ctx || ell == 0 || (batch && (!d_proof_points || !d_proof_scalars || !d_comp_vecs || !d_comp_DA || !d_comp_H || !d_state || !d_challenges || !d_tmp)))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_transcript_b_dev: bad argument");
    if (batch == 0) return CDP_OK;
    size_t n = ell + 4, m = 0;
    while ((size_t(1) << m) < n) m++;
    if ((size_t(1) << m) != n || m > 16) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_transcript_b_dev: ell + 4 must be a power of two");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_TRANSCRIPT, batch);
    CUDA_TRY(ctx, launch_verify_transcript_b(ctx->stream, d_proof_points, d_proof_scalars, d_comp_DA, d_comp_vecs, d_comp_H, (uint32_t)ell, (uint32_t)m,
                                             (uint32_t)(18 + 10 * m), (uint32_t)(27 + 4 * m), (uint32_t)batch, reinterpret_cast<uint64_t *>(d_state),
                                             reinterpret_cast<uint32_t *>(d_challenges), reinterpret_cast<const uint32_t *>(d_tmp)));
    return CDP_OK;
}

extern "C" int cdp_verify_coeffs_dev(cdp_ctx *ctx, const uint8_t *d_challenges, const uint8_t *d_vec_a, const cdp_vcoef_params *params, size_t batch,
                                     uint8_t *d_crs_scalars, uint8_t *d_var_scalars, uint8_t *d_exact_scalars) {
    if (!ctx || !params || (batch && (!d_challenges || !d_vec_a || !d_crs_scalars || !d_var_scalars || !d_exact_scalars)))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_coeffs_dev: bad argument");
    if (batch == 0) return CDP_OK;
    const cdp_vcoef_params &q = *params;
    if (q.m == 0 || q.m > 16 || q.n != (1u << q.m) || q.ell + 4 != q.n || q.vch != 27 + 4 * q.m || q.vw < q.big_n - (q.n + 5))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_verify_coeffs_dev: inconsistent layout");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, batch);
    vcoef_params_t P = {q.ell, q.n, q.m, q.big_n, q.vw, q.o_R, q.o_S, q.o_T, q.o_U, q.o_M, q.o_P, q.exact_eq, q.vch};
    CUDA_TRY(ctx, launch_verify_coeffs(ctx->stream, reinterpret_cast<const uint32_t *>(d_challenges), reinterpret_cast<const uint32_t *>(d_vec_a), P,
                                       (uint32_t)batch, reinterpret_cast<uint32_t *>(d_crs_scalars), reinterpret_cast<uint32_t *>(d_var_scalars),
                                       reinterpret_cast<uint32_t *>(d_exact_scalars)));
    return CDP_OK;
}

extern "C" int cdp_round_expand_dev(cdp_ctx *ctx, const uint8_t *d_compact, const uint8_t *d_u_canonical, size_t n, size_t h, size_t scalars_per_proof,
                                    size_t compact_per_proof, int mode, size_t batch, uint8_t *d_scalars_out) {
    if (!ctx || (batch && (!d_compact || !d_scalars_out)) || (mode != 0 && mode != 1) || (mode == 0 && batch && !d_u_canonical) || h == 0 ||
        (h & (h - 1)) || n < 2 * h || n % (2 * h))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_round_expand_dev: bad argument");
    const size_t Q = n / (2 * h);
    if (compact_per_proof < (mode == 0 ? 2 * Q + 4 * h + 2 : Q + 2 * h) || scalars_per_proof < (mode == 0 ? 2 * n + 2 : n + 2 * h))
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_round_expand_dev: inconsistent layout");
    if (batch == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, batch * n);
    round_expand_params_t P = {(uint32_t)n, (uint32_t)h, (uint32_t)scalars_per_proof, (uint32_t)compact_per_proof, (uint32_t)mode};
    CUDA_TRY(ctx, launch_round_expand(ctx->stream, reinterpret_cast<const uint32_t *>(d_compact), reinterpret_cast<const uint32_t *>(d_u_canonical), P,
                                      (uint32_t)batch, reinterpret_cast<uint32_t *>(d_scalars_out)));
    return CDP_OK;
}

extern "C" size_t cdp_prove_work_scalars(size_t ell) { return 11 * (ell + 4) + 64; }
extern "C" size_t cdp_prove_random_scalars(size_t ell) { return 3 * (ell + 4) + 11; }
extern "C" int cdp_prove_random_dev(cdp_ctx *ctx, const uint8_t *d_keys, const uint64_t *d_skip_words, size_t batch, size_t ell, uint8_t *d_random) {
    if (!ctx || !d_keys || !d_random || ell < 4) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_prove_random_dev: bad argument");
    if (batch == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_PROVE_STAGE, batch);
    CUDA_TRY(ctx, launch_prove_random(ctx->stream, reinterpret_cast<const uint32_t *>(d_keys), d_skip_words, (uint32_t)ell, (uint32_t)batch,
                                      reinterpret_cast<uint32_t *>(d_random)));
    return CDP_OK;
}
extern "C" int cdp_prove_stage_dev(cdp_ctx *ctx, const cdp_prove_dev *P, int stage, unsigned round) {
    if (!ctx || !P || stage < CDP_PS_S1 || stage > CDP_PS_SM_ROUND) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_prove_stage_dev: bad argument");
    if (P->batch == 0) return CDP_OK;
    const size_t n = (size_t)P->ell + 4;
    if (P->ell < 4 || ((size_t)1 << P->m) != n || round >= P->m || P->switch_round < 1 || P->switch_round > P->m || !P->d_state || !P->d_vec_a || !P->d_perm || !P->d_witness || !P->d_random ||
        !P->d_work || !P->d_comp0_vecs || !P->d_comp0_M || !P->d_comp_H || !P->d_comp || !P->d_side || !P->d_proofs || !P->d_scalars ||
        !P->d_fold_scalars)
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_prove_stage_dev: inconsistent parameters");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_PROVE_STAGE, P->batch);
    CUDA_TRY(ctx, launch_prove_stage(ctx->stream, *P, stage, round));
    return CDP_OK;
}

extern "C" int cdp_sum_scalars_dev(cdp_ctx *ctx, const uint8_t *d_scalars, size_t row_stride, size_t cols, size_t rows, uint8_t *d_out) {
    if (!ctx || (cols && rows && (!d_scalars || !d_out)) || row_stride < cols) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_sum_scalars_dev: bad argument");
    if (cols == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, cols * rows);
    CUDA_TRY(ctx, launch_sum_scalars(ctx->stream, reinterpret_cast<const uint32_t *>(d_scalars), (uint32_t)row_stride, (uint32_t)cols, (uint32_t)rows,
                                     reinterpret_cast<uint32_t *>(d_out)));
    return CDP_OK;
}

extern "C" int cdp_dev_zero(cdp_ctx *ctx, void *d_ptr, size_t bytes) {
    if (!ctx || (bytes && !d_ptr)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_dev_zero: bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemsetAsync(d_ptr, 0, bytes, ctx->stream));
    return CDP_OK;
}

extern "C" int cdp_decompress_dev(cdp_ctx *ctx, const uint8_t *d_compressed, const uint32_t *d_dst_index, size_t n, uint8_t *d_out_affine,
                                  uint8_t *d_status) {
    if (!ctx || (n && (!d_compressed || !d_out_affine || !d_status))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_decompress_dev: null argument");
    if (n == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, n);
    CUDA_TRY(ctx, launch_decompress(ctx->stream, d_compressed, d_dst_index, reinterpret_cast<uint32_t *>(d_out_affine), d_status, (uint32_t)n));
    return CDP_OK;
}

extern "C" int cdp_decompress_batch(cdp_ctx *ctx, const uint8_t *compressed, size_t n, uint8_t *out_affine, uint8_t *status) {
    if (!ctx || (n && (!compressed || !out_affine || !status))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_decompress_batch: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return CDP_OK;
    TRY(ensure_dev(ctx, ctx->d_aux, n * CDP_COMPRESSED_BYTES));
    TRY(ensure_dev(ctx, ctx->d_out, n * CDP_AFFINE_BYTES));
    TRY(ensure_dev(ctx, ctx->d_segs, n));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_aux.ptr, compressed, n * CDP_COMPRESSED_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    TRY(cdp_decompress_dev(ctx, (const uint8_t *)ctx->d_aux.ptr, nullptr, n, (uint8_t *)ctx->d_out.ptr, (uint8_t *)ctx->d_segs.ptr));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_affine, ctx->d_out.ptr, n * CDP_AFFINE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(status, ctx->d_segs.ptr, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int bad = 0;
    for (size_t i = 0; i < n; i++) bad |= status[i];
    return bad ? fail(ctx, CDP_ERR_NOT_ON_CURVE, "cdp_decompress_batch: at least one encoding is invalid (see status[])") : CDP_OK;
}

extern "C" int cdp_normalize_dev(cdp_ctx *ctx, const uint8_t *d_jac, size_t n, uint8_t *d_out_affine, uint8_t *d_out_compressed) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return normalize_dev(ctx, d_jac, n, d_out_affine, d_out_compressed);
}

// =================================================================================================== host-buffer drop-ins
// One MSM of any size: chunks of <= 2048 points (one CTA pair each), window sums reduced across chunks, one Horner pass.
static int msm_single_resident(cdp_ctx *ctx, const uint8_t *d_pts, const uint8_t *d_scalars, size_t n, uint8_t *d_out_jac) {
    // A lone MSM is latency-bound: a lane of the bucket kernel adds its ~m / 4 points (chunk of m points, c = 5) one after the other, ~8 us
    // per mixed addition for a warp running alone, and only 7 warps work on a chunk.  So a single MSM is cut into chunks of <= 64 points
    // (16 sequential additions per lane, n / 64 * 7 warps in flight); their bucket sums are added slot by slot and combined once -- the
    // ~130 dependent doublings of that one Horner pass (~0.7 ms) are what is left.  CDP_MSM_SINGLE_CHUNK overrides the chunk size.
    static const size_t single_chunk = [] { const char *e = getenv("CDP_MSM_SINGLE_CHUNK"); size_t v = e ? (size_t)atoll(e) : 64; return std::min<size_t>(SMALL_MSM_MAX_N, std::max<size_t>(v, 16)); }();
    size_t chunk = std::min<size_t>(single_chunk, n);
    size_t nchunks = (n + chunk - 1) / chunk;
    msm_cfg g = pick_cfg(chunk);
    std::vector<msm_seg_t> segs(nchunks);
    for (size_t i = 0; i < nchunks; i++) {
        segs[i].pts_off = (uint32_t)(i * chunk);
        segs[i].scalars_off = (uint32_t)(i * chunk);
        segs[i].n = (uint32_t)std::min(chunk, n - i * chunk);
        segs[i].extra = 0;
    }
    TRY(ensure_dev(ctx, ctx->d_segs, nchunks * sizeof(msm_seg_t)));
    TRY(ensure_host(ctx, ctx->h_stage, nchunks * sizeof(msm_seg_t)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // staging buffer reuse
    memcpy(ctx->h_stage.ptr, segs.data(), nchunks * sizeof(msm_seg_t));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_segs.ptr, ctx->h_stage.ptr, nchunks * sizeof(msm_seg_t), cudaMemcpyHostToDevice, ctx->stream));
    TRY(ensure_dev(ctx, ctx->d_win, msm_bucket_sums_bytes(g.c, nchunks + 1)));
    uint32_t *win = reinterpret_cast<uint32_t *>(ctx->d_win.ptr);
    TRY(msm_buckets_dev(ctx, g, d_pts, d_scalars, reinterpret_cast<const msm_seg_t *>(ctx->d_segs.ptr), nchunks, chunk, win, n));
    const uint32_t *win_final = win;
    if (nchunks > 1) {  // bucket sums of the chunks are added slot by slot, then combined once
        const uint32_t slots = (uint32_t)(g.nwin * msm_nb_for(g.c));
        uint32_t *red = win + 36 * nchunks * (size_t)slots;
        launch_scope ls(ctx, CDP_PROFILE_OTHER, nchunks);
        CUDA_TRY(ctx, launch_sum_groups(ctx->stream, win, red, slots, (uint32_t)nchunks, slots));
        win_final = red;
    }
    return combine_dev(ctx, g, win_final, 1, reinterpret_cast<uint32_t *>(d_out_jac));
}

// ---- large Pippenger (k_bigmsm.cu) ---------------------------------------------------------------------------
// below this size (2^16 pairs) the chunked small-MSM path (more additions per pair, but no sort and a much shorter launch chain) is faster;
// CDP_BIG_MIN_LOG2 overrides (tuning)
static const size_t BIG_MSM_MIN_N = [] { const char *e = getenv("CDP_BIG_MIN_LOG2"); int v = e ? atoi(e) : 16; return size_t(1) << (v >= 11 && v <= 24 ? v : 16); }();
static int big_c_for(size_t n) {
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("CDP_BIG_C"); forced = e ? atoi(e) : 0; }
    if (forced >= 12 && forced <= 20) return forced;
    // measured with load-ordered slots and quad reductions (tools/msm_latency.py, round 2): 2^16..2^19: 15, 2^20: 16, 2^21: 17, 2^22: 19 (7 windows)
    return n < (size_t(1) << 14) ? 12 : n < (size_t(1) << 16) ? 13 : n < (size_t(1) << 20) ? 15 : n < (size_t(1) << 21) ? 16 : n < (size_t(1) << 22) ? 17 : 19;
}
// The bucket reduction of the large Pippenger, shared by both accumulate paths: 16-ary running-sum levels while a window still has many nodes
// (throughput-bound), then -- from BITSUM_MAX nodes per window down -- the window totals as bit-decomposed plain sums and the final Horner
// (k_big_bitsums / k_big_horner2: ~45 dependent quad operations instead of ~50 per remaining level).  CDP_BIG_BITSUM=0 keeps the levels.
static int big_reduce_tail(cdp_ctx *ctx, const uint32_t *buckets, bool affine, uint32_t nb, int nwin, int c, uint8_t *ws, size_t o_A0, size_t o_B0,
                           size_t o_A1, size_t o_B1, uint8_t *d_out_jac) {
    static const uint32_t BITSUM_MAX = [] { const char *e = getenv("CDP_BIG_BITSUM"); int v = e ? atoi(e) : 1024; return (uint32_t)(v < 0 ? 0 : v); }();
    const uint32_t *Ain = nullptr, *Bin = nullptr;
    uint32_t *Aout = (uint32_t *)(ws + o_A0), *Bout = (uint32_t *)(ws + o_B0);
    uint32_t len = nb;
    int shift = 0, flip = 0;
    while (len > 1) {
        if (Ain && len <= BITSUM_MAX && nwin <= 16) {
            launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, nwin);
            CUDA_TRY(ctx, launch_big_bitsum_horner(ctx->stream, Ain, Bin, len, shift, nwin, c, Aout, reinterpret_cast<uint32_t *>(d_out_jac)));
            ctx->launches++;
            return CDP_OK;
        }
        // first level: LEAF_G buckets per thread (tuning: CDP_BIG_LEAF_G), 16 per node above
        static const uint32_t LEAF_G = [] { const char *e = getenv("CDP_BIG_LEAF_G"); int v = e ? atoi(e) : 16; return (uint32_t)(v == 8 || v == 32 || v == 64 ? v : 16); }();
        uint32_t g = std::min<uint32_t>(Ain ? 16 : LEAF_G, len);
        uint32_t n_out = (uint32_t)((size_t)nwin * (len / g));
        launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, n_out);
        if (!Ain && affine) CUDA_TRY(ctx, launch_big_reduce_leaf_affine(ctx->stream, buckets, n_out, g, Aout, Bout));
        else CUDA_TRY(ctx, launch_big_reduce_level(ctx->stream, Ain ? Ain : buckets, Bin, n_out, g, shift, Aout, Bout));
        Ain = Aout; Bin = Bout;
        flip ^= 1;
        Aout = (uint32_t *)(ws + (flip ? o_A1 : o_A0)); Bout = (uint32_t *)(ws + (flip ? o_B1 : o_B0));
        len /= g;
        while (g > 1) { shift++; g >>= 1; }
    }
    { launch_scope ls(ctx, CDP_PROFILE_MSM_COMBINE, nwin); CUDA_TRY(ctx, launch_big_horner(ctx->stream, Ain, Bin, nwin, c, reinterpret_cast<uint32_t *>(d_out_jac))); }
    return CDP_OK;
}
// Bucket sums by rounds of batched affine additions (k_batchaff.cu): the default; CDP_BIG_BA=0 keeps the one-thread-per-bucket XYZZ accumulate.
static const bool BIG_BA = [] { const char *e = getenv("CDP_BIG_BA"); return !e || atoi(e) != 0; }();
// from 2^19 pairs: below, the rounds' fixed costs (a scan, a job list and an inversion's latency per round, ~10 rounds) outweigh the cheaper
// additions.  At 2^19 a lone call takes the same time either way (5.7 ms), but the rounds are 40 % fewer multiply-adds: eight verifier lanes that
// each run a 2^19.1-pair merged check side by side go from 50 900 to 55 400 verifies/s
static const size_t BIG_BA_MIN_N = [] { const char *e = getenv("CDP_BIG_BA_MIN_LOG2"); int v = e ? atoi(e) : 19; return size_t(1) << (v >= 11 && v <= 30 ? v : 19); }();
static uint32_t env_u32(const char *name, uint32_t dflt) { const char *e = getenv(name); return e && atoi(e) > 0 ? (uint32_t)atoi(e) : dflt; }
static int msm_big_resident_ba(cdp_ctx *ctx, const uint8_t *d_pts, const uint8_t *d_scalars, size_t n, uint8_t *d_out_jac) {
    const int c = big_c_for(n), nwin = (130 + c - 1) / c;
    const uint32_t nb = 1u << (c - 1), n2 = (uint32_t)(2 * n);
    const size_t slots = (size_t)nwin * nb, items = (size_t)nwin * n2;
    const size_t sort_tmp = big_msm_sort_temp_bytes(n2, nwin, c), scan_tmp = ba_scan_temp_bytes(slots);
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t lvl = slots / 2 + nwin;
    const size_t s0n = (items + slots) / 2 + 64, s1n = (s0n + slots) / 2 + 64;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += al(bytes); return at; };
    const size_t o_keys = take(items * 4), o_vals = take(items * 4), o_keys2 = take(items * 4), o_vals2 = take(items * 4),
                 o_start = take((size_t)nwin * (nb + 1) * 4), o_bx = take(n * 48), o_baff = take(slots * 96), o_act0 = take(3 * slots * 4),
                 o_act1 = take(3 * slots * 4), o_sc0 = take(slots * 16), o_sc1 = take(slots * 16), o_incl = take(slots * 16),
                 o_A0 = take(lvl * 144), o_B0 = take(lvl * 144), o_A1 = take(lvl * 144), o_B1 = take(lvl * 144), o_misc = take(BA_STATS_ROUNDS * 3 * 4);
    // the sort's temporary storage is dead once the ids are sorted: the round buffers start there
    const size_t o_tmp = o;
    static const bool BA_STASH = [] { const char *e = getenv("CDP_BA_STASH"); return e && atoi(e) != 0; }();
    const size_t o_scan_tmp = take(scan_tmp), o_jobs = take((items / 2 + 1) * 16), o_s0 = take(s0n * 96), o_s1 = take(s1n * 96),
                 o_stash = take(BA_STASH ? (items / 2 + 1) * 192 : 0);
    const size_t total = std::max(o, o_tmp + al(sort_tmp));
    TRY(ensure_dev(ctx, ctx->d_big, total));
    uint8_t *ws = (uint8_t *)ctx->d_big.ptr;
    uint32_t *keys = (uint32_t *)(ws + o_keys), *vals = (uint32_t *)(ws + o_vals), *keys2 = (uint32_t *)(ws + o_keys2), *vals2 = (uint32_t *)(ws + o_vals2);
    uint32_t *start = (uint32_t *)(ws + o_start), *bx = (uint32_t *)(ws + o_bx), *baff = (uint32_t *)(ws + o_baff);
    const uint32_t *P = reinterpret_cast<const uint32_t *>(d_pts), *S = reinterpret_cast<const uint32_t *>(d_scalars);
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, n); CUDA_TRY(ctx, launch_big_digits(ctx->stream, P, S, (uint32_t)n, c, nwin, keys, vals, bx)); }
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, items); CUDA_TRY(ctx, launch_big_sort(ctx->stream, ws + o_tmp, sort_tmp, keys, keys2, vals, vals2, n2, nwin, c, nullptr)); }
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, items); CUDA_TRY(ctx, launch_big_offsets(ctx->stream, keys2, n2, nwin, nb, c, start)); }
    uint32_t *act[2] = {(uint32_t *)(ws + o_act0), (uint32_t *)(ws + o_act1)};
    void *sc[2] = {ws + o_sc0, ws + o_sc1};
    uint32_t *stats = (uint32_t *)(ws + o_misc);
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, slots); CUDA_TRY(ctx, launch_ba_init(ctx->stream, start, n2, nwin, nb, P, bx, vals2, act[0], sc[0], baff, stats)); }
    // (additions, elements kept, list length) of every round, exact: one read-back sizes all later launches
    uint32_t h_stats[BA_STATS_ROUNDS * 3];
    CUDA_TRY(ctx, cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // K additions share an inversion (~5 additions' worth of instructions).  All threads of a round do the same work, so the grid runs in
    // waves of WAVE resident threads (4 CTAs of 128 per SM at 128 registers): K is the smallest that packs the round into w FULL waves, w the fewest waves with K <= K_MAX
    static const uint32_t K_MAX = env_u32("CDP_BA_KMAX", 128), WAVE = env_u32("CDP_BA_WAVE", 148 * 4 * 128);
    uint32_t *sbuf[2] = {(uint32_t *)(ws + o_s0), (uint32_t *)(ws + o_s1)};
    int total_rounds = 0;
    while (total_rounds < BA_STATS_ROUNDS && h_stats[3 * total_rounds]) total_rounds++;
    // the tail of a deep tree (short list, at most 2^FIN_ROUNDS elements per bucket left; 0 = never): one thread per bucket instead of more rounds
    static const uint32_t FIN_LIST = env_u32("CDP_BA_FINISH_LIST", 32768), FIN_ROUNDS = getenv("CDP_BA_FINISH_ROUNDS") ? (uint32_t)atoi(getenv("CDP_BA_FINISH_ROUNDS")) : 5u;
    for (int r = 0; r < total_rounds; r++) {
        const uint32_t pairs = h_stats[3 * r], list_len = r == 0 ? (uint32_t)slots : h_stats[3 * r + 2];
        if (list_len <= FIN_LIST && (uint32_t)(total_rounds - r) <= FIN_ROUNDS && total_rounds - r >= 2) {
            launch_scope ls(ctx, CDP_PROFILE_MSM_BUCKETS, 0);
            CUDA_TRY(ctx, launch_ba_finish(ctx->stream, r == 0, act[r & 1], slots, list_len, P, bx, vals2, sbuf[(r + 1) & 1], baff));
            break;
        }
        if (h_stats[3 * r + 1] > ((r & 1) ? s1n : s0n)) return fail(ctx, CDP_ERR_TOO_LARGE, "cdp_msm: round buffer too small");  // cannot happen: sized for the worst case
        const uint32_t waves = std::max(1u, (uint32_t)(((uint64_t)pairs + (uint64_t)K_MAX * WAVE - 1) / ((uint64_t)K_MAX * WAVE)));
        const uint32_t K = std::max(1u, (uint32_t)(((uint64_t)pairs + (uint64_t)waves * WAVE - 1) / ((uint64_t)waves * WAVE)));
        launch_scope ls(ctx, CDP_PROFILE_MSM_BUCKETS, r == 0 ? (uint64_t)n : 0);
        CUDA_TRY(ctx, launch_ba_round(ctx->stream, r == 0, ws + o_scan_tmp, scan_tmp, sc[r & 1], ws + o_incl, list_len, pairs, K, act[r & 1], act[(r + 1) & 1],
                                      slots, sc[(r + 1) & 1], P, bx, vals2, sbuf[(r + 1) & 1], sbuf[r & 1], baff, ws + o_jobs,
                                      BA_STASH && r == 0 ? (uint32_t *)(ws + o_stash) : nullptr));
        ctx->launches += 3;
    }
    // hierarchical reduction: sum_b (b+1) B_b per window (the leaves are affine), Horner over the windows
    return big_reduce_tail(ctx, baff, true, nb, nwin, c, ws, o_A0, o_B0, o_A1, o_B1, d_out_jac);
}
static int msm_big_resident(cdp_ctx *ctx, const uint8_t *d_pts, const uint8_t *d_scalars, size_t n, uint8_t *d_out_jac) {
    if (BIG_BA && n >= (ctx->big_ba_min ? ctx->big_ba_min : BIG_BA_MIN_N) && (size_t)((130 + big_c_for(n) - 1) / big_c_for(n)) * 2 * n < (size_t(1) << 30)) return msm_big_resident_ba(ctx, d_pts, d_scalars, n, d_out_jac);
    const int c = big_c_for(n), nwin = (130 + c - 1) / c;
    const uint32_t nb = 1u << (c - 1), n2 = (uint32_t)(2 * n);
    const int top_bits = 128 - c * (nwin - 1);
    const uint32_t nbt = top_bits > 0 ? (1u << top_bits) : 1u;   // upper bound of the top window's digits (incl. the carry)
    const uint32_t sp_top = nb / nbt;                             // threads per top-window bucket
    const size_t slots = (size_t)nwin * nb, items = (size_t)nwin * n2;
    const size_t sort_tmp = big_msm_sort_temp_bytes(n2, nwin, c);
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    // workspace: keys/values (double-buffered for the sort), bucket offsets, buckets, two ping-pong (A, Bv) level buffers, sort temp
    const size_t lvl = slots / 2 + nwin;  // first level output is at most slots / 2 nodes
    size_t o_keys = 0, o_vals = o_keys + al(items * 4), o_keys2 = o_vals + al(items * 4), o_vals2 = o_keys2 + al(items * 4),
           o_start = o_vals2 + al(items * 4), o_bjac = o_start + al((size_t)nwin * (nb + 1) * 4), o_top = o_bjac + al(slots * 144),
           o_A0 = o_top + al((size_t)nb * 144), o_B0 = o_A0 + al(lvl * 144), o_A1 = o_B0 + al(lvl * 144), o_B1 = o_A1 + al(lvl * 144),
           o_tmp = o_B1 + al(lvl * 144), o_heavy = o_tmp + al(sort_tmp), o_order = o_heavy + al(big_heavy_bytes(n2, nwin)),
           total = o_order + al(4 * slots * 4 + big_order_temp_bytes(slots));
    TRY(ensure_dev(ctx, ctx->d_big, total));
    uint8_t *ws = (uint8_t *)ctx->d_big.ptr;
    uint32_t *keys = (uint32_t *)(ws + o_keys), *vals = (uint32_t *)(ws + o_vals), *keys2 = (uint32_t *)(ws + o_keys2), *vals2 = (uint32_t *)(ws + o_vals2);
    uint32_t *start = (uint32_t *)(ws + o_start), *bjac = (uint32_t *)(ws + o_bjac);
    const uint32_t *P = reinterpret_cast<const uint32_t *>(d_pts), *S = reinterpret_cast<const uint32_t *>(d_scalars);
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, n); CUDA_TRY(ctx, launch_big_digits(ctx->stream, P, S, (uint32_t)n, c, nwin, keys, vals, nullptr)); }
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, items); CUDA_TRY(ctx, launch_big_sort(ctx->stream, ws + o_tmp, sort_tmp, keys, keys2, vals, vals2, n2, nwin, c, nullptr)); }
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, items); CUDA_TRY(ctx, launch_big_offsets(ctx->stream, keys2, n2, nwin, nb, c, start)); }
    uint32_t *order_ws = (uint32_t *)(ws + o_order);
    { launch_scope ls(ctx, CDP_PROFILE_OTHER, slots); CUDA_TRY(ctx, launch_big_order(ctx->stream, start, nwin, nb, sp_top, order_ws, big_order_temp_bytes(slots))); }
    { launch_scope ls(ctx, CDP_PROFILE_MSM_BUCKETS, (uint64_t)n); CUDA_TRY(ctx, launch_big_accumulate(ctx->stream, P, vals2, start, n2, nwin, nb, sp_top, order_ws + 3 * slots, bjac, ws + o_heavy)); }
    // top window: fold the sp_top partial sums of each bucket (two steps when a bucket has many), back into the window's slot array
    {
        uint32_t *top_slots = bjac + 36 * (size_t)(nwin - 1) * nb, *tmp = (uint32_t *)(ws + o_top);
        uint32_t nw = sp_top > 1024 ? 32u : 1u;
        if (nw > 1) {
            launch_scope ls(ctx, CDP_PROFILE_OTHER, nbt * nw);
            CUDA_TRY(ctx, launch_big_fold_top(ctx->stream, top_slots, nbt, sp_top, nw, 0, tmp));
            CUDA_TRY(ctx, cudaMemcpyAsync(top_slots, tmp, (size_t)nbt * nw * 144, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        launch_scope ls(ctx, CDP_PROFILE_OTHER, nb);
        CUDA_TRY(ctx, launch_big_fold_top(ctx->stream, top_slots, nbt, nw > 1 ? nw : sp_top, 1, nb, tmp));
        CUDA_TRY(ctx, cudaMemcpyAsync(top_slots, tmp, (size_t)nb * 144, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // hierarchical reduction: sum_b (b+1) B_b per window, Horner over the windows
    return big_reduce_tail(ctx, bjac, false, nb, nwin, c, ws, o_A0, o_B0, o_A1, o_B1, d_out_jac);
}

extern "C" int cdp_sum_jacobian_dev(cdp_ctx *ctx, const uint8_t *d_jac_in, size_t count, uint8_t *d_out_jac) {
    if (!ctx || !d_jac_in || !d_out_jac || count == 0) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_sum_jacobian_dev: bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, count);
    CUDA_TRY(ctx, launch_sum_groups(ctx->stream, reinterpret_cast<const uint32_t *>(d_jac_in), reinterpret_cast<uint32_t *>(d_out_jac), 1,
                                    (uint32_t)count, 1));
    return CDP_OK;
}

extern "C" int cdp_sum_groups_dev(cdp_ctx *ctx, const uint8_t *d_jac_in, size_t n_out, size_t per_out, size_t group_stride, uint8_t *d_out_jac) {
    if (!ctx || !d_jac_in || !d_out_jac || per_out == 0) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_sum_groups_dev: bad argument");
    if (n_out == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, n_out * per_out);
    CUDA_TRY(ctx, launch_sum_groups(ctx->stream, reinterpret_cast<const uint32_t *>(d_jac_in), reinterpret_cast<uint32_t *>(d_out_jac), (uint32_t)n_out,
                                    (uint32_t)per_out, (uint32_t)group_stride));
    return CDP_OK;
}

extern "C" int cdp_sum_groups2_dev(cdp_ctx *ctx, const uint8_t *d_a, size_t per_a, size_t sa, size_t ga, const uint8_t *d_b, size_t per_b, size_t sb,
                                   size_t gb, size_t n_out, uint8_t *d_out_jac) {
    if (!ctx || !d_out_jac || (per_a && !d_a) || (per_b && !d_b) || per_a + per_b == 0) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_sum_groups2_dev: bad argument");
    if (n_out == 0) return CDP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    launch_scope ls(ctx, CDP_PROFILE_OTHER, n_out * (per_a + per_b));
    CUDA_TRY(ctx, launch_sum_groups2(ctx->stream, reinterpret_cast<const uint32_t *>(d_a), (uint32_t)per_a, (uint32_t)sa, (uint32_t)ga,
                                     reinterpret_cast<const uint32_t *>(d_b), (uint32_t)per_b, (uint32_t)sb, (uint32_t)gb,
                                     reinterpret_cast<uint32_t *>(d_out_jac), (uint32_t)n_out));
    return CDP_OK;
}

extern "C" int cdp_set_big_msm_min(cdp_ctx *ctx, size_t n_pairs) {
    if (!ctx || (n_pairs && n_pairs < 2048)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_set_big_msm_min: the sort-based path needs at least 2048 pairs");
    ctx->big_msm_min = n_pairs;
    return CDP_OK;
}
extern "C" int cdp_set_big_ba_min(cdp_ctx *ctx, size_t n_pairs) {
    if (!ctx || (n_pairs && n_pairs < 2048)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_set_big_ba_min: at least 2048 pairs");
    ctx->big_ba_min = n_pairs;
    return CDP_OK;
}
extern "C" int cdp_msm_dev(cdp_ctx *ctx, const uint8_t *d_affine_pts, const uint8_t *d_scalars, size_t n, uint8_t *d_out_jac) {
    if (!ctx || !d_out_jac || (n && (!d_affine_pts || !d_scalars))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_dev: null argument");
    if (n == 0 || n >= (size_t(1) << 31)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_dev: n out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n >= (ctx->big_msm_min ? ctx->big_msm_min : BIG_MSM_MIN_N)) return msm_big_resident(ctx, d_affine_pts, d_scalars, n, d_out_jac);
    return msm_single_resident(ctx, d_affine_pts, d_scalars, n, d_out_jac);
}

extern "C" int cdp_msm(cdp_ctx *ctx, const uint8_t *affine_pts, const uint8_t *scalars, size_t n, uint8_t out_jac[CDP_JACOBIAN_BYTES]) {
    if (!ctx || !out_jac || (n && (!affine_pts || !scalars))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm: null argument");
    if (n >= (size_t(1) << 31)) return fail(ctx, CDP_ERR_TOO_LARGE, "cdp_msm: n too large");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) {  // empty sum = infinity (Z = 0)
        memcpy(out_jac, INF_JAC_ZERO, CDP_JACOBIAN_BYTES);
        return CDP_OK;
    }
    TRY(ensure_dev(ctx, ctx->d_pts, n * CDP_AFFINE_BYTES));
    TRY(ensure_dev(ctx, ctx->d_scalars, n * CDP_SCALAR_BYTES));
    TRY(ensure_dev(ctx, ctx->d_out, CDP_JACOBIAN_BYTES));
    TRY(h2d_pipelined(ctx, ctx->d_pts.ptr, affine_pts, n * CDP_AFFINE_BYTES));
    TRY(h2d_pipelined(ctx, ctx->d_scalars.ptr, scalars, n * CDP_SCALAR_BYTES));
    if (n >= BIG_MSM_MIN_N) TRY(msm_big_resident(ctx, (const uint8_t *)ctx->d_pts.ptr, (const uint8_t *)ctx->d_scalars.ptr, n, (uint8_t *)ctx->d_out.ptr));
    else TRY(msm_single_resident(ctx, (const uint8_t *)ctx->d_pts.ptr, (const uint8_t *)ctx->d_scalars.ptr, n, (uint8_t *)ctx->d_out.ptr));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_jac, ctx->d_out.ptr, CDP_JACOBIAN_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}

extern "C" int cdp_msm_from_projective(cdp_ctx *ctx, const uint8_t *jac_pts, const uint8_t *scalars, size_t n,
                                       uint8_t out_jac[CDP_JACOBIAN_BYTES]) {
    if (!ctx || !out_jac || (n && (!jac_pts || !scalars))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_from_projective: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) {
        memcpy(out_jac, INF_JAC_ZERO, CDP_JACOBIAN_BYTES);
        return CDP_OK;
    }
    TRY(ensure_dev(ctx, ctx->d_jac, n * CDP_JACOBIAN_BYTES));
    TRY(ensure_dev(ctx, ctx->d_pts, n * CDP_AFFINE_BYTES));
    TRY(ensure_dev(ctx, ctx->d_scalars, n * CDP_SCALAR_BYTES));
    TRY(ensure_dev(ctx, ctx->d_out, CDP_JACOBIAN_BYTES));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_jac.ptr, jac_pts, n * CDP_JACOBIAN_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scalars.ptr, scalars, n * CDP_SCALAR_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    TRY(normalize_dev(ctx, (const uint8_t *)ctx->d_jac.ptr, n, (uint8_t *)ctx->d_pts.ptr, nullptr));
    TRY(msm_single_resident(ctx, (const uint8_t *)ctx->d_pts.ptr, (const uint8_t *)ctx->d_scalars.ptr, n, (uint8_t *)ctx->d_out.ptr));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_jac, ctx->d_out.ptr, CDP_JACOBIAN_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}

extern "C" int cdp_msm_batch(cdp_ctx *ctx, const cdp_msm_desc *descs, size_t count, uint8_t *out_jac) {
    if (!ctx || (count && (!descs || !out_jac))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_batch: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return CDP_OK;
    // pack all bases / scalars; MSMs are grouped into launches by window-size class
    size_t total = 0;
    for (size_t i = 0; i < count; i++) {
        if (descs[i].n && (!descs[i].affine_pts || !descs[i].scalars)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_batch: null segment");
        total += descs[i].n;
    }
    if (total >= (size_t(1) << 31)) return fail(ctx, CDP_ERR_TOO_LARGE, "cdp_msm_batch: too many points");
    // segments longer than the CTA capacity go through the single-MSM path
    std::vector<size_t> order[5], big;
    auto cls = [](size_t n) { return 6 - pick_cfg(n).c; };
    for (size_t i = 0; i < count; i++) {
        if (descs[i].n > SMALL_MSM_MAX_N) big.push_back(i);
        else if (descs[i].n == 0) memcpy(out_jac + i * CDP_JACOBIAN_BYTES, INF_JAC_ZERO, CDP_JACOBIAN_BYTES);
        else order[cls(descs[i].n)].push_back(i);
    }
    size_t small_total = 0, small_count = 0;
    for (auto &o : order) for (size_t i : o) { small_total += descs[i].n; small_count++; }
    if (small_count) {
        size_t bytes_pts = small_total * CDP_AFFINE_BYTES, bytes_sc = small_total * CDP_SCALAR_BYTES, bytes_seg = small_count * sizeof(msm_seg_t);
        TRY(ensure_host(ctx, ctx->h_stage, bytes_pts + bytes_sc + bytes_seg + small_count * CDP_JACOBIAN_BYTES));
        TRY(ensure_dev(ctx, ctx->d_pts, bytes_pts));
        TRY(ensure_dev(ctx, ctx->d_scalars, bytes_sc));
        TRY(ensure_dev(ctx, ctx->d_segs, bytes_seg));
        TRY(ensure_dev(ctx, ctx->d_jac, small_count * CDP_JACOBIAN_BYTES));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        uint8_t *hp = (uint8_t *)ctx->h_stage.ptr, *hs = hp + bytes_pts;
        msm_seg_t *hseg = (msm_seg_t *)(hs + bytes_sc);
        uint8_t *hout = (uint8_t *)(hseg + small_count);
        size_t off = 0, k = 0;
        for (auto &o : order)
            for (size_t i : o) {
                memcpy(hp + off * CDP_AFFINE_BYTES, descs[i].affine_pts, descs[i].n * CDP_AFFINE_BYTES);
                memcpy(hs + off * CDP_SCALAR_BYTES, descs[i].scalars, descs[i].n * CDP_SCALAR_BYTES);
                hseg[k].pts_off = (uint32_t)off; hseg[k].scalars_off = (uint32_t)off; hseg[k].n = (uint32_t)descs[i].n; hseg[k].extra = 0;
                off += descs[i].n; k++;
            }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pts.ptr, hp, bytes_pts, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scalars.ptr, hs, bytes_sc, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_segs.ptr, hseg, bytes_seg, cudaMemcpyHostToDevice, ctx->stream));
        size_t first = 0;
        for (auto &o : order) {
            if (o.empty()) continue;
            size_t max_n = 0, pairs = 0;
            for (size_t i : o) { max_n = std::max(max_n, descs[i].n); pairs += descs[i].n; }
            TRY(cdp_msm_batch_dev(ctx, (const uint8_t *)ctx->d_pts.ptr, (const uint8_t *)ctx->d_scalars.ptr,
                                  reinterpret_cast<const cdp_msm_seg *>(ctx->d_segs.ptr) + first, o.size(), max_n, pairs,
                                  (uint8_t *)ctx->d_jac.ptr + first * CDP_JACOBIAN_BYTES));
            first += o.size();
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(hout, ctx->d_jac.ptr, small_count * CDP_JACOBIAN_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        k = 0;
        for (auto &o : order)
            for (size_t i : o) memcpy(out_jac + i * CDP_JACOBIAN_BYTES, hout + (k++) * CDP_JACOBIAN_BYTES, CDP_JACOBIAN_BYTES);
    }
    for (size_t i : big) TRY(cdp_msm(ctx, descs[i].affine_pts, descs[i].scalars, descs[i].n, out_jac + i * CDP_JACOBIAN_BYTES));
    return CDP_OK;
}

// host buffers -> one job over a combined device array [src n | add n]; the result overwrites the src range
static int smul_host(cdp_ctx *ctx, const uint8_t *pts, const uint8_t *scalars, size_t n_scalars, int bcast, const uint8_t *add, size_t n,
                     uint8_t *out_affine) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return CDP_OK;
    if (n >= (size_t(1) << 30)) return fail(ctx, CDP_ERR_TOO_LARGE, "too many points");
    TRY(ensure_dev(ctx, ctx->d_pts, 2 * n * CDP_AFFINE_BYTES));
    TRY(ensure_dev(ctx, ctx->d_scalars, n_scalars * CDP_SCALAR_BYTES));
    TRY(ensure_dev(ctx, ctx->d_segs, sizeof(smul_job_t)));
    TRY(ensure_host(ctx, ctx->h_stage, sizeof(smul_job_t)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    smul_job_t *job = (smul_job_t *)ctx->h_stage.ptr;
    memset(job, 0, sizeof *job);
    job->src_off = 0;
    job->add_off = add ? (uint32_t)n : 0xFFFFFFFFu;
    job->out_off = 0;
    job->scalar_off = 0;
    job->scalar_stride = bcast ? 0 : 1;
    uint8_t *dp = (uint8_t *)ctx->d_pts.ptr;
    CUDA_TRY(ctx, cudaMemcpyAsync(dp, pts, n * CDP_AFFINE_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    if (add) CUDA_TRY(ctx, cudaMemcpyAsync(dp + n * CDP_AFFINE_BYTES, add, n * CDP_AFFINE_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scalars.ptr, scalars, n_scalars * CDP_SCALAR_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_segs.ptr, job, sizeof *job, cudaMemcpyHostToDevice, ctx->stream));
    TRY(smul_jobs_dev(ctx, dp, (const uint8_t *)ctx->d_scalars.ptr, (const smul_job_t *)ctx->d_segs.ptr, 1, n));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_affine, dp, n * CDP_AFFINE_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}

extern "C" int cdp_fold(cdp_ctx *ctx, const uint8_t *L_affine, const uint8_t *R_affine, const uint8_t gamma[CDP_SCALAR_BYTES], size_t n,
                        uint8_t *out_affine) {
    if (!ctx || !gamma || (n && (!L_affine || !R_affine || !out_affine))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_fold: null argument");
    return smul_host(ctx, R_affine, gamma, 1, 1, L_affine, n, out_affine);
}
extern "C" int cdp_scalar_mul_batch(cdp_ctx *ctx, const uint8_t *affine_pts, const uint8_t *scalars, size_t n, uint8_t *out_affine) {
    if (!ctx || (n && (!affine_pts || !scalars || !out_affine))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_scalar_mul_batch: null argument");
    return smul_host(ctx, affine_pts, scalars, n, 0, nullptr, n, out_affine);
}

static int normalize_host(cdp_ctx *ctx, const uint8_t *jac_pts, size_t n, uint8_t *out, bool compressed) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return CDP_OK;
    size_t out_bytes = n * (compressed ? CDP_COMPRESSED_BYTES : CDP_AFFINE_BYTES);
    TRY(ensure_dev(ctx, ctx->d_jac, n * CDP_JACOBIAN_BYTES));
    TRY(ensure_dev(ctx, ctx->d_out, out_bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_jac.ptr, jac_pts, n * CDP_JACOBIAN_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    TRY(normalize_dev(ctx, (const uint8_t *)ctx->d_jac.ptr, n, compressed ? nullptr : (uint8_t *)ctx->d_out.ptr,
                      compressed ? (uint8_t *)ctx->d_out.ptr : nullptr));
    CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}
extern "C" int cdp_normalize_batch(cdp_ctx *ctx, const uint8_t *jac_pts, size_t n, uint8_t *out_affine) {
    if (!ctx || (n && (!jac_pts || !out_affine))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_normalize_batch: null argument");
    return normalize_host(ctx, jac_pts, n, out_affine, false);
}
extern "C" int cdp_compress_batch(cdp_ctx *ctx, const uint8_t *jac_pts, size_t n, uint8_t *out_compressed) {
    if (!ctx || (n && (!jac_pts || !out_compressed))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_compress_batch: null argument");
    return normalize_host(ctx, jac_pts, n, out_compressed, true);
}

// =================================================================================================== fixed-base MSM
struct cdp_fixed_table {
    uint32_t *d_table = nullptr;
    size_t n_bases = 0, bytes = 0;
    int device = 0;
    fixed_kparams_t kp;
};

extern "C" int cdp_fixed_table_create(cdp_ctx *ctx, const uint8_t *affine_pts, size_t n_bases, int window_bits, cdp_fixed_table **out) {
    if (!ctx || !out || !affine_pts || n_bases == 0) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_fixed_table_create: bad argument");
    *out = nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // default 16 bits: 16 additions per pair, 12 GiB for the ell = 252 CRS.  Wider windows work (19 bits: 14 additions, 86 GiB, the
    // kernel alone 10% faster) but did not move the whole prover (3725 vs 3795 proofs/s), so they stay opt-in (CDP_FIXED_BITS / argument).
    if (window_bits == 0) window_bits = 16;
    if (window_bits < 2 || window_bits > 20) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_fixed_table_create: window_bits must be 2..20");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int c = window_bits, nw = (256 + c - 1) / c;
    const uint32_t nd = 1u << (c - 1);
    const size_t chains = n_bases * (size_t)nw;
    if (chains * nd >= (size_t(1) << 32)) return fail(ctx, CDP_ERR_TOO_LARGE, "cdp_fixed_table_create: table too large");
    cdp_fixed_table *t = new cdp_fixed_table();
    t->n_bases = n_bases;
    t->device = ctx->device;
    t->kp.c = c; t->kp.nw = nw; t->kp.nd = nd;
    for (int i = 0; i < 8; i++) t->kp.recode[i] = 0;
    for (int w = 0; w + 1 < nw; w++) {
        int bit = c * w + c - 1;
        t->kp.recode[bit >> 5] |= 1u << (bit & 31);
    }
    // entry stride: 128 bytes -- one DRAM line per gathered entry (ncu: 2.1 KB of DRAM traffic per pair for 1.5 KB of entries; packed at
    // 96 bytes an entry straddles two lines half of the time: 3.1 KB per pair, profiles/r02_fixed_stride.txt).  Same kernel time either
    // way (the gather is latency-hidden, not bandwidth-bound); CDP_FIXED_STRIDE=96 selects the packed table (3/4 of the memory).
    static const uint32_t stride_bytes = [] { const char *e = getenv("CDP_FIXED_STRIDE"); return e && atoi(e) == 96 ? 96u : 128u; }();
    t->kp.es = stride_bytes / 4;
    t->bytes = chains * nd * (size_t)stride_bytes;
    if (cudaMalloc(&t->d_table, t->bytes) != cudaSuccess) {
        delete t;
        return fail(ctx, CDP_ERR_CUDA, "cdp_fixed_table_create: cudaMalloc of the digit table failed");
    }
    auto bail = [&](int rc) { cudaStreamSynchronize(ctx->stream); cudaFree(t->d_table); delete t; return rc; };
    // bases -> 2^(c w) B (Jacobian) -> affine -> digit 1 of every chain -> c - 1 doubling levels
    if (int rc = ensure_dev(ctx, ctx->d_pts, n_bases * CDP_AFFINE_BYTES)) return bail(rc);
    if (int rc = ensure_dev(ctx, ctx->d_jac, chains * CDP_JACOBIAN_BYTES)) return bail(rc);
    if (int rc = ensure_dev(ctx, ctx->d_out, chains * CDP_AFFINE_BYTES)) return bail(rc);
    cudaError_t e = cudaMemcpyAsync(ctx->d_pts.ptr, affine_pts, n_bases * CDP_AFFINE_BYTES, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_fixed_pow(ctx->stream, (const uint32_t *)ctx->d_pts.ptr, (uint32_t)n_bases, c, nw, (uint32_t *)ctx->d_jac.ptr);
    if (e == cudaSuccess) e = launch_normalize(ctx->stream, 1, (const uint32_t *)ctx->d_jac.ptr, (uint32_t *)ctx->d_out.ptr, nullptr, (uint32_t)chains, nullptr, 1);
    if (e == cudaSuccess) e = launch_fixed_seed(ctx->stream, (const uint32_t *)ctx->d_out.ptr, (uint32_t)chains, nd, t->kp.es, t->d_table);
    ctx->launches += 3;
    for (uint32_t half = 1; half < nd && e == cudaSuccess; half <<= 1) {
        e = launch_fixed_level(ctx->stream, t->d_table, (uint32_t)chains, nd, half, t->kp.es);
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, CDP_ERR_CUDA, std::string("cdp_fixed_table_create: ") + cudaGetErrorString(e)));
    *out = t;
    return CDP_OK;
}
extern "C" void cdp_fixed_table_destroy(cdp_ctx *ctx, cdp_fixed_table *t) {
    if (!t) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(t->d_table);
    delete t;
}
extern "C" size_t cdp_fixed_table_bytes(const cdp_fixed_table *t) { return t ? t->bytes : 0; }
extern "C" size_t cdp_fixed_table_bases(const cdp_fixed_table *t) { return t ? t->n_bases : 0; }

extern "C" int cdp_msm_fixed_batch_dev_lanes(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                                             size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac, int lanes_per_segment) {
    if (!ctx || !t || (count && (!d_scalars || !d_segs || !d_out_jac))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed_batch_dev: null argument");
    if (lanes_per_segment != 0 && lanes_per_segment != 8 && lanes_per_segment != 16 && lanes_per_segment != 32)
        return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed_batch_dev_lanes: lanes_per_segment must be 0 (= 32), 8, 16 or 32");
    if (count == 0) return CDP_OK;
    if (t->device != ctx->device) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed_batch_dev: table lives on another device");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cdp_fixed_seg) == sizeof(fixed_seg_t), "fixed segment layout");
    launch_scope ls(ctx, CDP_PROFILE_MSM_FIXED, total_pairs);
    CUDA_TRY(ctx, launch_fixed_msm(ctx->stream, t->d_table, reinterpret_cast<const uint32_t *>(d_scalars), reinterpret_cast<const fixed_seg_t *>(d_segs),
                                   (uint32_t)count, t->kp, reinterpret_cast<const uint32_t *>(d_var_pts), reinterpret_cast<uint32_t *>(d_out_jac),
                                   lanes_per_segment ? lanes_per_segment : 32));
    return CDP_OK;
}
// The same sums as a tree of batched affine additions (k_fixed.cu): for launches of many long segments.  CDP_FIXED_TREE_ROUNDS (default 5),
// CDP_FIXED_TREE_KMAX (64) and CDP_FIXED_TREE_THREADS are tuning knobs.
extern "C" int cdp_msm_fixed_batch_dev_tree(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                                            size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac, size_t max_pairs_per_segment) {
    if (!ctx || !t || (count && (!d_scalars || !d_segs || !d_out_jac))) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed_batch_dev_tree: null argument");
    if (count == 0) return CDP_OK;
    if (t->device != ctx->device) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed_batch_dev_tree: table lives on another device");
    static const int rounds = (int)env_u32("CDP_FIXED_TREE_ROUNDS", 5);
    static const uint32_t kmax = env_u32("CDP_FIXED_TREE_KMAX", 128), t_target = env_u32("CDP_FIXED_TREE_WAVE", 148 * 4 * 128);
    if (max_pairs_per_segment == 0 || max_pairs_per_segment * (size_t)t->kp.nw < (size_t(4) << rounds) ||
        count * (max_pairs_per_segment * (size_t)t->kp.nw + 64) >= (size_t(1) << 30))  // too short for a tree (or too many slots for 30-bit job numbers): the lane kernel
        return cdp_msm_fixed_batch_dev_lanes(ctx, t, d_scalars, d_segs, count, total_pairs, d_var_pts, d_out_jac, 32);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(ensure_dev(ctx, ctx->d_big, fixed_ba_scratch_bytes((uint32_t)count, (uint32_t)max_pairs_per_segment, t->kp.nw, rounds)));
    launch_scope ls(ctx, CDP_PROFILE_MSM_FIXED, total_pairs);
    CUDA_TRY(ctx, launch_fixed_msm_ba(ctx->stream, t->d_table, reinterpret_cast<const uint32_t *>(d_scalars), reinterpret_cast<const fixed_seg_t *>(d_segs),
                                      (uint32_t)count, t->kp, reinterpret_cast<const uint32_t *>(d_var_pts), reinterpret_cast<uint32_t *>(d_out_jac),
                                      (uint32_t)max_pairs_per_segment, rounds, (uint32_t *)ctx->d_big.ptr, t_target, kmax));
    ctx->launches += (uint64_t)rounds;
    return CDP_OK;
}
extern "C" int cdp_msm_fixed_batch_dev(cdp_ctx *ctx, const cdp_fixed_table *t, const uint8_t *d_scalars, const cdp_fixed_seg *d_segs, size_t count,
                                       size_t total_pairs, const uint8_t *d_var_pts, uint8_t *d_out_jac) {
    return cdp_msm_fixed_batch_dev_lanes(ctx, t, d_scalars, d_segs, count, total_pairs, d_var_pts, d_out_jac, 32);
}

extern "C" int cdp_msm_fixed(cdp_ctx *ctx, const cdp_fixed_table *t, size_t base_off, const uint8_t *scalars, size_t n,
                             uint8_t out_jac[CDP_JACOBIAN_BYTES]) {
    if (!ctx || !t || !out_jac || (n && !scalars)) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed: null argument");
    if (base_off + n > t->n_bases) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_msm_fixed: base range outside the table");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) {
        memcpy(out_jac, INF_JAC_ZERO, CDP_JACOBIAN_BYTES);
        return CDP_OK;
    }
    // one warp per 64 pairs, partial sums added by one more warp
    const size_t chunk = 64, nseg = (n + chunk - 1) / chunk;
    TRY(ensure_dev(ctx, ctx->d_scalars, n * CDP_SCALAR_BYTES));
    TRY(ensure_dev(ctx, ctx->d_segs, nseg * sizeof(fixed_seg_t)));
    TRY(ensure_dev(ctx, ctx->d_jac, (nseg + 1) * CDP_JACOBIAN_BYTES));
    TRY(ensure_host(ctx, ctx->h_stage, nseg * sizeof(fixed_seg_t)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    fixed_seg_t *hs = (fixed_seg_t *)ctx->h_stage.ptr;
    for (size_t i = 0; i < nseg; i++) {
        memset(&hs[i], 0, sizeof hs[i]);
        hs[i].base_off = (uint32_t)(base_off + i * chunk);
        hs[i].scalars_off = (uint32_t)(i * chunk);
        hs[i].n = (uint32_t)std::min(chunk, n - i * chunk);
        hs[i].remap_from = 0xFFFFFFFFu;
        hs[i].out_idx = (uint32_t)i;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scalars.ptr, scalars, n * CDP_SCALAR_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_segs.ptr, hs, nseg * sizeof(fixed_seg_t), cudaMemcpyHostToDevice, ctx->stream));
    uint8_t *dj = (uint8_t *)ctx->d_jac.ptr;
    TRY(cdp_msm_fixed_batch_dev(ctx, t, (const uint8_t *)ctx->d_scalars.ptr, (const cdp_fixed_seg *)ctx->d_segs.ptr, nseg, n, nullptr, dj));
    const uint8_t *res = dj;
    if (nseg > 1) {
        TRY(cdp_sum_jacobian_dev(ctx, dj, nseg, dj + nseg * CDP_JACOBIAN_BYTES));
        res = dj + nseg * CDP_JACOBIAN_BYTES;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out_jac, res, CDP_JACOBIAN_BYTES, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CDP_OK;
}

// =================================================================================================== profiling
// =================================================================================================== multi-GPU (NCCL)
// NCCL is resolved at run time: the library stays loadable (and single-GPU use unaffected) where NCCL is absent, and inside a PyTorch
// process dlopen by soname returns the copy torch already loaded.
namespace {
struct nccl_api {
    void *handle = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
nccl_api *nccl() {
    static nccl_api api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("CDP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
            api.err = dlerror();
        }
        if (!api.handle) return;
        bool ok = true;
        auto sym = [&](const char *nm) { void *p = dlsym(api.handle, nm); if (!p) { ok = false; api.err = std::string("missing symbol ") + nm; } return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return &api;
}
}  // namespace

struct cdp_comm {
    cdp_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    uint8_t *d_mine = nullptr, *d_all = nullptr;  // this rank's partial sum; the gathered n_ranks partial sums
    std::string err = "ok";
};
namespace {
int comm_fail(cdp_comm *c, int code, const std::string &msg) {
    if (c) { c->err = msg; if (c->ctx) c->ctx->err = msg; }
    return code;
}
#define NCCL_TRY(c, expr)                                                                                                   \
    do {                                                                                                                    \
        ncclResult_t r__ = (expr);                                                                                          \
        if (r__ != ncclSuccess) return comm_fail(c, CDP_ERR_NCCL, std::string(#expr) + ": " + nccl()->GetErrorString(r__)); \
    } while (0)
int comm_alloc(cdp_comm *c) {
    if (cudaSetDevice(c->ctx->device) != cudaSuccess || cudaMalloc(&c->d_mine, 144) != cudaSuccess ||
        cudaMalloc(&c->d_all, 144 * (size_t)c->n_ranks) != cudaSuccess)
        return comm_fail(c, CDP_ERR_CUDA, "cdp_comm: allocation failed");
    return CDP_OK;
}
// the local part: this rank's Pippenger over its shard (or the point at infinity for an empty shard) into d_mine
int sharded_local(cdp_comm *c, const uint8_t *d_pts, const uint8_t *d_scalars, size_t n_local) {
    cdp_ctx *ctx = c->ctx;
    if (n_local == 0) {
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        CUDA_TRY(ctx, cudaMemsetAsync(c->d_mine, 0, 144, ctx->stream));  // Z = 0
        return CDP_OK;
    }
    return cdp_msm_dev(ctx, d_pts, d_scalars, n_local, c->d_mine);
}
int sharded_gather(cdp_comm *c, const uint8_t *d_partial) {
    CUDA_TRY(c->ctx, cudaSetDevice(c->ctx->device));
    NCCL_TRY(c, nccl()->AllGather(d_partial, c->d_all, 144, ncclUint8, c->comm, c->ctx->stream));
    return CDP_OK;
}
}  // namespace

extern "C" int cdp_comm_unique_id(uint8_t out_id[CDP_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == CDP_COMM_ID_BYTES, "ncclUniqueId size");
    if (!out_id) return CDP_ERR_INVALID_ARG;
    if (!nccl()->handle) return CDP_ERR_NCCL;
    ncclUniqueId id;
    if (nccl()->GetUniqueId(&id) != ncclSuccess) return CDP_ERR_NCCL;
    memcpy(out_id, &id, sizeof id);
    return CDP_OK;
}
extern "C" int cdp_comm_create(cdp_comm **out, cdp_ctx *ctx, const uint8_t id[CDP_COMM_ID_BYTES], int n_ranks, int rank) {
    if (!out || !ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_comm_create: bad argument");
    *out = nullptr;
    if (!nccl()->handle) return fail(ctx, CDP_ERR_NCCL, "cdp_comm_create: NCCL not available: " + nccl()->err);
    cdp_comm *c = new cdp_comm();
    c->ctx = ctx; c->n_ranks = n_ranks; c->rank = rank;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclResult_t r = cudaSetDevice(ctx->device) == cudaSuccess ? nccl()->CommInitRank(&c->comm, n_ranks, uid, rank) : ncclUnhandledCudaError;
    if (r != ncclSuccess) {
        fail(ctx, CDP_ERR_NCCL, std::string("ncclCommInitRank: ") + nccl()->GetErrorString(r));
        delete c;
        return CDP_ERR_NCCL;
    }
    if (int rc = comm_alloc(c)) { cdp_comm_destroy(c); return rc; }
    *out = c;
    return CDP_OK;
}
extern "C" int cdp_comm_create_all(cdp_comm **out, cdp_ctx *const *ctxs, int n) {
    if (!out || !ctxs || n < 1) return CDP_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) { out[i] = nullptr; if (!ctxs[i]) return CDP_ERR_INVALID_ARG; }
    if (!nccl()->handle) return fail(ctxs[0], CDP_ERR_NCCL, "cdp_comm_create_all: NCCL not available: " + nccl()->err);
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    ncclResult_t r = nccl()->CommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) return fail(ctxs[0], CDP_ERR_NCCL, std::string("ncclCommInitAll: ") + nccl()->GetErrorString(r));
    for (int i = 0; i < n; i++) {
        cdp_comm *c = new cdp_comm();
        c->ctx = ctxs[i]; c->comm = comms[i]; c->n_ranks = n; c->rank = i;
        out[i] = c;
    }
    for (int i = 0; i < n; i++)
        if (int rc = comm_alloc(out[i])) {
            for (int j = 0; j < n; j++) { cdp_comm_destroy(out[j]); out[j] = nullptr; }
            return rc;
        }
    return CDP_OK;
}
extern "C" void cdp_comm_destroy(cdp_comm *c) {
    if (!c) return;
    if (c->ctx) { cudaSetDevice(c->ctx->device); cudaStreamSynchronize(c->ctx->stream); }
    if (c->comm && nccl()->handle) nccl()->CommDestroy(c->comm);
    if (c->d_mine) cudaFree(c->d_mine);
    if (c->d_all) cudaFree(c->d_all);
    delete c;
}
extern "C" int cdp_comm_rank(const cdp_comm *c) { return c ? c->rank : -1; }
extern "C" int cdp_comm_size(const cdp_comm *c) { return c ? c->n_ranks : 0; }
extern "C" cdp_ctx *cdp_comm_ctx(const cdp_comm *c) { return c ? c->ctx : nullptr; }
extern "C" const char *cdp_comm_last_error(const cdp_comm *c) { return c ? c->err.c_str() : "null communicator"; }
extern "C" void cdp_shard_range(size_t n, int rank, int n_ranks, size_t *lo, size_t *hi) {
    const size_t w = (size_t)(n_ranks > 0 ? n_ranks : 1), r = (size_t)(rank > 0 ? rank : 0), base = n / w, rem = n % w;
    const size_t l = r * base + std::min(r, rem);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (r < rem ? 1 : 0);
}
extern "C" int cdp_allreduce_jacobian_dev(cdp_comm *c, const uint8_t *d_partial_jac, uint8_t *d_out_jac) {
    if (!c || !d_partial_jac || !d_out_jac) return comm_fail(c, CDP_ERR_INVALID_ARG, "cdp_allreduce_jacobian_dev: bad argument");
    TRY(sharded_gather(c, d_partial_jac));
    return cdp_sum_jacobian_dev(c->ctx, c->d_all, (size_t)c->n_ranks, d_out_jac);
}
extern "C" int cdp_msm_sharded_dev(cdp_comm *c, const uint8_t *d_pts, const uint8_t *d_scalars, size_t n_local, uint8_t *d_out_jac) {
    if (!c || !d_out_jac || (n_local && (!d_pts || !d_scalars))) return comm_fail(c, CDP_ERR_INVALID_ARG, "cdp_msm_sharded_dev: bad argument");
    TRY(sharded_local(c, d_pts, d_scalars, n_local));
    return cdp_allreduce_jacobian_dev(c, c->d_mine, d_out_jac);
}
extern "C" int cdp_msm_sharded_group(cdp_comm *const *comms, int n, const uint8_t *const *d_pts, const uint8_t *const *d_scalars, const size_t *n_local,
                                     uint8_t *const *d_out_jac) {
    if (!comms || n < 1 || !d_pts || !d_scalars || !n_local || !d_out_jac) return CDP_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++)
        if (!comms[i] || comms[i]->n_ranks != n || !d_out_jac[i]) return CDP_ERR_INVALID_ARG;
    for (int i = 0; i < n; i++) TRY(sharded_local(comms[i], d_pts[i], d_scalars[i], n_local[i]));
    NCCL_TRY(comms[0], nccl()->GroupStart());
    for (int i = 0; i < n; i++) {
        const int rc = sharded_gather(comms[i], comms[i]->d_mine);
        if (rc != CDP_OK) { nccl()->GroupEnd(); return rc; }
    }
    NCCL_TRY(comms[0], nccl()->GroupEnd());
    for (int i = 0; i < n; i++) TRY(cdp_sum_jacobian_dev(comms[i]->ctx, comms[i]->d_all, (size_t)n, d_out_jac[i]));
    return CDP_OK;
}

static void prof_drain(cdp_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    static const char *dump = getenv("CDP_PROFILE_DUMP");  // optional per-launch log: "kind units ms" per line
    FILE *f = (dump && !ctx->prof_pending.empty()) ? fopen(dump, "a") : nullptr;
    for (auto &r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            if (f) fprintf(f, "%d %llu %.4f\n", r.kind, (unsigned long long)r.units, ms);
            ctx->prof_ms[r.kind] += ms;
            ctx->prof_launches[r.kind]++;
            ctx->prof_units[r.kind] += r.units;
        }
        ctx->prof_pool.push_back(r.e0);
        ctx->prof_pool.push_back(r.e1);
    }
    if (f) fclose(f);
    ctx->prof_pending.clear();
}
extern "C" int cdp_profile_enable(cdp_ctx *ctx, int on) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    prof_drain(ctx);
    ctx->profiling = on != 0;
    return CDP_OK;
}
extern "C" int cdp_profile_reset(cdp_ctx *ctx) {
    if (!ctx) return CDP_ERR_INVALID_ARG;
    prof_drain(ctx);
    for (int k = 0; k < CDP_PROFILE_KINDS; k++) { ctx->prof_ms[k] = 0; ctx->prof_launches[k] = 0; ctx->prof_units[k] = 0; }
    return CDP_OK;
}
extern "C" int cdp_profile_read(cdp_ctx *ctx, double ms[CDP_PROFILE_KINDS], uint64_t launches[CDP_PROFILE_KINDS], uint64_t units[CDP_PROFILE_KINDS]) {
    if (!ctx || !ms || !launches || !units) return CDP_ERR_INVALID_ARG;
    prof_drain(ctx);
    for (int k = 0; k < CDP_PROFILE_KINDS; k++) { ms[k] = ctx->prof_ms[k]; launches[k] = ctx->prof_launches[k]; units[k] = ctx->prof_units[k]; }
    return CDP_OK;
}

// =================================================================================================== diagnostics
extern "C" int cdp_bench_kernel(cdp_ctx *ctx, int which, int blocks, int threads, int iters, float *ms_out) {
    if (!ctx || !ms_out || blocks <= 0 || threads <= 0 || threads > 256) return fail(ctx, CDP_ERR_INVALID_ARG, "cdp_bench_kernel: bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(ensure_dev(ctx, ctx->d_out, (size_t)blocks * threads * 4));
    if (which >= 9 && which <= 32) TRY(ensure_dev(ctx, ctx->d_big, (size_t)blocks * threads * iters * 288));
    auto launch_bench = [&](cudaStream_t st, int w, uint32_t *out, int b, int t, int it) {
        // 9 + INL: batched affine additions (INL: which products are expanded in place, batch_affine.cuh), blocks * threads threads of `iters` additions each over a streamed operand array
        if (w >= 9 && w <= 32) return launch_bench_ba(st, (uint32_t *)ctx->d_big.ptr, (uint32_t)(blocks * threads), (uint32_t)iters, it != iters, w - 9);
        return cdp::launch_bench(st, w, out, b, t, it);
    };
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    CUDA_TRY(ctx, launch_bench(ctx->stream, which, (uint32_t *)ctx->d_out.ptr, blocks, threads, 8));  // warm-up
    CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    CUDA_TRY(ctx, launch_bench(ctx->stream, which, (uint32_t *)ctx->d_out.ptr, blocks, threads, iters));
    CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(e1));
    CUDA_TRY(ctx, cudaEventElapsedTime(ms_out, e0, e1));
    ctx->launches += 2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return CDP_OK;
}
