// Context life cycle, error reporting, host-side step-size controllers.
#include <math.h>

#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>

#include "fedmx_internal.h"

// FEDMX_NVTX=1: every profiled kernel class is also an NVTX range on the calling thread (visible in Nsight Systems / ncu
// --nvtx; the reference has no tracing of its own).  Header-only NVTX v3: no extra library.
static const bool g_nvtx = getenv("FEDMX_NVTX") && atoi(getenv("FEDMX_NVTX")) != 0;
static const char *const kProfNames[FEDMX_PROF_KINDS] = {"fedmx:assemble_jacobian", "fedmx:assemble_residual", "fedmx:spmv",
                                                         "fedmx:orthogonalisation", "fedmx:preconditioner", "fedmx:mg_fine",
                                                         "fedmx:mg_mid", "fedmx:mg_coarse", "fedmx:comm"};

extern "C" const char *fedmx_version(void) { return "fedmx 0.1 (sm_100a)"; }

extern "C" int fedmx_ctx_create(int device, fedmx_ctx **out) {
    if (!out) return FEDMX_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return FEDMX_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FEDMX_ECUDA;
    fedmx_ctx *ctx = new fedmx_ctx();
    ctx->device = device;
    ctx->mp = ModelParams{};
    ctx->mp.bc_comp = -1;
    // The library runs on its own BLOCKING stream: it keeps the implicit ordering with work the caller queues on the legacy
    // default stream (torch's copies between solves) and, unlike the legacy stream itself, can be captured into CUDA graphs.
    if (cudaStreamCreate(&ctx->stream) == cudaSuccess) ctx->own_stream = true;
    else ctx->stream = 0;
    *out = ctx;
    return FEDMX_OK;
}

extern "C" int fedmx_ctx_destroy(fedmx_ctx *ctx) {
    if (!ctx) return FEDMX_EINVAL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    fedmx_peer_destroy(ctx);
    fedmx_comm_destroy_internal(ctx);
    fedmx_ilu0_free(ctx);
    for (auto &g : ctx->mg_graph)
        if (g) cudaGraphExecDestroy(g);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    for (void *p : ctx->owned_device) cudaFree(p);
    for (void *p : ctx->amg_device) cudaFree(p);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (auto &pool : ctx->prof_ev)
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    delete ctx;
    return FEDMX_OK;
}

extern "C" const char *fedmx_last_error(const fedmx_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int fedmx_set_stream(fedmx_ctx *ctx, void *cuda_stream) {
    if (!ctx) return FEDMX_EINVAL;
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return FEDMX_OK;
}

extern "C" int64_t fedmx_launch_count(fedmx_ctx *ctx, int reset) {
    if (!ctx) return -1;
    int64_t n = ctx->launches;
    if (reset) ctx->launches = 0;
    return n;
}

// ---- profiling -----------------------------------------------------------------------------------------
ProfScope::ProfScope(fedmx_ctx *ctx, int k) : c(ctx), kind(k) {
    if (g_nvtx) nvtxRangePushA(kProfNames[k]);
    if (!c->prof_on || (k >= FEDMX_PROF_MG_FINE && c->prof_on < 2)) return;
    auto &pool = c->prof_ev[kind];
    if (c->prof_used[kind] + 2 > pool.size()) {
        if (pool.size() >= 8192) fedmx_profile_flush(c, kind);
        else
            for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    }
    cudaEvent_t start = pool[c->prof_used[kind]++];
    stop = pool[c->prof_used[kind]++];
    cudaEventRecord(start, c->stream);
}
ProfScope::~ProfScope() {
    if (g_nvtx) nvtxRangePop();
    if (stop) { cudaEventRecord(stop, c->stream); c->prof_count[kind]++; }
}
int fedmx_profile_flush(fedmx_ctx *ctx, int kind) {
    cudaStreamSynchronize(ctx->stream);
    auto &pool = ctx->prof_ev[kind];
    for (size_t i = 0; i + 1 < ctx->prof_used[kind]; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pool[i], pool[i + 1]) == cudaSuccess) ctx->prof_ms[kind] += ms;
    }
    ctx->prof_used[kind] = 0;
    return FEDMX_OK;
}
extern "C" int fedmx_profile_enable(fedmx_ctx *ctx, int on) {
    if (!ctx) return FEDMX_EINVAL;
    ctx->prof_on = on;
    return FEDMX_OK;
}
extern "C" int fedmx_profile_get(fedmx_ctx *ctx, int kind, int64_t *count, double *total_ms) {
    if (!ctx || kind < 0 || kind >= FEDMX_PROF_KINDS) return FEDMX_EINVAL;
    fedmx_profile_flush(ctx, kind);
    if (count) *count = ctx->prof_count[kind];
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    ctx->prof_count[kind] = 0;
    ctx->prof_ms[kind] = 0.0;
    return FEDMX_OK;
}

// Step-size controllers: plain libm pow() so that the dt sequence is bit-identical to the reference's
// Python floats (reference: fedm/functions.py:915-951; pinned by the golden relative-error logs).
// Python's max(min(dt, dt_max), dt_min): a NaN step size propagates (fmin/fmax would silently turn it into dt_max)
static double clampd(double dt, double dt_min, double dt_max) {
    const double lo = (dt_max < dt) ? dt_max : dt;
    return (dt_min > lo) ? dt_min : lo;
}

extern "C" double fedmx_pid_step(double dt, const double err[3], double tol, double dt_min, double dt_max) {
    dt *= pow(err[1] / err[0], 0.075) * pow(tol / err[0], 0.175) * pow(err[1] * err[1] / (err[0] * err[2]), 0.01);
    return clampd(dt, dt_min, dt_max);
}

extern "C" double fedmx_pi34_step(double dt, const double err[3], double tol, double dt_min, double dt_max) {
    dt *= pow(0.8 * tol / err[0], 0.3 / 3) * pow(0.8 * err[1] / err[0], 0.4 / 3);
    return clampd(dt, dt_min, dt_max);
}

extern "C" double fedmx_h211b_step(double dt, double dt_old, const double err[3], double tol, double dt_min, double dt_max) {
    dt *= pow(0.8 * tol / err[0], 1.0 / 12) * pow(0.8 * tol / err[1], 1.0 / 12) * pow(dt / dt_old, -1.0 / 4);
    return clampd(dt, dt_min, dt_max);
}
