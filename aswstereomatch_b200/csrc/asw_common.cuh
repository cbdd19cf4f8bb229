// asw_common.cuh -- context, device buffers, launch/profiling helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/asw/asw.h"

#define ASW_MAX_PROFILE 48

// Superseded kernels (kept for A/B measurements) are compiled, and selectable through ASW_* environment variables, only
// with -DASW_DEV_KERNELS.  The product build contains one path per method plus the size-generic fallbacks and never
// reads the environment.
#ifdef ASW_DEV_KERNELS
#include <stdlib.h>
static inline bool asw_dev(const char* name) { return getenv(name) != nullptr; }
#else
#define asw_dev(name) false
#endif

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    unsigned gen = 0;          // bumped on every (re)allocation: cached contents are keyed on it
};

struct ProfEntry {
    const char* name;
    double total_ms;
    long long launches;
};

struct asw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // batch transfers overlap the compute stream
    // second compute stream + workspace bank: the two views of an LR frame are independent until the consistency check and
    // run concurrently (view_begin / view_end); `stream` is always the stream the current view launches on
    cudaStream_t stream_main = nullptr, stream_view = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int ws_bank = 0;
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;      // asw_timer_*
    cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;      // per-kernel profiling
    char err[512] = {0};
    int sm_count = 148;
    // named workspaces (grow-only)
    std::vector<DevBuf> bufs;
    // profiling
    bool profiling = false;
    ProfEntry prof[ASW_MAX_PROFILE];
    int n_prof = 0;
    long long launches = 0;
    // capture hook
    float* capture_host = nullptr;
    size_t capture_cap = 0;
    // L2 flush scratch
    DevBuf flush;
    // pinned staging for small D2H results
    void* pinned = nullptr;
    size_t pinned_cap = 0;
    // tuning knobs (asw_set_tuning; 0 = the built-in choice)
    int tune[ASW_TUNE_COUNT] = {0};
    // host-built tables are cached per parameter set: nothing is uploaded (and nothing synchronises) on a repeated call
    std::string table_key[12];                         // 6 per workspace bank
};

// workspace slots
enum {
    WS_IMG_L = 0, WS_IMG_R, WS_FEAT_REF, WS_FEAT_TGT, WS_GUIDE_I, WS_GUIDE_MI, WS_GUIDE_DEN,
    WS_TMP0, WS_TMP1, WS_TMP2, WS_TMP3, WS_VOL0, WS_VOL1, WS_AB, WS_KEYS, WS_KEYS2, WS_DISP_L, WS_DISP_R,
    WS_MASK, WS_FILLED, WS_OUT, WS_SLICE_MM, WS_GRAY_L, WS_GRAY_R, WS_GEO_L, WS_GEO_R, WS_GRID_S, WS_GRID_C,
    WS_TABLE0, WS_TABLE1, WS_MISC0, WS_MISC1, WS_MISC2, WS_MISC3, WS_CAPTURE, WS_GUIDE_RDEN,
    WS_FEATF_REF, WS_FEATF_TGT, WS_GUIDE_NM, WS_GUIDE_RD2, WS_AFF, WS_REFINE_LIST,
    WS_TRAD_C2, WS_TRAD_TABLE, WS_GRID_TI, WS_GRID_TD, WS_IMG_PAD, WS_GUIDE6, WS_COUNT
};

static inline asw_status asw_fail(asw_ctx* ctx, asw_status st, const char* fmt, const char* a = "", const char* b = "") {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b);
    return st;
}

#define ASW_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            return asw_fail((ctx), ASW_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
        }                                                                                \
    } while (0)

#define ASW_TRY(expr)                        \
    do {                                     \
        asw_status s__ = (expr);             \
        if (s__ != ASW_OK) return s__;       \
    } while (0)

// cached host-built table: true when slot `ks` already holds the table described by `what` (same buffer generation)
static inline bool table_cached(asw_ctx* ctx, int ks, int ws_slot, const char* what) {
    char key[160];
    snprintf(key, sizeof(key), "%u|%s", ctx->bufs[ws_slot + ctx->ws_bank * WS_COUNT].gen, what);
    ks += 6 * ctx->ws_bank;
    if (ctx->table_key[ks] == key) return true;
    ctx->table_key[ks] = key;
    return false;
}

// reserve a workspace slot (grow-only; contents undefined after growth)
static inline asw_status ws_reserve(asw_ctx* ctx, int slot, size_t bytes, void** out) {
    if ((int)ctx->bufs.size() < 2 * WS_COUNT) ctx->bufs.resize(2 * WS_COUNT);
    slot += ctx->ws_bank * WS_COUNT;
    DevBuf& b = ctx->bufs[slot];
    if (b.cap < bytes) {
        if (b.p) {
            ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->stream_view) ASW_CUDA(ctx, cudaStreamSynchronize(ctx->ws_bank ? ctx->stream_main : ctx->stream_view));
            ASW_CUDA(ctx, cudaFree(b.p));
            b.p = nullptr; b.cap = 0;
        }
        size_t cap = (bytes + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&b.p, cap);
        if (e != cudaSuccess) {
            b.p = nullptr;
            return asw_fail(ctx, ASW_ERR_NOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
        }
        b.cap = cap;
        b.gen++;
    }
    *out = b.p;
    return ASW_OK;
}
template <typename T>
static inline asw_status ws_get(asw_ctx* ctx, int slot, size_t count, T** out) {
    void* p = nullptr;
    asw_status s = ws_reserve(ctx, slot, count * sizeof(T), &p);
    *out = (T*)p;
    return s;
}

// profiling-aware launch bracket: PROF_BEGIN(ctx); kernel<<<...>>>(...); PROF_END(ctx, "name");
static inline void prof_begin(asw_ctx* ctx) {
    if (ctx->profiling) cudaEventRecord(ctx->ev_p0, ctx->stream);
}
static inline asw_status prof_end(asw_ctx* ctx, const char* name) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return asw_fail(ctx, ASW_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e));
    if (ctx->profiling) {
        cudaEventRecord(ctx->ev_p1, ctx->stream);
        cudaEventSynchronize(ctx->ev_p1);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_p0, ctx->ev_p1);
        int i = 0;
        for (; i < ctx->n_prof; i++)
            if (strcmp(ctx->prof[i].name, name) == 0) break;
        if (i == ctx->n_prof && ctx->n_prof < ASW_MAX_PROFILE) {
            ctx->prof[i].name = name; ctx->prof[i].total_ms = 0; ctx->prof[i].launches = 0;
            ctx->n_prof++;
        }
        if (i < ctx->n_prof) { ctx->prof[i].total_ms += ms; ctx->prof[i].launches++; }
    }
    return ASW_OK;
}
#define LAUNCH(ctx, name, ...)            \
    do {                                  \
        prof_begin(ctx);                  \
        __VA_ARGS__;                      \
        ASW_TRY(prof_end(ctx, name));     \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Run what follows on the second compute stream with the second workspace bank (ordered after everything already on the
// main stream), until view_switch_back() / view_join().  Used for the right view of an LR frame.
static inline asw_status view_begin(asw_ctx* ctx) {
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream_main));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_view, ctx->ev_fork, 0));
    ctx->stream = ctx->stream_view; ctx->ws_bank = 1;
    return ASW_OK;
}
// back to the main stream / bank; the second stream keeps running
static inline void view_switch_back(asw_ctx* ctx) { ctx->stream = ctx->stream_main; ctx->ws_bank = 0; }
// the main stream waits for everything launched on the second one
static inline asw_status view_join(asw_ctx* ctx) {
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream_view));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream_main, ctx->ev_join, 0));
    return ASW_OK;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// cv::borderInterpolate: delta = 0 BORDER_REFLECT, delta = 1 BORDER_REFLECT_101
__device__ __forceinline__ int border_idx(int p, int len, int delta) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = len - 1 - (p - len) - delta;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// total order on float bits such that unsigned compare == float compare; NaN -> 0xFFFFFFFF
__device__ __forceinline__ uint32_t orderable_u32(float f) {
    if (f != f) return 0xFFFFFFFFu;
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long orderable_u64(double f) {
    if (f != f) return 0xFFFFFFFFFFFFFFFFull;
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
// 64-bit WTA key: top 48 bits = orderable double cost (sign, exponent, 36 mantissa bits: float costs are
// represented exactly), low 16 bits = disparity label.  Unsigned min over keys == the reference's WTA
// (A.cpp:3032-3048): strict <, ascending d so the lowest d wins ties; NaN and +inf never beat DBL_MAX.
__device__ __forceinline__ unsigned long long wta_key_d(double cost, int d) {
    return (orderable_u64(cost) & 0xFFFFFFFFFFFF0000ull) | (unsigned long long)(d & 0xFFFF);
}
__device__ __forceinline__ unsigned long long wta_key(float cost, int d) { return wta_key_d((double)cost, d); }
#define WTA_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define WTA_KEY_INF_TOP 0xFFF0000000000000ull   /* orderable(+inf) */
