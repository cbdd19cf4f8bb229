// asw_lib.cu -- C ABI (include/asw/asw.h) over the sm_100a kernels.  Single translation unit.
// Host orchestration only: argument checks mirroring the reference's early-outs, H2D / D2H,
// workspace management, kernel launches.  No CPU compute path: every stage is a CUDA kernel.
#include "asw_common.cuh"
#include "k_prep.cuh"
#include "k_cost.cuh"
#include "k_guided.cuh"
#include "k_guided_fast.cuh"
#include "k_guided_stream.cuh"
#include "k_refine.cuh"
#include "k_ncc.cuh"

#include <math.h>
#include <stdlib.h>

// =================================================================================================
// context
// =================================================================================================
extern "C" int asw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
#ifdef ASW_DEV_KERNELS
extern "C" const char* asw_version(void) { return "aswstereomatch_b200 0.2 (sm_100a) +dev-kernels"; }
#else
extern "C" const char* asw_version(void) { return "aswstereomatch_b200 0.2 (sm_100a)"; }
#endif
extern "C" asw_status asw_set_tuning(asw_ctx* ctx, int key, int value) {
    if (!ctx || key < 0 || key >= ASW_TUNE_COUNT || value < 0) return ASW_ERR_BAD_ARG;
    ctx->tune[key] = value;
    return ASW_OK;
}

extern "C" asw_status asw_create(int device, asw_ctx** out) {
    if (!out) return ASW_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return ASW_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ASW_ERR_CUDA;
    asw_ctx* ctx = new asw_ctx();
    ctx->device = device;
    ctx->bufs.resize(2 * WS_COUNT);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream_view, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_t0) != cudaSuccess || cudaEventCreate(&ctx->ev_t1) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_p0) != cudaSuccess || cudaEventCreate(&ctx->ev_p1) != cudaSuccess) {
        // release whatever was created before the failure (the handles start out null)
        if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
        if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
        if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
        if (ctx->stream_view) cudaStreamDestroy(ctx->stream_view);
        if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
        if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
        if (ctx->ev_p0) cudaEventDestroy(ctx->ev_p0);
        if (ctx->ev_p1) cudaEventDestroy(ctx->ev_p1);
        if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
        if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return ASW_ERR_CUDA;
    }
    ctx->stream_main = ctx->stream;
    *out = ctx;
    return ASW_OK;
}
extern "C" void asw_destroy(asw_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamSynchronize(ctx->stream_main);
    cudaStreamSynchronize(ctx->stream_view);
    cudaStreamSynchronize(ctx->d2h_stream);
    for (auto& b : ctx->bufs) if (b.p) cudaFree(b.p);
    if (ctx->flush.p) cudaFree(ctx->flush.p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1);
    cudaEventDestroy(ctx->ev_p0); cudaEventDestroy(ctx->ev_p1); cudaEventDestroy(ctx->ev_copy);
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->h2d_stream); cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->stream_main); cudaStreamDestroy(ctx->stream_view);
    delete ctx;
}
extern "C" const char* asw_last_error(const asw_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
extern "C" asw_status asw_sync(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream_view));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    return ASW_OK;
}
extern "C" void* asw_stream(asw_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" void* asw_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void asw_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" asw_status asw_timer_start(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_timer_stop(asw_ctx* ctx, float* ms) {
    if (!ctx || !ms) return ASW_ERR_BAD_ARG;
    // the timed region ends when the batch transfer streams have drained too
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->d2h_stream));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
    ASW_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    ASW_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return ASW_OK;
}
extern "C" asw_status asw_profile_enable(asw_ctx* ctx, int on) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ctx->profiling = on != 0;
    return ASW_OK;
}
extern "C" asw_status asw_profile_reset(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ctx->n_prof = 0;
    ctx->launches = 0;
    return ASW_OK;
}
extern "C" int asw_profile_count(asw_ctx* ctx) { return ctx ? ctx->n_prof : 0; }
extern "C" asw_status asw_profile_entry(asw_ctx* ctx, int i, const char** name, double* total_ms, long long* launches) {
    if (!ctx || i < 0 || i >= ctx->n_prof) return ASW_ERR_BAD_ARG;
    if (name) *name = ctx->prof[i].name;
    if (total_ms) *total_ms = ctx->prof[i].total_ms;
    if (launches) *launches = ctx->prof[i].launches;
    return ASW_OK;
}
extern "C" long long asw_launch_count(asw_ctx* ctx) { return ctx ? ctx->launches : 0; }

__global__ void k_flush_l2(uint4* __restrict__ p, size_t n, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(v, v, v, v);
}
extern "C" asw_status asw_flush_l2(asw_ctx* ctx) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
    if (!ctx->flush.p) {
        ASW_CUDA(ctx, cudaMalloc(&ctx->flush.p, bytes));
        ctx->flush.cap = bytes;
    }
    static uint32_t tick = 0;
    k_flush_l2<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4*)ctx->flush.p, bytes / 16, ++tick);
    ASW_CUDA(ctx, cudaGetLastError());
    return ASW_OK;
}
extern "C" asw_status asw_capture_aggregated(asw_ctx* ctx, float* host_volume, size_t capacity_floats) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ctx->capture_host = host_volume;
    ctx->capture_cap = host_volume ? capacity_floats : 0;
    return ASW_OK;
}

// =================================================================================================
// host <-> device helpers
// =================================================================================================
static asw_status check_u8(asw_ctx* ctx, const asw_u8_image* im, int channels) {
    if (!im || !im->data || im->rows <= 0 || im->cols <= 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "empty image%s%s");
    if (im->channels != channels) return asw_fail(ctx, ASW_ERR_BAD_ARG, "unexpected channel count%s%s");
    if (im->step < (size_t)im->cols * channels) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad step%s%s");
    return ASW_OK;
}
static asw_status check_f32(asw_ctx* ctx, const asw_f32_image* im) {
    if (!im || !im->data || im->rows <= 0 || im->cols <= 0 || im->step < (size_t)im->cols * 4)
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad float image%s%s");
    return ASW_OK;
}
static asw_status check_mask(asw_ctx* ctx, const asw_mask_image* im) {
    if (!im || !im->data || im->rows <= 0 || im->cols <= 0 || im->step < (size_t)im->cols)
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad mask image%s%s");
    return ASW_OK;
}
static asw_status check_pair(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, const asw_f32_image* disp) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_u8(ctx, L, 3));
    ASW_TRY(check_u8(ctx, R, 3));
    if (L->rows != R->rows || L->cols != R->cols) return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "left/right sizes differ%s%s");
    if (disp) {
        ASW_TRY(check_f32(ctx, disp));
        if (disp->rows != L->rows || disp->cols != L->cols)
            return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "disparity map size differs from the images%s%s");
    }
    return ASW_OK;
}
static asw_status upload_u8(asw_ctx* ctx, const asw_u8_image* im, uint8_t* dst) {
    size_t rowb = (size_t)im->cols * im->channels;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(dst, rowb, im->data, im->step, rowb, im->rows, cudaMemcpyHostToDevice, ctx->stream));
    return ASW_OK;
}
static asw_status upload_f32(asw_ctx* ctx, const asw_f32_image* im, float* dst) {
    size_t rowb = (size_t)im->cols * 4;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(dst, rowb, im->data, im->step, rowb, im->rows, cudaMemcpyHostToDevice, ctx->stream));
    return ASW_OK;
}
static asw_status upload_mask(asw_ctx* ctx, const asw_mask_image* im, uint8_t* dst) {
    ASW_CUDA(ctx, cudaMemcpy2DAsync(dst, im->cols, im->data, im->step, im->cols, im->rows, cudaMemcpyHostToDevice, ctx->stream));
    return ASW_OK;
}
static asw_status download_f32(asw_ctx* ctx, const float* src, asw_f32_image* im) {
    size_t rowb = (size_t)im->cols * 4;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(im->data, im->step, src, rowb, rowb, im->rows, cudaMemcpyDeviceToHost, ctx->stream));
    return ASW_OK;
}
static asw_status download_mask(asw_ctx* ctx, const uint8_t* src, asw_mask_image* im) {
    ASW_CUDA(ctx, cudaMemcpy2DAsync(im->data, im->step, src, im->cols, im->cols, im->rows, cudaMemcpyDeviceToHost, ctx->stream));
    return ASW_OK;
}
static asw_status upload_pair(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, uint8_t** dL, uint8_t** dR) {
    size_t n = (size_t)L->rows * L->cols * 3;
    ASW_TRY(ws_get(ctx, WS_IMG_L, n, dL));
    ASW_TRY(ws_get(ctx, WS_IMG_R, n, dR));
    ASW_TRY(upload_u8(ctx, L, *dL));
    ASW_TRY(upload_u8(ctx, R, *dR));
    return ASW_OK;
}
// copy the aggregated volume to the capture buffer (if armed) and disarm
static asw_status finish_capture(asw_ctx* ctx, const float* dev_agg, size_t count) {
    if (!ctx->capture_host) return ASW_OK;
    float* dst = ctx->capture_host;
    size_t cap = ctx->capture_cap;
    ctx->capture_host = nullptr; ctx->capture_cap = 0;
    if (cap < count) return asw_fail(ctx, ASW_ERR_BAD_ARG, "capture buffer too small%s%s");
    ASW_CUDA(ctx, cudaMemcpyAsync(dst, dev_agg, count * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
static asw_status init_keys(asw_ctx* ctx, unsigned long long* keys, size_t n) {
    LAUNCH(ctx, "fill_u64", (k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, WTA_KEY_EMPTY)));
    return ASW_OK;
}
static asw_status keys_to_disp(asw_ctx* ctx, const unsigned long long* keys, size_t n, float* disp) {
    LAUNCH(ctx, "keys_to_disp", (k_keys_to_disp<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(keys, n, disp)));
    return ASW_OK;
}

// =================================================================================================
// device-level pipelines (inputs and outputs resident in HBM)
// =================================================================================================
struct ViewGeom {
    const uint8_t* ref; const uint8_t* tgt;   // tightly packed BGR on device
    int H, W, max_off, pad_l, pad_r, Wp, x0_base, x0_step;
};
// LEFT: pad the right image on the left by max_off, crop at max_off - offset (A.cpp:442, 455)
// RIGHT: pad the left image on the right, crop at offset (A.cpp:491, 503)
static ViewGeom make_view(const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int min_d, int num_d) {
    ViewGeom v;
    v.H = H; v.W = W;
    v.max_off = min_d + num_d - 1;
    if (disp_type == ASW_DISPARITY_LEFT) {
        v.ref = dL; v.tgt = dR; v.pad_l = v.max_off; v.pad_r = 0;
        v.x0_base = v.max_off - min_d; v.x0_step = -1;
    } else {
        v.ref = dR; v.tgt = dL; v.pad_l = 0; v.pad_r = v.max_off;
        v.x0_base = min_d; v.x0_step = 1;
    }
    v.Wp = W + v.max_off;
    return v;
}

__global__ void k_init_slice_mm(uint32_t* mm, int D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) { mm[2 * i] = 0xFFFFFFFFu; mm[2 * i + 1] = 0u; }
}
__global__ void k_init_mm_u8(int* mm) { mm[2 * threadIdx.x] = 255; mm[2 * threadIdx.x + 1] = 0; }   // one pair per thread
// both of the above in one launch (streaming guided path)
__global__ void k_gfs_init(uint32_t* slice_mm, int D, int* guide_mm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) { slice_mm[2 * i] = 0xFFFFFFFFu; slice_mm[2 * i + 1] = 0u; }
    if (i == 0) { guide_mm[0] = 255; guide_mm[1] = 0; }
}

// guidance moments for a C-channel u8 guide: planes I (C), mean_I (C), den (C) + packed records for C == 3
struct GuidePrep { float* I; float* mI; float* den; float4 *Gi, *Gm, *Gd; const int* mm; };
static asw_status prep_guide(asw_ctx* ctx, const uint8_t* guide, int H, int W, int C, int r, double eps, GuidePrep* gp) {
    size_t n = (size_t)H * W;
    float *planes, *boxed;
    int* mm;
    ASW_TRY(ws_get(ctx, WS_TMP0, n * 2 * C, &planes));
    ASW_TRY(ws_get(ctx, WS_TMP1, n * 2 * C, &boxed));
    ASW_TRY(ws_get(ctx, WS_MISC0, (size_t)64, &mm));
    gp->Gi = gp->Gm = gp->Gd = nullptr;
    if (C == 3) {
        ASW_TRY(ws_get(ctx, WS_GUIDE_I, n, &gp->Gi));
        ASW_TRY(ws_get(ctx, WS_GUIDE_MI, n, &gp->Gm));
        ASW_TRY(ws_get(ctx, WS_GUIDE_DEN, n, &gp->Gd));
    }
    LAUNCH(ctx, "init_mm", (k_init_mm_u8<<<1, 1, 0, ctx->stream>>>(mm)));
    LAUNCH(ctx, "minmax_u8", (k_minmax_u8<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(guide, n * C, mm)));
    LAUNCH(ctx, "guide_normalize", (k_guide_normalize<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(guide, n, C, mm, planes, gp->Gi)));
    ASW_TRY(launch_box_f32(ctx, planes, boxed, H, W, r, 2 * C, n));
    LAUNCH(ctx, "guide_finish", (k_guide_finish<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(boxed, n, C, (float)eps, gp->Gm, gp->Gd)));
    gp->I = planes; gp->mI = boxed; gp->den = boxed + n * C; gp->mm = mm;
    return ASW_OK;
}

static size_t gf_smem_bytes(int k) {
    int IW = GF_TW + k - 1, IH = GF_TH + k - 1;
    return ((size_t)IH * (IW | 1) + (size_t)IH * GF_HP) * sizeof(float4);
}
static int gf_chunk_slices(asw_ctx* ctx, size_t n) {
    // keep one chunk's a/b planes (16 B per DE) inside the 126 MB L2 so pass 2 reads them from L2
    size_t budget = (size_t)64 << 20;
    if (ctx->tune[ASW_TUNE_GF_CHUNK_MB] > 0) budget = (size_t)ctx->tune[ASW_TUNE_GF_CHUNK_MB] << 20;
    size_t s = budget / (n * 16);
    if (s < 1) s = 1;
    return (int)s;
}

// GuidedF_2 for one view, fast path.  keys must be initialised by the caller (allows d-range splits).
// d_lo/d_hi: slice index range [d_lo, d_hi) evaluated (labels = min_d + index).
// q' workspace budget of the streaming path (4 B per evaluation) and the slices one chunk of it holds
static const size_t GFS_Q_BUDGET = (size_t)8 << 30;
static int gfs_chunk_slices(int H, int W, int win, int span) {
    const int QW = GFS_IW - 2 * (win - 1);
    const size_t per_slice = (size_t)H * (cdiv(W, QW) * QW);
    return (int)std::min<size_t>((size_t)span, std::max<size_t>(GFS_NS, GFS_Q_BUDGET / (per_slice * 4) / GFS_NS * GFS_NS));
}
static bool gfs_streaming(int H, int W, int win) {
    return gfs_supported(H, W, win) && !asw_dev("ASW_GF_TILED") && !asw_dev("ASW_GF_GENERIC");
}
// disp_direct (optional): when the range is the method's whole range and one chunk holds it, the streaming path's WTA
// pass writes the disparity map itself; *direct_done tells the caller (who then needs neither keys nor keys_to_disp).
static asw_status dev_guidedf2_keys(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type,
                                    double eps, int win, int min_d, int num_d, int d_lo, int d_hi,
                                    unsigned long long* keys, float* agg_dev, float* disp_direct = nullptr) {
    size_t n = (size_t)H * W;
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    Feat *fref, *ftgt;
    ASW_TRY(ws_get(ctx, WS_FEAT_REF, n, &fref));
    ASW_TRY(ws_get(ctx, WS_FEAT_TGT, (size_t)H * v.Wp, &ftgt));
    const bool streaming = gfs_streaming(H, W, win);
    GuidePrep gp;
    if (!streaming) {
        LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(v.ref, H, W, 0, 0, fref)));
        LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, ftgt)));
        ASW_TRY(prep_guide(ctx, v.ref, H, W, 3, win, eps, &gp));                    // A.cpp:3004 / 3019
    }
    uint32_t* slice_mm;
    ASW_TRY(ws_get(ctx, WS_SLICE_MM, (size_t)2 * num_d, &slice_mm));
    if (!streaming) LAUNCH(ctx, "init_slice_mm", (k_init_slice_mm<<<cdiv(num_d, 128), 128, 0, ctx->stream>>>(slice_mm, num_d)));
    TadParams tp = make_tad_params(0.4, 10, 50);                                     // A.cpp:2990
    if (streaming) {
        // streaming kernel (k_guided_stream.cuh): cost, both box levels and q' on chip; 4 B per evaluation to HBM
        FeatF *rff, *tff; GfsMoments* gmom; float2* aff;
        int* gmm;
        ASW_TRY(ws_get(ctx, WS_MISC0, (size_t)64, &gmm));
        ASW_TRY(ws_get(ctx, WS_GUIDE_I, n, &gp.Gi));
        gp.mm = gmm;
        ASW_TRY(ws_get(ctx, WS_FEATF_REF, n, &rff));
        ASW_TRY(ws_get(ctx, WS_FEATF_TGT, (size_t)H * v.Wp, &tff));
        ASW_TRY(ws_get(ctx, WS_GUIDE_NM, n, &gmom));
        ASW_TRY(ws_get(ctx, WS_AFF, (size_t)num_d, &aff));
        // float feature records straight from the images (the target's gradients negated)
        LAUNCH(ctx, "features", (k_features_f<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(v.ref, H, W, 0, 0, 1.0f, rff)));
        LAUNCH(ctx, "features", (k_features_f<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, -1.0f, tff)));
        {
            // guidance: global min / max (cv::normalize is over all channels), then the fused moments kernel
            LAUNCH(ctx, "gfs_init", (k_gfs_init<<<cdiv(num_d, 128), 128, 0, ctx->stream>>>(slice_mm, num_d, gmm)));
            LAUNCH(ctx, "minmax_u8", (k_minmax_u8<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(v.ref, n * 3, gmm)));
            const int strips_g = cdiv(W, GFS_GM_COLS - (win - 1));
            const int bands_g = std::max(1, std::min(cdiv(H, 2 * win), cdiv(4 * ctx->sm_count, strips_g)));
            const int band_rows = cdiv(H, bands_g);
            LAUNCH(ctx, "guide_moments", (k_gfs_guide_moments<<<dim3(strips_g, cdiv(H, band_rows)), GFS_GM_COLS, 0, ctx->stream>>>(
                                             v.ref, gmm, H, W, win, (float)eps, band_rows, gp.Gi, gmom)));
        }
        const int QW = GFS_IW - 2 * (win - 1);
        GfsGeom g;
        g.H = H; g.W = W; g.Wp = v.Wp; g.Wq = cdiv(W, QW) * QW; g.x0_step = v.x0_step; g.D = 0; g.nbands = 2;
        size_t per_slice = (size_t)H * g.Wq;
        int span = d_hi - d_lo;
        int chunk = gfs_chunk_slices(H, W, win, span);
        if (chunk < span) disp_direct = nullptr;
        float* qv;
        ASW_TRY(ws_get(ctx, WS_AB, per_slice * (size_t)chunk, &qv));
        for (int c0 = d_lo; c0 < d_hi; c0 += chunk) {
            int cn = (d_hi - c0 < chunk) ? d_hi - c0 : chunk;
            g.x0_base = v.x0_base + v.x0_step * c0;
            float* agg_c = agg_dev ? agg_dev + (size_t)(c0 - d_lo) * n : nullptr;
            if (win == 9) ASW_TRY(gfs_launch<9>(ctx, rff, tff, gp.Gi, gmom, gp.mm, g, tp, qv, slice_mm + 2 * c0, aff + c0, cn, min_d + c0, keys, agg_c, disp_direct));
            else if (win == 7) ASW_TRY(gfs_launch<7>(ctx, rff, tff, gp.Gi, gmom, gp.mm, g, tp, qv, slice_mm + 2 * c0, aff + c0, cn, min_d + c0, keys, agg_c, disp_direct));
            else ASW_TRY(gfs_launch<5>(ctx, rff, tff, gp.Gi, gmom, gp.mm, g, tp, qv, slice_mm + 2 * c0, aff + c0, cn, min_d + c0, keys, agg_c, disp_direct));
        }
        return ASW_OK;
    }
    // windows 11, 13, 15 (15 = the reference driver's own call, main.cpp:94): the tiled pair of k_guided_fast.cuh (fixed
    // 32 x 64 input tile, (a,b) through HBM).  The streaming kernel's shared-memory layout does not fit them (241 KB at 15 with
    // 2 slices per CTA).  Dev builds can route 5 / 7 / 9 here as well (superseded by the streaming kernel).
    const bool tiled_big = win == 11 || win == 13 || win == 15;
    if ((tiled_big || (asw_dev("ASW_GF_TILED") && (win == 5 || win == 7 || win == 9))) && !asw_dev("ASW_GF_GENERIC")) {
        // tuned kernels (k_guided_fast.cuh): fixed 32x64 input tile, register-resident reference-side data
        const int DC1 = 8;
        float4* grd;
        ASW_TRY(ws_get(ctx, WS_GUIDE_RDEN, n, &grd));
        LAUNCH(ctx, "reciprocal4", (k_reciprocal4<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(gp.Gd, n, grd)));
        size_t budget = (size_t)16 << 30;                 // a/b workspace budget (16 B per disparity evaluation)
        if (ctx->tune[ASW_TUNE_GF_CHUNK_MB] > 0) budget = (size_t)ctx->tune[ASW_TUNE_GF_CHUNK_MB] << 20;
        int span = d_hi - d_lo;
        int chunk = (int)(budget / (n * 16));
        if (chunk < DC1) chunk = DC1;
        chunk = (chunk / DC1) * DC1;
        if (chunk > span) chunk = ((span + DC1 - 1) / DC1) * DC1;
        float4* ab;
        ASW_TRY(ws_get(ctx, WS_AB, n * (size_t)chunk, &ab));
        GffGeom g;
        g.H = H; g.W = W; g.Wp = v.Wp; g.a = win / 2; g.x0_step = v.x0_step; g.D = 0;
        for (int c0 = d_lo; c0 < d_hi; c0 += chunk) {
            int cn = (d_hi - c0 < chunk) ? d_hi - c0 : chunk;
            g.x0_base = v.x0_base + v.x0_step * c0;
            float* agg_c = agg_dev ? agg_dev + (size_t)(c0 - d_lo) * n : nullptr;
            if (win == 15) ASW_TRY(gff_launch<15>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
            else if (win == 13) ASW_TRY(gff_launch<13>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
            else if (win == 11) ASW_TRY(gff_launch<11>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
#ifdef ASW_DEV_KERNELS
            else if (win == 9) ASW_TRY(gff_launch<9>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
            else if (win == 7) ASW_TRY(gff_launch<7>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
            else if (win == 5) ASW_TRY(gff_launch<5>(ctx, fref, ftgt, gp.Gi, gp.Gm, grd, g, tp, ab, slice_mm + 2 * c0, cn, min_d + c0, keys, agg_c));
#endif
        }
        return ASW_OK;
    }
    const int DC = 4;
    int chunk = gf_chunk_slices(ctx, n);
    chunk = ((chunk + DC - 1) / DC) * DC;
    int span = d_hi - d_lo;
    if (chunk > span) chunk = ((span + DC - 1) / DC) * DC;
    float4* ab;
    ASW_TRY(ws_get(ctx, WS_AB, n * (size_t)chunk, &ab));
    size_t smem = gf_smem_bytes(win);
    if (smem > 220 * 1024) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "guided-filter window too large for the tiled kernels%s%s");
    cudaFuncSetAttribute(k_gf_ab<DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_gf_q<DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 tiles(cdiv(W, GF_TW), cdiv(H, GF_TH), 1);
    for (int c0 = d_lo; c0 < d_hi; c0 += chunk) {
        int cn = (d_hi - c0 < chunk) ? d_hi - c0 : chunk;
        GfGeom g;
        g.H = H; g.W = W; g.Wp = v.Wp; g.k = win; g.a = win / 2;
        g.x0_base = v.x0_base + v.x0_step * c0; g.x0_step = v.x0_step; g.D = cn;
        dim3 grid(tiles.x, tiles.y, cdiv(cn, DC));
        LAUNCH(ctx, "gf_ab", (k_gf_ab<DC><<<grid, GF_THREADS, smem, ctx->stream>>>(fref, ftgt, gp.Gi, gp.Gm, gp.Gd, g, tp, ab, slice_mm + 2 * c0)));
        LAUNCH(ctx, "gf_q", (k_gf_q<DC><<<grid, GF_THREADS, smem, ctx->stream>>>(ab, gp.Gi, g, tp.c0, slice_mm + 2 * c0, min_d + c0, keys,
                                                                              agg_dev ? agg_dev + (size_t)(c0 - d_lo) * n : nullptr)));
    }
    return ASW_OK;
}

static asw_status dev_guidedf2(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type,
                               double eps, int win, int min_d, int num_d, float* disp_dev, float* agg_dev) {
    size_t n = (size_t)H * W;
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    if (gfs_streaming(H, W, win) && gfs_chunk_slices(H, W, win, num_d) >= num_d)     // one WTA launch sees every candidate
        return dev_guidedf2_keys(ctx, dL, dR, H, W, disp_type, eps, win, min_d, num_d, 0, num_d, keys, agg_dev, disp_dev);
    ASW_TRY(init_keys(ctx, keys, n));
    ASW_TRY(dev_guidedf2_keys(ctx, dL, dR, H, W, disp_type, eps, win, min_d, num_d, 0, num_d, keys, agg_dev));
    return keys_to_disp(ctx, keys, n, disp_dev);
}

// generic guided filter of one slice: cost (device, n floats) -> q (device); guide prepared by prep_guide
static asw_status dev_gf_generic_slice(asw_ctx* ctx, const GuidePrep& gp, const float* cost, int H, int W, int C, int r, float* q) {
    size_t n = (size_t)H * W;
    float *planes, *boxed;
    uint32_t* mm;
    ASW_TRY(ws_get(ctx, WS_TMP2, n * (1 + C), &planes));
    ASW_TRY(ws_get(ctx, WS_TMP3, n * (1 + C), &boxed));
    ASW_TRY(ws_get(ctx, WS_MISC1, (size_t)64, &mm));
    LAUNCH(ctx, "init_slice_mm", (k_init_slice_mm<<<1, 32, 0, ctx->stream>>>(mm, 1)));
    LAUNCH(ctx, "minmax_f32", (k_minmax_f32_slices<<<dim3(ctx->sm_count * 2, 1), 256, 0, ctx->stream>>>(cost, n, mm)));
    unsigned nb = (unsigned)((n + 255) / 256);
    LAUNCH(ctx, "gfg_products", (k_gfg_products<<<nb, 256, 0, ctx->stream>>>(cost, mm, gp.I, n, C, planes, 0)));
    ASW_TRY(launch_box_f32(ctx, planes, boxed, H, W, r, 1 + C, n));
    LAUNCH(ctx, "gfg_ab", (k_gfg_ab<<<nb, 256, 0, ctx->stream>>>(boxed, gp.mI, gp.den, n, C, 0)));
    ASW_TRY(launch_box_f32(ctx, boxed, planes, H, W, r, 1 + C, n));
    LAUNCH(ctx, "gfg_q", (k_gfg_q<<<nb, 256, 0, ctx->stream>>>(planes, gp.I, n, C, q, 0)));
    return ASW_OK;
}

// The generic guided filter of `ns` consecutive slices in ONE set of 13 launches per chunk (the per-slice form costs 13 launches
// a slice: 839 launches / 12.5 ms for 64 slices of 640 x 360).  Slice s filters cost[s] -> q[s].  merge = true: 6-channel guidance
// merge(ref, tpad cropped at column x0_first + s * x0_step), prepared per slice (GuidedF A.cpp:2905-2912, GuidedF_3 LEFT
// A.cpp:3090-3097); merge = false: the 3-channel image `ref` guides every slice (GuidedF_3 RIGHT, A.cpp:3104).
static asw_status dev_gf_generic_batch(asw_ctx* ctx, bool merge, const uint8_t* ref, const uint8_t* tpad, int H, int W, int Wp,
                                       int ref_first, int x0_first, int x0_step, int ns, int win, double eps, const float* cost, float* q) {
    const size_t n = (size_t)H * W;
    const int C = merge ? 6 : 3;
    const size_t per_slice = n * ((merge ? 6 + 4 * 2 * C * 2 : 0) + 4 * 2 * (1 + C));
    int sb = (int)std::max<size_t>(1, std::min<size_t>((size_t)ns, ((size_t)3 << 30) / per_slice));
    sb = std::min(sb, 1024);                                          // init_mm: one thread per slice
    GuidePrep gp;
    uint8_t* guide6 = nullptr;
    float *gplanes = nullptr, *gboxed = nullptr, *planes, *boxed;
    int* gmm = nullptr; uint32_t* smm;
    if (merge) {
        ASW_TRY(ws_get(ctx, WS_GUIDE6, n * 6 * sb, &guide6));
        ASW_TRY(ws_get(ctx, WS_TMP0, n * 2 * C * sb, &gplanes));
        ASW_TRY(ws_get(ctx, WS_TMP1, n * 2 * C * sb, &gboxed));
        ASW_TRY(ws_get(ctx, WS_MISC0, (size_t)2 * sb + 64, &gmm));
    } else {
        ASW_TRY(prep_guide(ctx, ref, H, W, 3, win, eps, &gp));
    }
    ASW_TRY(ws_get(ctx, WS_TMP2, n * (1 + C) * sb, &planes));
    ASW_TRY(ws_get(ctx, WS_TMP3, n * (1 + C) * sb, &boxed));
    ASW_TRY(ws_get(ctx, WS_MISC1, (size_t)2 * sb + 64, &smm));
    const unsigned nb = (unsigned)((n + 255) / 256);
    for (int s0 = 0; s0 < ns; s0 += sb) {
        const int cn = std::min(sb, ns - s0);
        const float *I, *mI, *den; size_t gstride;
        if (merge) {
            LAUNCH(ctx, "merge_guide6", (k_merge_guide6<<<dim3(cdiv(W, 128), H, cn), 128, 0, ctx->stream>>>(
                                            ref, tpad, H, W, Wp, x0_first + s0 * x0_step, ref_first, guide6, x0_step)));
            LAUNCH(ctx, "init_mm", (k_init_mm_u8<<<1, cn, 0, ctx->stream>>>(gmm)));
            LAUNCH(ctx, "minmax_u8", (k_minmax_u8<<<dim3(std::max(1, ctx->sm_count * 4 / cn), cn), 256, 0, ctx->stream>>>(guide6, n * 6, gmm)));
            LAUNCH(ctx, "guide_normalize", (k_guide_normalize<<<dim3(nb, cn), 256, 0, ctx->stream>>>(guide6, n, C, gmm, gplanes, nullptr)));
            ASW_TRY(launch_box_f32(ctx, gplanes, gboxed, H, W, win, 2 * C * cn, n));
            LAUNCH(ctx, "guide_finish", (k_guide_finish<<<dim3(nb, cn), 256, 0, ctx->stream>>>(gboxed, n, C, (float)eps, nullptr, nullptr)));
            I = gplanes; mI = gboxed; den = gboxed + n * C; gstride = n * 2 * C;
        } else {
            I = gp.I; mI = gp.mI; den = gp.den; gstride = 0;
        }
        const float* c = cost + (size_t)s0 * n;
        LAUNCH(ctx, "init_slice_mm", (k_init_slice_mm<<<cdiv(cn, 128), 128, 0, ctx->stream>>>(smm, cn)));
        LAUNCH(ctx, "minmax_f32", (k_minmax_f32_slices<<<dim3(std::max(1, ctx->sm_count * 2 / cn), cn), 256, 0, ctx->stream>>>(c, n, smm)));
        LAUNCH(ctx, "gfg_products", (k_gfg_products<<<dim3(nb, cn), 256, 0, ctx->stream>>>(c, smm, I, n, C, planes, gstride)));
        ASW_TRY(launch_box_f32(ctx, planes, boxed, H, W, win, (1 + C) * cn, n));
        LAUNCH(ctx, "gfg_ab", (k_gfg_ab<<<dim3(nb, cn), 256, 0, ctx->stream>>>(boxed, mI, den, n, C, gstride)));
        ASW_TRY(launch_box_f32(ctx, boxed, planes, H, W, win, (1 + C) * cn, n));
        LAUNCH(ctx, "gfg_q", (k_gfg_q<<<dim3(nb, cn), 256, 0, ctx->stream>>>(planes, I, n, C, q + (size_t)s0 * n, gstride)));
    }
    return ASW_OK;
}

// gray + SAD box cost volume on device (A.cpp:2442-2503 for every d), vol [num_d][n]
static asw_status dev_cost_sad_box(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type,
                                   int win, int min_d, int num_d, float* vol, uint8_t** gray_ref_out, uint8_t** gray_tgt_out) {
    size_t n = (size_t)H * W;
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    uint8_t *gref, *gtgt;
    ASW_TRY(ws_get(ctx, WS_GRAY_L, n, &gref));
    ASW_TRY(ws_get(ctx, WS_GRAY_R, (size_t)H * v.Wp, &gtgt));
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(v.ref, H, W, 0, 0, gref)));
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(v.Wp, 256), H), 256, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, gtgt)));
    if (win - 1 < SADBOX_COLS / 2) {
        // bands: a band is entered once (win - 1 warm-up rows); pick the count that minimises (waves of the 8 CTAs an SM
        // holds) x (rows a CTA walks)
        const int per_band = cdiv(W, SADBOX_COLS - (win - 1)) * num_d, slots = 8 * ctx->sm_count;
        int bands = 1; long long best_cost = -1;
        for (int b = 1; b <= std::max(1, H / (2 * win)); b++) {
            const int rows = cdiv(cdiv(H, b), 4) * 4;
            const long long cost = (long long)cdiv(per_band * cdiv(H, rows), slots) * (rows + win - 1);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; bands = b; }
        }
        int band_rows = cdiv(cdiv(H, bands), 4) * 4;
        dim3 grid(cdiv(W, SADBOX_COLS - (win - 1)), cdiv(H, band_rows), num_d);
#define SAD_BOX_LAUNCH(WT) LAUNCH(ctx, "sad_box", (k_sad_box_u8<4, WT><<<grid, SADBOX_COLS, 0, ctx->stream>>>(gref, gtgt, H, W, v.Wp, v.x0_base, v.x0_step, win, band_rows, vol)))
        switch (win) {
            case 5: SAD_BOX_LAUNCH(5); break;
            case 7: SAD_BOX_LAUNCH(7); break;
            case 9: SAD_BOX_LAUNCH(9); break;
            case 15: SAD_BOX_LAUNCH(15); break;
            case 25: SAD_BOX_LAUNCH(25); break;
            case 35: SAD_BOX_LAUNCH(35); break;
            default: SAD_BOX_LAUNCH(0); break;
        }
#undef SAD_BOX_LAUNCH
    } else {
        float* ad;
        ASW_TRY(ws_get(ctx, WS_VOL1, n * num_d, &ad));
        LAUNCH(ctx, "gray_absdiff", (k_gray_absdiff<<<dim3(cdiv(W, 256), H, num_d), 256, 0, ctx->stream>>>(gref, gtgt, H, W, v.Wp, v.x0_base, v.x0_step, ad)));
        ASW_TRY(launch_box_f32(ctx, ad, vol, H, W, win, num_d, n));
    }
    if (gray_ref_out) *gray_ref_out = gref;
    if (gray_tgt_out) *gray_tgt_out = gtgt;
    return ASW_OK;
}

// GuidedF (6-channel guidance, SAD cost), A.cpp:2867-2963
static asw_status dev_guidedf(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, double eps,
                              int win, int min_d, int num_d, float* disp_dev, float* agg_dev) {
    size_t n = (size_t)H * W;
    float* cost;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &cost));
    ASW_TRY(dev_cost_sad_box(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, cost, nullptr, nullptr));
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    // padded target BGR image for the 6-channel guide (A.cpp:2877-2878)
    uint8_t* tpad;
    ASW_TRY(ws_get(ctx, WS_MISC2, (size_t)H * v.Wp * 3, &tpad));
    LAUNCH(ctx, "pad_bgr", (k_pad_cols_bgr<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, tpad)));
    float* q = agg_dev;
    if (!q) ASW_TRY(ws_get(ctx, WS_VOL1, n * num_d, &q));   // the |L-R| planes in WS_VOL1 are dead after dev_cost_sad_box
    // slice i: LEFT crops at Rect(numDisparity - i - 1, ..) (A.cpp:2909); RIGHT at Rect(i + minDisparity, ..) (A.cpp:2926)
    const bool left = disp_type == ASW_DISPARITY_LEFT;
    ASW_TRY(dev_gf_generic_batch(ctx, true, v.ref, tpad, H, W, v.Wp, left ? 1 : 0, left ? num_d - 1 : min_d, left ? -1 : 1,
                                 num_d, win, eps, cost, q));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    LAUNCH(ctx, "wta_keys", (k_wta_keys<<<(unsigned)((n / 4 + 256) / 256), 256, 0, ctx->stream>>>(q, num_d, n, min_d, keys)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}

// NCC set-up on device: gray (RGB2GRAY on BGR bytes) reference, padded target, statistics of both
struct NccDev { uint8_t *ref, *tgt; float *mr, *mt; double *sr, *st; ViewGeom v; };
static asw_status dev_ncc_setup(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int win,
                                int min_d, int num_d, NccDev* s) {
    const size_t n = (size_t)H * W;
    s->v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    const size_t nt = (size_t)H * s->v.Wp;
    ASW_TRY(ws_get(ctx, WS_GRAY_L, n, &s->ref));
    ASW_TRY(ws_get(ctx, WS_GRAY_R, nt, &s->tgt));
    ASW_TRY(ws_get(ctx, WS_TMP0, n, &s->mr));
    ASW_TRY(ws_get(ctx, WS_TMP1, nt, &s->mt));
    ASW_TRY(ws_get(ctx, WS_MISC2, n, &s->sr));
    ASW_TRY(ws_get(ctx, WS_MISC3, nt, &s->st));
    LAUNCH(ctx, "rgb2gray", (k_rgb2gray_on_bgr_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(s->v.ref, H, W, 0, 0, s->ref)));
    LAUNCH(ctx, "rgb2gray", (k_rgb2gray_on_bgr_pad<<<dim3(cdiv(s->v.Wp, 256), H), 256, 0, ctx->stream>>>(s->v.tgt, H, W, s->v.pad_l, s->v.pad_r, s->tgt)));
    LAUNCH(ctx, "ncc_stats", (k_ncc_stats<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(s->ref, H, W, win, s->mr, s->sr)));
    LAUNCH(ctx, "ncc_stats", (k_ncc_stats<<<dim3(cdiv(s->v.Wp, 128), H), 128, 0, ctx->stream>>>(s->tgt, H, s->v.Wp, win, s->mt, s->st)));
    return ASW_OK;
}
// computeNCC, vector overload (A.cpp:924-1013): vol [num_d][n], every slice min-max normalised
static asw_status dev_cost_ncc(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int win,
                               int min_d, int num_d, float* vol) {
    const size_t n = (size_t)H * W;
    NccDev s;
    ASW_TRY(dev_ncc_setup(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, &s));
    uint32_t* mm;
    ASW_TRY(ws_get(ctx, WS_SLICE_MM, (size_t)2 * num_d, &mm));
    LAUNCH(ctx, "init_slice_mm", (k_init_slice_mm<<<cdiv(num_d, 128), 128, 0, ctx->stream>>>(mm, num_d)));
    if (ncc_tile_smem(win) <= 160 * 1024 && !asw_dev("ASW_NCC_DIRECT")) {
        cudaFuncSetAttribute(k_ncc_cost_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ncc_tile_smem(win));
        LAUNCH(ctx, "ncc_cost", (k_ncc_cost_tile<false><<<dim3(cdiv(W, NCC_COLS), H, cdiv(num_d, NCC_DC)), NCC_COLS, ncc_tile_smem(win), ctx->stream>>>(
                                    s.ref, s.mr, s.sr, s.tgt, s.mt, s.st, H, W, s.v.Wp, win, s.v.x0_base, s.v.x0_step, num_d, min_d, vol, nullptr)));
    } else {
        LAUNCH(ctx, "ncc_cost", (k_ncc_cost<false><<<dim3(cdiv(W, 128), H, num_d), 128, 0, ctx->stream>>>(
                                    s.ref, s.mr, s.sr, s.tgt, s.mt, s.st, H, W, s.v.Wp, win, s.v.x0_base, s.v.x0_step, 0, min_d, vol, nullptr)));
    }
    LAUNCH(ctx, "minmax_f32", (k_minmax_f32_slices<<<dim3(std::max(1, ctx->sm_count * 2 / num_d), num_d), 256, 0, ctx->stream>>>(vol, n, mm)));
    LAUNCH(ctx, "normalize_slices", (k_normalize_slices<<<dim3((unsigned)((n + 255) / 256), num_d), 256, 0, ctx->stream>>>(vol, n, mm)));
    return ASW_OK;
}
// computeNCC, Mat overload (A.cpp:812-912; the dispatcher's NCC): LEFT scans offsets min .. max - 1 and keeps the MINIMUM
// raw cost; RIGHT never writes its map (cost > DBL_MAX never holds)
static asw_status dev_ncc(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int win, int min_d,
                          int num_d, float* disp_dev) {
    const size_t n = (size_t)H * W;
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    if (disp_type == ASW_DISPARITY_LEFT && num_d > 1) {
        NccDev s;
        ASW_TRY(dev_ncc_setup(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, &s));
        if (ncc_tile_smem(win) <= 160 * 1024 && !asw_dev("ASW_NCC_DIRECT")) {
            cudaFuncSetAttribute(k_ncc_cost_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ncc_tile_smem(win));
            LAUNCH(ctx, "ncc_cost", (k_ncc_cost_tile<true><<<dim3(cdiv(W, NCC_COLS), H, cdiv(num_d - 1, NCC_DC)), NCC_COLS, ncc_tile_smem(win), ctx->stream>>>(
                                        s.ref, s.mr, s.sr, s.tgt, s.mt, s.st, H, W, s.v.Wp, win, s.v.x0_base, s.v.x0_step, num_d - 1, min_d, nullptr, keys)));
        } else {
            LAUNCH(ctx, "ncc_cost", (k_ncc_cost<true><<<dim3(cdiv(W, 128), H, num_d - 1), 128, 0, ctx->stream>>>(
                                        s.ref, s.mr, s.sr, s.tgt, s.mt, s.st, H, W, s.v.Wp, win, s.v.x0_base, s.v.x0_step, 0, min_d, nullptr, keys)));
        }
    }
    return keys_to_disp(ctx, keys, n, disp_dev);
}
// computeAdaptiveWeight_GuidedF_3 (A.cpp:3063-3137): NCC cost; LEFT filters with the 6-channel merge of the left image and the
// shifted right one, RIGHT builds the merge but hands the plain right image to the filter (A.cpp:3104)
static asw_status dev_guidedf3(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, double eps,
                               int win, int min_d, int num_d, float* disp_dev, float* agg_dev) {
    const size_t n = (size_t)H * W;
    float* cost;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &cost));
    ASW_TRY(dev_cost_ncc(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, cost));
    float* q = agg_dev;
    if (!q) ASW_TRY(ws_get(ctx, WS_VOL1, n * num_d, &q));
    if (disp_type == ASW_DISPARITY_LEFT) {
        ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
        uint8_t* tpad;
        ASW_TRY(ws_get(ctx, WS_IMG_PAD, (size_t)H * v.Wp * 3, &tpad));
        LAUNCH(ctx, "pad_bgr", (k_pad_cols_bgr<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, tpad)));
        ASW_TRY(dev_gf_generic_batch(ctx, true, v.ref, tpad, H, W, v.Wp, 1, num_d - 1, -1, num_d, win, eps, cost, q));
    } else {
        ASW_TRY(dev_gf_generic_batch(ctx, false, dR, nullptr, H, W, 0, 0, 0, 0, num_d, win, eps, cost, q));
    }
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    LAUNCH(ctx, "wta_keys", (k_wta_keys<<<(unsigned)((n / 4 + 256) / 256), 256, 0, ctx->stream>>>(q, num_d, n, min_d, keys)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}

// stage 4 on device
static asw_status dev_lr_refine(asw_ctx* ctx, const uint8_t* dL, const float* dl, const float* dr, int H, int W, float tol,
                                int win, double rate_s, double rate_r, uint8_t* valid, float* filled, float* out) {
    dim3 grid(cdiv(W, 128), H);
    double alpha_r = (1.0 / rate_r) * (-1);
    float alpha_s = (float)((1.0 / rate_s) * (-1));
    const size_t ne = (size_t)win * win, n = (size_t)H * W;
    const int bt = (ne * (2 * 64 + 1)) * sizeof(float) <= 96 * 1024 ? 64 : (ne * (2 * 32 + 1)) * sizeof(float) <= 96 * 1024 ? 32 : 0;
    const bool listed = bt && n < ((size_t)1 << 31) && !asw_dev("ASW_REFINE_DENSE");
    if (listed && W <= 48 * 1024) {
        int* list;
        ASW_TRY(ws_get(ctx, WS_REFINE_LIST, n + 1, &list));               // list[0] = count
        ASW_CUDA(ctx, cudaMemsetAsync(list, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, "lr_fill_compact", (k_lr_fill_compact<<<H, 256, W, ctx->stream>>>(dl, dr, H, W, tol, valid, filled, out, list + 1, list)));
    } else {
        LAUNCH(ctx, "lr_check", (k_lr_check<<<grid, 128, 0, ctx->stream>>>(dl, dr, H, W, tol, valid)));
        LAUNCH(ctx, "fill_invalid", (k_fill_invalid<<<grid, 128, 0, ctx->stream>>>(dl, valid, H, W, filled)));
    }
    if (listed) {
        int* list;
        ASW_TRY(ws_get(ctx, WS_REFINE_LIST, n + 1, &list));               // list[0] = count
        if (W > 48 * 1024) {
            ASW_CUDA(ctx, cudaMemsetAsync(list, 0, sizeof(int), ctx->stream));
            LAUNCH(ctx, "refine_compact", (k_refine_compact<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(filled, valid, (int)n, out, list + 1, list)));
        }
        const size_t smem = ne * (2 * bt + 1) * sizeof(float);
        const unsigned blocks = (unsigned)std::min<size_t>((n + bt - 1) / bt, (size_t)ctx->sm_count * 16);
        if (bt == 64) {
            cudaFuncSetAttribute(k_wmedian_refine_list<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            LAUNCH(ctx, "wmedian_refine", (k_wmedian_refine_list<64><<<blocks, 64, smem, ctx->stream>>>(dL, filled, list + 1, list, H, W, win, alpha_r, alpha_s, out)));
        } else {
            cudaFuncSetAttribute(k_wmedian_refine_list<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            LAUNCH(ctx, "wmedian_refine", (k_wmedian_refine_list<32><<<blocks, 32, smem, ctx->stream>>>(dL, filled, list + 1, list, H, W, win, alpha_r, alpha_s, out)));
        }
        return ASW_OK;
    }
    LAUNCH(ctx, "wmedian_refine", (k_wmedian_refine<<<dim3(cdiv(W, 32), H), 32, 0, ctx->stream>>>(dL, filled, valid, H, W, win, alpha_r, alpha_s, out)));
    return ASW_OK;
}

static asw_status dev_guidedf2_lr_refine(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, double eps, int win,
                                         int min_d, int num_d, float tol, double rate_s, double rate_r, float* out,
                                         float** dl_out, float** dr_out, uint8_t** valid_out) {
    size_t n = (size_t)H * W;
    float *dl, *dr, *filled;
    uint8_t* valid;
    ASW_TRY(ws_get(ctx, WS_DISP_L, n, &dl));
    ASW_TRY(ws_get(ctx, WS_DISP_R, n, &dr));
    ASW_TRY(ws_get(ctx, WS_FILLED, n, &filled));
    ASW_TRY(ws_get(ctx, WS_MASK, n, &valid));
    // the two views are independent until the consistency check: the right one runs on the second stream / workspace bank
    // (small frames gain the most: its kernels fill the tails of the left view's; with per-kernel profiling on, every
    // launch synchronises and the views simply alternate)
    ASW_TRY(view_begin(ctx));
    const asw_status st_r = dev_guidedf2(ctx, dL, dR, H, W, ASW_DISPARITY_RIGHT, eps, win, min_d, num_d, dr, nullptr);
    view_switch_back(ctx);
    const asw_status st_l = st_r == ASW_OK ? dev_guidedf2(ctx, dL, dR, H, W, ASW_DISPARITY_LEFT, eps, win, min_d, num_d, dl, nullptr) : st_r;
    ASW_TRY(view_join(ctx));
    ASW_TRY(st_l);
    ASW_TRY(dev_lr_refine(ctx, dL, dl, dr, H, W, tol, win, rate_s, rate_r, valid, filled, out));
    if (dl_out) *dl_out = dl;
    if (dr_out) *dr_out = dr;
    if (valid_out) *valid_out = valid;
    return ASW_OK;
}

#include "k_traditional.cuh"
#include "k_geodesic.cuh"
#include "k_grid.cuh"
#include "k_blo1.cuh"
#include "k_wmedian.cuh"
#include "k_preproc.cuh"
#include "asw_methods.inl"
#include "asw_pool.inl"
