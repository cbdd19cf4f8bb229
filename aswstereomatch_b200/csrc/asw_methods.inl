// asw_methods.inl -- extern "C" entry points (included at the end of asw_lib.cu)

// -------------------------------------------------------------------------------------------------
// stage-level API
// -------------------------------------------------------------------------------------------------
extern "C" asw_status asw_cost_tad_cg(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, float* host_volume,
                                      double regularity, double thres_c, double thres_g, int disp_type,
                                      int min_d, int num_d) {
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (!host_volume || num_d <= 0 || min_d < 0 || (disp_type != 0 && disp_type != 1))
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad cost arguments%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols;
    size_t n = (size_t)H * W;
    uint8_t *dL, *dR;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    Feat *fref, *ftgt;
    float* vol;
    ASW_TRY(ws_get(ctx, WS_FEAT_REF, n, &fref));
    ASW_TRY(ws_get(ctx, WS_FEAT_TGT, (size_t)H * v.Wp, &ftgt));
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &vol));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(v.ref, H, W, 0, 0, fref)));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, ftgt)));
    TadParams tp = make_tad_params(regularity, thres_c, thres_g);
    LAUNCH(ctx, "cost_tad_volume", (k_cost_tad_volume<<<dim3(cdiv(W, 128), H, num_d), 128, 0, ctx->stream>>>(
                                       fref, ftgt, H, W, v.Wp, v.x0_base, v.x0_step, tp, vol)));
    ASW_CUDA(ctx, cudaMemcpyAsync(host_volume, vol, n * num_d * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

// copyMakeBorder(slice, h, h, h, h, BORDER_REFLECT) of every cost slice (A.cpp:651-668): out [D][H + 2h][W + 2h]
__global__ void k_pad_reflect_slices(const float* __restrict__ vol, int H, int W, int h, float* __restrict__ out) {
    const int Wp = W + 2 * h, Hp = H + 2 * h;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y, d = blockIdx.z;
    if (xp >= Wp) return;
    out[((size_t)d * Hp + yp) * Wp + xp] = vol[((size_t)d * H + border_idx(yp - h, H, 0)) * W + border_idx(xp - h, W, 0)];
}
extern "C" asw_status asw_cost_tad_cg_padded(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, float* host_volume,
                                             double regularity, double thres_c, double thres_g, int disp_type,
                                             int win, int min_d, int num_d) {      // the reference's order (A.h:112-114)
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (!host_volume || num_d <= 0 || min_d < 0 || (disp_type != 0 && disp_type != 1) || win <= 0 || win % 2 == 0)
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad cost arguments%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    const int H = L->rows, W = L->cols, h = win / 2;
    const size_t n = (size_t)H * W, np_ = (size_t)(H + 2 * h) * (W + 2 * h);
    uint8_t *dL, *dR;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    Feat *fref, *ftgt;
    float *vol, *padded;
    ASW_TRY(ws_get(ctx, WS_FEAT_REF, n, &fref));
    ASW_TRY(ws_get(ctx, WS_FEAT_TGT, (size_t)H * v.Wp, &ftgt));
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &vol));
    ASW_TRY(ws_get(ctx, WS_VOL1, np_ * num_d, &padded));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(v.ref, H, W, 0, 0, fref)));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, ftgt)));
    TadParams tp = make_tad_params(regularity, thres_c, thres_g);
    LAUNCH(ctx, "cost_tad_volume", (k_cost_tad_volume<<<dim3(cdiv(W, 128), H, num_d), 128, 0, ctx->stream>>>(
                                       fref, ftgt, H, W, v.Wp, v.x0_base, v.x0_step, tp, vol)));
    LAUNCH(ctx, "pad_reflect", (k_pad_reflect_slices<<<dim3(cdiv(W + 2 * h, 128), H + 2 * h, num_d), 128, 0, ctx->stream>>>(vol, H, W, h, padded)));
    ASW_CUDA(ctx, cudaMemcpyAsync(host_volume, padded, np_ * num_d * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

extern "C" asw_status asw_cost_sad_box(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, float* host_volume,
                                       int disp_type, int win, int min_d, int num_d) {
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (!host_volume || num_d <= 0 || min_d < 0 || (disp_type != 0 && disp_type != 1))
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad cost arguments%s%s");
    if (win <= 0 || win % 2 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "winsize must be odd%s%s");   // A.cpp:2458-2462
    if (min_d + num_d - 1 <= 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "padded image not wider than the image%s%s"); // A.cpp:2472
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols;
    size_t n = (size_t)H * W;
    uint8_t *dL, *dR;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    float* vol;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &vol));
    ASW_TRY(dev_cost_sad_box(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, vol, nullptr, nullptr));
    ASW_CUDA(ctx, cudaMemcpyAsync(host_volume, vol, n * num_d * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

extern "C" asw_status asw_wta(asw_ctx* ctx, const float* host_volume, int D, int rows, int cols, int min_d,
                              asw_f32_image* disp) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    if (!host_volume || D <= 0 || rows <= 0 || cols <= 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad volume%s%s");
    ASW_TRY(check_f32(ctx, disp));
    if (disp->rows != rows || disp->cols != cols) return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "map size differs from the volume%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)rows * cols;
    float *vol, *dd;
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * D, &vol));
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(ws_get(ctx, WS_OUT, n, &dd));
    ASW_CUDA(ctx, cudaMemcpyAsync(vol, host_volume, n * D * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ASW_TRY(init_keys(ctx, keys, n));
    LAUNCH(ctx, "wta_keys", (k_wta_keys<<<(unsigned)((n / 4 + 256) / 256), 256, 0, ctx->stream>>>(vol, D, n, min_d, keys)));
    ASW_TRY(keys_to_disp(ctx, keys, n, dd));
    ASW_TRY(download_f32(ctx, dd, disp));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

extern "C" asw_status asw_guided_filter(asw_ctx* ctx, const asw_u8_image* guide, const asw_f32_image* p, int r, double eps,
                                        asw_f32_image* out) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    if (!guide || (guide->channels != 1 && guide->channels != 3 && guide->channels != 6))
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "guide must have 1, 3 or 6 channels%s%s");
    ASW_TRY(check_u8(ctx, guide, guide->channels));
    ASW_TRY(check_f32(ctx, p));
    ASW_TRY(check_f32(ctx, out));
    if (guide->rows != p->rows || guide->cols != p->cols || out->rows != p->rows || out->cols != p->cols)
        return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "guide / input / output sizes differ%s%s");    // A.cpp:2768-2769
    if (r <= 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad window%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = p->rows, W = p->cols, C = guide->channels;
    size_t n = (size_t)H * W;
    uint8_t* dg;
    float *dp, *dq;
    ASW_TRY(ws_get(ctx, WS_IMG_L, n * C, &dg));
    ASW_TRY(ws_get(ctx, WS_VOL0, n, &dp));
    ASW_TRY(ws_get(ctx, WS_OUT, n, &dq));
    ASW_TRY(upload_u8(ctx, guide, dg));
    ASW_TRY(upload_f32(ctx, p, dp));
    GuidePrep gp;
    ASW_TRY(prep_guide(ctx, dg, H, W, C, r, eps, &gp));
    ASW_TRY(dev_gf_generic_slice(ctx, gp, dp, H, W, C, r, dq));
    ASW_TRY(download_f32(ctx, dq, out));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

// -------------------------------------------------------------------------------------------------
// stage 4
// -------------------------------------------------------------------------------------------------
extern "C" asw_status asw_lr_check(asw_ctx* ctx, const asw_f32_image* dl, const asw_f32_image* dr, float tol,
                                   asw_mask_image* valid) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_f32(ctx, dl)); ASW_TRY(check_f32(ctx, dr)); ASW_TRY(check_mask(ctx, valid));
    if (dl->rows != dr->rows || dl->cols != dr->cols || valid->rows != dl->rows || valid->cols != dl->cols)
        return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "map sizes differ%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = dl->rows, W = dl->cols; size_t n = (size_t)H * W;
    float *a, *b; uint8_t* m;
    ASW_TRY(ws_get(ctx, WS_DISP_L, n, &a)); ASW_TRY(ws_get(ctx, WS_DISP_R, n, &b)); ASW_TRY(ws_get(ctx, WS_MASK, n, &m));
    ASW_TRY(upload_f32(ctx, dl, a)); ASW_TRY(upload_f32(ctx, dr, b));
    LAUNCH(ctx, "lr_check", (k_lr_check<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(a, b, H, W, tol, m)));
    ASW_TRY(download_mask(ctx, m, valid));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_fill_invalid(asw_ctx* ctx, const asw_f32_image* d, const asw_mask_image* valid, asw_f32_image* out) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_f32(ctx, d)); ASW_TRY(check_mask(ctx, valid)); ASW_TRY(check_f32(ctx, out));
    if (valid->rows != d->rows || valid->cols != d->cols || out->rows != d->rows || out->cols != d->cols)
        return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "map sizes differ%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = d->rows, W = d->cols; size_t n = (size_t)H * W;
    float *a, *o; uint8_t* m;
    ASW_TRY(ws_get(ctx, WS_DISP_L, n, &a)); ASW_TRY(ws_get(ctx, WS_FILLED, n, &o)); ASW_TRY(ws_get(ctx, WS_MASK, n, &m));
    ASW_TRY(upload_f32(ctx, d, a)); ASW_TRY(upload_mask(ctx, valid, m));
    LAUNCH(ctx, "fill_invalid", (k_fill_invalid<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(a, m, H, W, o)));
    ASW_TRY(download_f32(ctx, o, out));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_wmedian_refine(asw_ctx* ctx, const asw_u8_image* img, const asw_f32_image* filled,
                                         const asw_mask_image* valid, int win, double rate_s, double rate_r,
                                         asw_f32_image* out) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_u8(ctx, img, 3)); ASW_TRY(check_f32(ctx, filled)); ASW_TRY(check_mask(ctx, valid)); ASW_TRY(check_f32(ctx, out));
    if (win <= 0 || win % 2 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "window must be odd%s%s");     // A.cpp:3238
    if (filled->rows != img->rows || filled->cols != img->cols || valid->rows != img->rows || valid->cols != img->cols ||
        out->rows != img->rows || out->cols != img->cols)
        return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "sizes differ%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = img->rows, W = img->cols; size_t n = (size_t)H * W;
    uint8_t *di, *m; float *f, *o;
    ASW_TRY(ws_get(ctx, WS_IMG_L, n * 3, &di)); ASW_TRY(ws_get(ctx, WS_MASK, n, &m));
    ASW_TRY(ws_get(ctx, WS_FILLED, n, &f)); ASW_TRY(ws_get(ctx, WS_OUT, n, &o));
    ASW_TRY(upload_u8(ctx, img, di)); ASW_TRY(upload_f32(ctx, filled, f)); ASW_TRY(upload_mask(ctx, valid, m));
    double alpha_r = (1.0 / rate_r) * (-1);
    float alpha_s = (float)((1.0 / rate_s) * (-1));
    LAUNCH(ctx, "wmedian_refine", (k_wmedian_refine<<<dim3(cdiv(W, 32), H), 32, 0, ctx->stream>>>(di, f, m, H, W, win, alpha_r, alpha_s, o)));
    ASW_TRY(download_f32(ctx, o, out));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

// driver post-processing on the device (aswStereoMatch.cpp:97-98): the 8-bit, min-max stretched map the driver writes
extern "C" asw_status asw_disparity_to_u8(asw_ctx* ctx, const asw_f32_image* disp, asw_mask_image* out) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_f32(ctx, disp)); ASW_TRY(check_mask(ctx, out));
    if (out->rows != disp->rows || out->cols != disp->cols) return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "map sizes differ%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)disp->rows * disp->cols;
    float* d; uint8_t* o; int* mm;
    ASW_TRY(ws_get(ctx, WS_DISP_L, n, &d)); ASW_TRY(ws_get(ctx, WS_MASK, n, &o)); ASW_TRY(ws_get(ctx, WS_MISC0, (size_t)64, &mm));
    ASW_TRY(upload_f32(ctx, disp, d));
    LAUNCH(ctx, "init_mm", (k_init_mm_u8<<<1, 1, 0, ctx->stream>>>(mm)));
    LAUNCH(ctx, "disp_round_minmax", (k_disp_round_minmax<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d, n, o, mm)));
    LAUNCH(ctx, "disp_normalize_u8", (k_disp_normalize_u8<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(o, n, mm)));
    ASW_TRY(download_mask(ctx, o, out));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

// driver pre-processing (aswStereoMatch.cpp:30-31, 67-89): resize to dst's size + V-channel bilateral detail boost
extern "C" asw_status asw_preprocess(asw_ctx* ctx, const asw_u8_image* src, asw_u8_image* dst) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_u8(ctx, src, 3)); ASW_TRY(check_u8(ctx, dst, 3));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *raw, *out;
    ASW_TRY(ws_get(ctx, WS_IMG_L, (size_t)src->rows * src->cols * 3, &raw));
    ASW_TRY(ws_get(ctx, WS_IMG_R, (size_t)dst->rows * dst->cols * 3, &out));
    ASW_TRY(upload_u8(ctx, src, raw));
    ASW_TRY(dev_preprocess(ctx, raw, src->rows, src->cols, out, dst->rows, dst->cols));
    const size_t rowb = (size_t)dst->cols * 3;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(const_cast<uint8_t*>(dst->data), dst->step, out, rowb, rowb, dst->rows, cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

static bool valid_disp_args(int disp_type, int min_d, int num_d) {
    return (disp_type == 0 || disp_type == 1) && min_d >= 0 && num_d > 0;
}

extern "C" asw_status asw_guidedf2_lr_refine(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* refined,
                                             double eps, int win, int min_d, int num_d, float tol, double rate_s, double rate_r,
                                             asw_f32_image* raw_left, asw_f32_image* raw_right, asw_mask_image* valid) {
    ASW_TRY(check_pair(ctx, L, R, refined));
    if (!valid_disp_args(0, min_d, num_d) || win <= 0 || win % 2 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad arguments%s%s");
    if (raw_left) ASW_TRY(check_f32(ctx, raw_left));
    if (raw_right) ASW_TRY(check_f32(ctx, raw_right));
    if (valid) ASW_TRY(check_mask(ctx, valid));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols; size_t n = (size_t)H * W;
    uint8_t *dL, *dR, *dv; float *out, *dl, *dr;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    ASW_TRY(ws_get(ctx, WS_OUT, n, &out));
    ASW_TRY(dev_guidedf2_lr_refine(ctx, dL, dR, H, W, eps, win, min_d, num_d, tol, rate_s, rate_r, out, &dl, &dr, &dv));
    ASW_TRY(download_f32(ctx, out, refined));
    if (raw_left) ASW_TRY(download_f32(ctx, dl, raw_left));
    if (raw_right) ASW_TRY(download_f32(ctx, dr, raw_right));
    if (valid) ASW_TRY(download_mask(ctx, dv, valid));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

// -------------------------------------------------------------------------------------------------
// per-method entry points
// -------------------------------------------------------------------------------------------------
enum MethodId { M_TRAD, M_GEO, M_GRID, M_BLO1, M_GF1, M_GF2, M_WMED, M_D8, M_GF3, M_NCC };
struct MethodArgs {
    int id, disp_type, win, min_d, num_d;
    double p0, p1;   // method-specific: (gamma_c, gamma_g) | (rate_s, rate_r) | (rate_r) | (eps)
};
static int method_n_eval(const MethodArgs& m) {
    return (m.id == M_TRAD || m.id == M_GEO || m.id == M_GRID || m.id == M_D8) ? m.num_d + 1 : m.num_d;
}
// argument checks mirroring the reference's early-outs (SURVEY 8b "errors")
static asw_status method_check(asw_ctx* ctx, const MethodArgs& m) {
    if ((m.disp_type != 0 && m.disp_type != 1) || m.min_d < 0 || m.num_d <= 0)
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad disparity arguments%s%s");
    bool needs_odd = m.id != M_GRID;
    if (needs_odd && (m.win <= 0 || m.win % 2 == 0))
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "window size must be odd%s%s");           // A.cpp:1440, 2458, 3238
    // DISPARITY_RIGHT: the TAD-cost methods throw in the reference (Appendix A-3; observed on the reference's own code,
    // tests/test_cpu_ref.py), the grid reads out of bounds (A-7).  GuidedF_2 RIGHT is offered only through the explicit
    // LR pipeline (asw_guidedf2_lr_refine), with the mirrored-LEFT cost.  BLO(1), traditional, geodesic and GuidedF
    // define RIGHT (A.cpp:2538-2546, 2600-2631, 2685-2722 for BLO(1)).
    // The 8-direction method's RIGHT branch indexes its weight lists out of bounds (A.cpp:1291: [i] for [count]).
    if (m.disp_type == 1 && (m.id == M_GRID || m.id == M_WMED || m.id == M_D8))
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "DISPARITY_RIGHT is undefined for this method in the reference%s%s");
    if (m.id == M_BLO1 && m.min_d != 0)
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "BLO1 indexes planes by offset (A.cpp:2666): minDisparity must be 0%s%s");
    return ASW_OK;
}
static asw_status dev_run_method(asw_ctx* ctx, const MethodArgs& m, const uint8_t* dL, const uint8_t* dR, int H, int W,
                                 float* disp_dev, float* agg_dev) {
    switch (m.id) {
    case M_GF2: return dev_guidedf2(ctx, dL, dR, H, W, m.disp_type, m.p0, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_GF1: return dev_guidedf(ctx, dL, dR, H, W, m.disp_type, m.p0, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_GF3: return dev_guidedf3(ctx, dL, dR, H, W, m.disp_type, m.p0, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_NCC: return dev_ncc(ctx, dL, dR, H, W, m.disp_type, m.win, m.min_d, m.num_d, disp_dev);
    case M_TRAD: return dev_traditional(ctx, dL, dR, H, W, m.p0, m.p1, m.disp_type, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_D8:    // gamma_c = 30, gamma_g = win * 2 / 3 in integer arithmetic (A.cpp:1175)
        if (m.win * 2 / 3 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "8-direction ASW: window too small (gamma_g = 0)%s%s");
        return dev_traditional(ctx, dL, dR, H, W, 30.0, (double)(m.win * 2 / 3), m.disp_type, m.win, m.min_d, m.num_d, disp_dev, agg_dev, 1);
    case M_GEO: return dev_geodesic(ctx, dL, dR, H, W, m.disp_type, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_GRID: return dev_bilateral_grid(ctx, dL, dR, H, W, m.p0, m.p1, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_BLO1: return dev_blo1(ctx, dL, dR, H, W, m.disp_type, m.p0, m.win, m.min_d, m.num_d, disp_dev, agg_dev);
    case M_WMED: return dev_weighted_median(ctx, dL, dR, H, W, m.win, m.p0, m.p1, m.min_d, m.num_d, disp_dev, agg_dev);
    }
    return ASW_ERR_UNSUPPORTED;
}
static asw_status run_method_host(asw_ctx* ctx, const MethodArgs& m, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp) {
    ASW_TRY(check_pair(ctx, L, R, disp));
    ASW_TRY(method_check(ctx, m));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols; size_t n = (size_t)H * W;
    uint8_t *dL, *dR; float* out; float* agg = nullptr;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    ASW_TRY(ws_get(ctx, WS_OUT, n, &out));
    size_t agg_count = n * (size_t)method_n_eval(m);
    if (ctx->capture_host) ASW_TRY(ws_get(ctx, WS_CAPTURE, agg_count, &agg));
    asw_status st = dev_run_method(ctx, m, dL, dR, H, W, out, agg);
    if (st != ASW_OK) { ctx->capture_host = nullptr; return st; }
    ASW_TRY(download_f32(ctx, out, disp));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (agg) ASW_TRY(finish_capture(ctx, agg, agg_count));
    return ASW_OK;
}

extern "C" asw_status asw_adaptive_weight(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                          double gamma_c, double gamma_g, int disp_type, int win, int min_d, int num_d) {
    MethodArgs m = {M_TRAD, disp_type, win, min_d, num_d, gamma_c, gamma_g};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_direct8(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                  int disp_type, int win, int min_d, int num_d) {
    MethodArgs m = {M_D8, disp_type, win, min_d, num_d, 0, 0};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_geodesic(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                   int disp_type, int win, int min_d, int num_d) {
    MethodArgs m = {M_GEO, disp_type, win, min_d, num_d, 0, 0};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_bilateral_grid(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                         int disp_type, double rate_s, double rate_r, int min_d, int num_d) {
    MethodArgs m = {M_GRID, disp_type, 1, min_d, num_d, rate_s, rate_r};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_blo1(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                               int disp_type, double rate_r, int win, int min_d, int num_d) {
    MethodArgs m = {M_BLO1, disp_type, win, min_d, num_d, rate_r, 0};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_guidedf(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                  int disp_type, double eps, int win, int min_d, int num_d) {
    MethodArgs m = {M_GF1, disp_type, win, min_d, num_d, eps, 0};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_guidedf_3(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                    int disp_type, double eps, int win, int min_d, int num_d) {
    MethodArgs m = {M_GF3, disp_type, win, min_d, num_d, eps, 0};
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_ncc(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp, int disp_type, int win,
                              int min_d, int num_d) {
    MethodArgs m = {M_NCC, disp_type, win, min_d, num_d, 0, 0};
    if (ctx && ctx->capture_host) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "the Mat-returning computeNCC has no cost volume; use asw_cost_ncc%s%s");
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_cost_ncc(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, float* host_volume, int disp_type,
                                   int win, int min_d, int num_d) {
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (!host_volume || num_d <= 0 || min_d < 0 || (disp_type != 0 && disp_type != 1) || win <= 0 || win % 2 == 0)
        return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad cost arguments%s%s");                                  // A.cpp:939-942
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols;
    size_t n = (size_t)H * W;
    uint8_t *dL, *dR;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    float* vol;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &vol));
    ASW_TRY(dev_cost_ncc(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, vol));
    ASW_CUDA(ctx, cudaMemcpyAsync(host_volume, vol, n * num_d * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_adaptive_weight_guidedf_2(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                    int disp_type, double eps, int win, int min_d, int num_d) {
    MethodArgs m = {M_GF2, disp_type, win, min_d, num_d, eps, 0};
    // the reference's RIGHT cost throws (Appendix A-3); the per-method entry keeps that contract
    if (ctx && disp_type == 1) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "GuidedF_2 DISPARITY_RIGHT throws in the reference; use asw_guidedf2_lr_refine%s%s");
    return run_method_host(ctx, m, L, R, disp);
}
extern "C" asw_status asw_adaptive_weight_weighted_median(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                          int disp_type, int win, double rate_s, double rate_r, int min_d, int num_d) {
    MethodArgs m = {M_WMED, disp_type, win, min_d, num_d, rate_s, rate_r};
    return run_method_host(ctx, m, L, R, disp);
}

// dispatcher with the reference's literals (A.cpp:46-88)
static bool dispatcher_args(int algorithm, int disp_type, int win, int min_d, int num_d, MethodArgs* m) {
    switch (algorithm) {
    case ASW_ALG_ADAPTIVE_WEIGHT: *m = {M_TRAD, disp_type, win, min_d, num_d, 30, 20}; return true;              // A.cpp:58
    case ASW_ALG_ADAPTIVE_WEIGHT_8DIRECT: *m = {M_D8, disp_type, win, min_d, num_d, 0, 0}; return true;           // A.cpp:61
    case ASW_ALG_ADAPTIVE_WEIGHT_GEODESIC: *m = {M_GEO, disp_type, win, min_d, num_d, 0, 0}; return true;         // A.cpp:64
    case ASW_ALG_ADAPTIVE_WEIGHT_BILATERAL_GRID: *m = {M_GRID, disp_type, 1, min_d, num_d, 10, 10}; return true;  // A.cpp:67
    case ASW_ALG_ADAPTIVE_WEIGHT_BLO1: *m = {M_BLO1, disp_type, win, min_d, num_d, 0.015, 0}; return true;        // A.cpp:70
    case ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER: *m = {M_GF1, disp_type, win, min_d, num_d, 1e-6, 0}; return true; // A.cpp:73
    case ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER_2: *m = {M_GF2, disp_type, win, min_d, num_d, 1e-6, 0}; return true; // A.cpp:76
    case ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER_3: *m = {M_GF3, disp_type, win, min_d, num_d, 1e-6, 0}; return true; // A.cpp:79
    case ASW_ALG_NCC: *m = {M_NCC, disp_type, win, min_d, num_d, 0, 0}; return true;                              // A.cpp:85
    case ASW_ALG_ADAPTIVE_WEIGHT_MEDIAN: *m = {M_WMED, disp_type, win, min_d, num_d, 10, 10}; return true;        // A.cpp:82
    default: return false;   // BM, SGBM: OpenCV's own matchers, outside the hot path
    }
}
// candidates the dispatcher's method scans for a named numDisparity (D + 1 where the reference loop runs to
// max_offset inclusive, A.cpp:1074, 1467, 2279); -1 for algorithms outside the hot path
extern "C" int asw_method_candidates(int algorithm, int num_d) {
    MethodArgs m;
    if (num_d <= 0 || !dispatcher_args(algorithm, 0, 1, 0, num_d, &m)) return -1;
    return method_n_eval(m);
}
extern "C" asw_status asw_stereo_matching(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                          int disp_type, int algorithm, int win, int min_d, int num_d) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    MethodArgs m;
    if (!dispatcher_args(algorithm, disp_type, win, min_d, num_d, &m))
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "algorithm outside the dense-matching hot path%s%s");
    if (m.id == M_GF2 && disp_type == 1) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "GuidedF_2 DISPARITY_RIGHT throws in the reference%s%s");
    return run_method_host(ctx, m, L, R, disp);
}

// -------------------------------------------------------------------------------------------------
// device-resident batches
// -------------------------------------------------------------------------------------------------
// Transfers run on their own streams (H2D and D2H separately: PCIe is full duplex) and are ordered against the
// compute stream per pair: uploaded[i] -> compute of pair i -> computed[i] -> download of pair i -> downloaded[i]
// -> next compute of pair i; next upload of pair i waits for computed[i].  Uploading pair i+1 and downloading
// pair i-1 therefore overlap the kernels of pair i.
struct asw_batch {
    asw_ctx* ctx; int n, H, W;
    int active;        // the run calls process pairs [0, active)
    uint8_t* imgs;     // [n][2][H][W][3]
    float* disp;       // [n][H][W]
    std::vector<cudaEvent_t> uploaded, computed, downloaded;
    // asw_batch_upload_raw: staging for one raw pair (any size) + the event after which it may be overwritten
    uint8_t* raw = nullptr; size_t raw_cap = 0; cudaEvent_t raw_free = nullptr;
};
static asw_status batch_pair_begin(asw_batch* b, int i) {
    asw_ctx* ctx = b->ctx;
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->uploaded[i], 0));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->downloaded[i], 0));
    return ASW_OK;
}
static asw_status batch_pair_end(asw_batch* b, int i) {
    ASW_CUDA(b->ctx, cudaEventRecord(b->computed[i], b->ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_batch_create(asw_ctx* ctx, int n_pairs, int rows, int cols, asw_batch** out) {
    if (!ctx || !out || n_pairs <= 0 || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    asw_batch* b = new asw_batch();
    b->ctx = ctx; b->n = n_pairs; b->active = n_pairs; b->H = rows; b->W = cols; b->imgs = nullptr; b->disp = nullptr;
    size_t n = (size_t)rows * cols;
    if (cudaMalloc(&b->imgs, n * 6 * n_pairs) != cudaSuccess || cudaMalloc(&b->disp, n * 4 * n_pairs) != cudaSuccess) {
        if (b->imgs) cudaFree(b->imgs);
        delete b;
        return asw_fail(ctx, ASW_ERR_NOMEM, "batch allocation failed%s%s");
    }
    b->uploaded.resize(n_pairs); b->computed.resize(n_pairs); b->downloaded.resize(n_pairs);
    for (int i = 0; i < n_pairs; i++) {
        cudaEventCreateWithFlags(&b->uploaded[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&b->computed[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&b->downloaded[i], cudaEventDisableTiming);
    }
    *out = b;
    return ASW_OK;
}
extern "C" void asw_batch_destroy(asw_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->h2d_stream);
    cudaStreamSynchronize(b->ctx->stream);
    cudaStreamSynchronize(b->ctx->d2h_stream);
    for (size_t i = 0; i < b->uploaded.size(); i++) {
        cudaEventDestroy(b->uploaded[i]); cudaEventDestroy(b->computed[i]); cudaEventDestroy(b->downloaded[i]);
    }
    cudaFree(b->imgs); cudaFree(b->disp);
    if (b->raw) cudaFree(b->raw);
    if (b->raw_free) cudaEventDestroy(b->raw_free);
    delete b;
}
extern "C" asw_status asw_batch_set_active(asw_batch* b, int n_active) {
    if (!b || n_active < 0 || n_active > b->n) return ASW_ERR_BAD_ARG;
    b->active = n_active;
    return ASW_OK;
}
extern "C" asw_status asw_batch_upload(asw_batch* b, int i, const asw_u8_image* L, const asw_u8_image* R) {
    if (!b || i < 0 || i >= b->n) return ASW_ERR_BAD_ARG;
    asw_ctx* ctx = b->ctx;
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (L->rows != b->H || L->cols != b->W) return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "pair size differs from the batch%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n3 = (size_t)b->H * b->W * 3;
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d_stream, b->computed[i], 0));     // the previous run still reads this pair
    size_t rowb = (size_t)b->W * 3;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(b->imgs + (size_t)i * 2 * n3, rowb, L->data, L->step, rowb, b->H, cudaMemcpyHostToDevice, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaMemcpy2DAsync(b->imgs + (size_t)i * 2 * n3 + n3, rowb, R->data, R->step, rowb, b->H, cudaMemcpyHostToDevice, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaEventRecord(b->uploaded[i], ctx->h2d_stream));
    return ASW_OK;
}
// raw frames (any size, CV_8UC3) -> the driver's pre-processing on the device -> the batch slot.  The raw pair is staged in
// one buffer per batch: the next raw upload waits until the previous pre-processing has read it.
extern "C" asw_status asw_batch_upload_raw(asw_batch* b, int i, const asw_u8_image* L, const asw_u8_image* R) {
    if (!b || i < 0 || i >= b->n) return ASW_ERR_BAD_ARG;
    asw_ctx* ctx = b->ctx;
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t raw3 = (size_t)L->rows * L->cols * 3, n3 = (size_t)b->H * b->W * 3;
    if (!b->raw_free) ASW_CUDA(ctx, cudaEventCreateWithFlags(&b->raw_free, cudaEventDisableTiming));
    if (b->raw_cap < 2 * raw3) {
        ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ASW_CUDA(ctx, cudaStreamSynchronize(ctx->h2d_stream));
        if (b->raw) cudaFree(b->raw);
        b->raw = nullptr; b->raw_cap = 0;
        if (cudaMalloc(&b->raw, 2 * raw3) != cudaSuccess) return asw_fail(ctx, ASW_ERR_NOMEM, "raw staging allocation failed%s%s");
        b->raw_cap = 2 * raw3;
    }
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d_stream, b->raw_free, 0));          // the previous pair's pre-processing
    const size_t rowb = (size_t)L->cols * 3;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(b->raw, rowb, L->data, L->step, rowb, L->rows, cudaMemcpyHostToDevice, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaMemcpy2DAsync(b->raw + raw3, rowb, R->data, R->step, rowb, R->rows, cudaMemcpyHostToDevice, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->h2d_stream));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->computed[i], 0));           // ordered anyway (same stream); explicit
    ASW_TRY(dev_preprocess(ctx, b->raw, L->rows, L->cols, b->imgs + (size_t)i * 2 * n3, b->H, b->W));
    ASW_TRY(dev_preprocess(ctx, b->raw + raw3, R->rows, R->cols, b->imgs + (size_t)i * 2 * n3 + n3, b->H, b->W));
    ASW_CUDA(ctx, cudaEventRecord(b->raw_free, ctx->stream));
    ASW_CUDA(ctx, cudaEventRecord(b->uploaded[i], ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_batch_run_guidedf2_lr_refine(asw_batch* b, double eps, int win, int min_d, int num_d, float tol,
                                                       double rate_s, double rate_r) {
    if (!b) return ASW_ERR_BAD_ARG;
    asw_ctx* ctx = b->ctx;
    if (!valid_disp_args(0, min_d, num_d) || win <= 0 || win % 2 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad arguments%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)b->H * b->W, n3 = n * 3;
    for (int i = 0; i < b->active; i++) {
        const uint8_t* dL = b->imgs + (size_t)i * 2 * n3;
        ASW_TRY(batch_pair_begin(b, i));
        ASW_TRY(dev_guidedf2_lr_refine(ctx, dL, dL + n3, b->H, b->W, eps, win, min_d, num_d, tol, rate_s, rate_r,
                                       b->disp + (size_t)i * n, nullptr, nullptr, nullptr));
        ASW_TRY(batch_pair_end(b, i));
    }
    return ASW_OK;
}
extern "C" asw_status asw_batch_run_method(asw_batch* b, int algorithm, int disp_type, int win, int min_d, int num_d) {
    if (!b) return ASW_ERR_BAD_ARG;
    asw_ctx* ctx = b->ctx;
    MethodArgs m;
    if (!dispatcher_args(algorithm, disp_type, win, min_d, num_d, &m))
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "algorithm outside the dense-matching hot path%s%s");
    ASW_TRY(method_check(ctx, m));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)b->H * b->W, n3 = n * 3;
    for (int i = 0; i < b->active; i++) {
        const uint8_t* dL = b->imgs + (size_t)i * 2 * n3;
        ASW_TRY(batch_pair_begin(b, i));
        ASW_TRY(dev_run_method(ctx, m, dL, dL + n3, b->H, b->W, b->disp + (size_t)i * n, nullptr));
        ASW_TRY(batch_pair_end(b, i));
    }
    return ASW_OK;
}
extern "C" asw_status asw_batch_download(asw_batch* b, int i, asw_f32_image* disp) {
    if (!b || i < 0 || i >= b->n) return ASW_ERR_BAD_ARG;
    asw_ctx* ctx = b->ctx;
    ASW_TRY(check_f32(ctx, disp));
    if (disp->rows != b->H || disp->cols != b->W) return asw_fail(ctx, ASW_ERR_SIZE_MISMATCH, "map size differs from the batch%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    ASW_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, b->computed[i], 0));
    size_t rowb = (size_t)b->W * 4;
    ASW_CUDA(ctx, cudaMemcpy2DAsync(disp->data, disp->step, b->disp + (size_t)i * b->H * b->W, rowb, rowb, b->H, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    ASW_CUDA(ctx, cudaEventRecord(b->downloaded[i], ctx->d2h_stream));
    return ASW_OK;
}

// -------------------------------------------------------------------------------------------------
// disparity-range split (multi-GPU): local keys for [d_begin, d_end), merge, finalise
// -------------------------------------------------------------------------------------------------
// device part of the split: keys of candidates [d_begin, d_end) left in WS_KEYS (asynchronous on the ctx stream)
static asw_status split_local_keys_dev(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, int algorithm,
                                       int disp_type, int win, int min_d, int num_d, int d_begin, int d_end,
                                       void** device_keys) {
    ASW_TRY(check_pair(ctx, L, R, nullptr));
    if (!device_keys) return ASW_ERR_BAD_ARG;
    MethodArgs m;
    if (!dispatcher_args(algorithm, disp_type, win, min_d, num_d, &m))
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "algorithm outside the dense-matching hot path%s%s");
    ASW_TRY(method_check(ctx, m));
    // [d_begin, d_end) indexes the candidates the method scans: num_d of them, or num_d + 1 for the methods whose
    // reference loop runs to max_offset inclusive (traditional, geodesic, grid)
    if (m.id != M_GF2 && m.id != M_TRAD && m.id != M_GEO && m.id != M_GRID && m.id != M_BLO1 && m.id != M_D8)
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "disparity split is not implemented for this method%s%s");
    if (m.id == M_GF2 && disp_type == 1) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "GuidedF_2 DISPARITY_RIGHT throws in the reference%s%s");
    if (d_begin < 0 || d_end > method_n_eval(m) || d_begin > d_end) return asw_fail(ctx, ASW_ERR_BAD_ARG, "bad disparity range%s%s");
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = L->rows, W = L->cols; size_t n = (size_t)H * W;
    uint8_t *dL, *dR;
    ASW_TRY(upload_pair(ctx, L, R, &dL, &dR));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    if (d_end > d_begin) {
        if (m.id == M_GF2) {
            ASW_TRY(dev_guidedf2_keys(ctx, dL, dR, H, W, disp_type, m.p0, win, min_d, num_d, d_begin, d_end, keys, nullptr));
        } else {
            // every candidate of these methods is independent of the range it is evaluated in (weights, distances and
            // grids do not depend on it; BLO(1) keeps the full range's normaliser and crop geometry): the method runs on
            // the sub-range and leaves its keys in WS_KEYS
            float* tmp;
            ASW_TRY(ws_get(ctx, WS_OUT, n, &tmp));
            const int cnt = d_end - d_begin;
            if (m.id == M_TRAD) ASW_TRY(dev_traditional(ctx, dL, dR, H, W, m.p0, m.p1, disp_type, win, min_d + d_begin, cnt - 1, tmp, nullptr));
            else if (m.id == M_D8) ASW_TRY(dev_traditional(ctx, dL, dR, H, W, 30.0, (double)(win * 2 / 3), disp_type, win, min_d + d_begin, cnt - 1, tmp, nullptr, 1));
            else if (m.id == M_GEO) ASW_TRY(dev_geodesic(ctx, dL, dR, H, W, disp_type, win, min_d + d_begin, cnt - 1, tmp, nullptr));
            else if (m.id == M_GRID) ASW_TRY(dev_bilateral_grid(ctx, dL, dR, H, W, m.p0, m.p1, min_d + d_begin, cnt - 1, tmp, nullptr));
            else ASW_TRY(dev_blo1_range(ctx, dL, dR, H, W, disp_type, m.p0, win, min_d, num_d, d_begin, d_end, tmp, nullptr));
        }
    }
    *device_keys = keys;
    return ASW_OK;
}
extern "C" asw_status asw_split_local_keys(asw_ctx* ctx, const asw_u8_image* L, const asw_u8_image* R, int algorithm,
                                           int disp_type, int win, int min_d, int num_d, int d_begin, int d_end,
                                           void** device_keys) {
    ASW_TRY(split_local_keys_dev(ctx, L, R, algorithm, disp_type, win, min_d, num_d, d_begin, d_end, device_keys));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
// in-place x ^= 1 << 63 on a device key buffer: the order-preserving map u64 <-> i64 for collectives that only know signed
// 64-bit MIN (torch.distributed); asynchronous on the ctx stream
__global__ void k_keys_flip_sign(unsigned long long* k, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k[i] ^= 0x8000000000000000ull;
}
extern "C" asw_status asw_keys_flip_sign(asw_ctx* ctx, void* device_keys, int rows, int cols) {
    if (!ctx || !device_keys || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)rows * cols;
    LAUNCH(ctx, "keys_flip_sign", (k_keys_flip_sign<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((unsigned long long*)device_keys, n)));
    return ASW_OK;
}
extern "C" asw_status asw_keys_alloc(asw_ctx* ctx, int rows, int cols, void** device_keys) {
    if (!ctx || !device_keys || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS2, (size_t)rows * cols, &keys));
    ASW_TRY(init_keys(ctx, keys, (size_t)rows * cols));
    *device_keys = keys;
    return ASW_OK;
}
extern "C" asw_status asw_keys_download(asw_ctx* ctx, const void* device_keys, int rows, int cols, uint64_t* host_keys) {
    if (!ctx || !device_keys || !host_keys || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    ASW_CUDA(ctx, cudaMemcpyAsync(host_keys, device_keys, (size_t)rows * cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_keys_upload(asw_ctx* ctx, const uint64_t* host_keys, int rows, int cols, void* device_keys) {
    if (!ctx || !device_keys || !host_keys || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    ASW_CUDA(ctx, cudaMemcpyAsync(device_keys, host_keys, (size_t)rows * cols * 8, cudaMemcpyHostToDevice, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}
extern "C" asw_status asw_keys_min_merge(asw_ctx* ctx, void* inout, const void* other, int rows, int cols) {
    if (!ctx || !inout || !other || rows <= 0 || cols <= 0) return ASW_ERR_BAD_ARG;
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)rows * cols;
    LAUNCH(ctx, "keys_min_merge", (k_keys_min_merge<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
                                      (unsigned long long*)inout, (const unsigned long long*)other, n)));
    return ASW_OK;
}
extern "C" asw_status asw_keys_to_disparity(asw_ctx* ctx, const void* device_keys, asw_f32_image* disp) {
    if (!ctx || !device_keys) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_f32(ctx, disp));
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n = (size_t)disp->rows * disp->cols;
    float* out;
    ASW_TRY(ws_get(ctx, WS_OUT, n, &out));
    ASW_TRY(keys_to_disp(ctx, (const unsigned long long*)device_keys, n, out));
    ASW_TRY(download_f32(ctx, out, disp));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

extern "C" asw_status asw_geodesic_dist(asw_ctx* ctx, const asw_u8_image* img, int win, float* host_dist) {
    if (!ctx) return ASW_ERR_BAD_ARG;
    ASW_TRY(check_u8(ctx, img, 3));
    if (!host_dist || win <= 0 || win % 2 == 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "window must be odd%s%s");   // A.cpp:1394-1397
    ASW_CUDA(ctx, cudaSetDevice(ctx->device));
    int H = img->rows, W = img->cols; size_t n = (size_t)H * W;
    uint8_t* di;
    ASW_TRY(ws_get(ctx, WS_IMG_L, n * 3, &di));
    ASW_TRY(upload_u8(ctx, img, di));
    return dev_geodesic_dist_to_host(ctx, di, H, W, win, host_dist);
}
