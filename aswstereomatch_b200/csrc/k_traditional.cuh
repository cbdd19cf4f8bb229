// k_traditional.cuh -- placeholder
#pragma once
static asw_status dev_traditional(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, double gamma_c, double gamma_g, int disp_type, int win, int min_d, int num_d, float* disp_dev, float* agg_dev) { return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "not built yet%s%s"); }
