// k_wmedian.cuh -- placeholder
#pragma once
static asw_status dev_weighted_median(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int win, double rate_s, double rate_r, int min_d, int num_d, float* disp_dev, float* agg_dev) { return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "not built yet%s%s"); }
