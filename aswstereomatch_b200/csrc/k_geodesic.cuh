// k_geodesic.cuh -- placeholder
#pragma once
static asw_status dev_geodesic(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int win, int min_d, int num_d, float* disp_dev, float* agg_dev) { return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "not built yet%s%s"); }
static asw_status dev_geodesic_dist_to_host(asw_ctx* ctx, const uint8_t* img, int H, int W, int win, float* host) { return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "not built yet%s%s"); }
