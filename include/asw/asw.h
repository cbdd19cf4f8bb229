/*
 * asw.h -- C ABI of the B200-native aswStereoMatch dense-matching hot path.
 *
 * Drop-in boundary for ZhangYY12345/aswStereoMatch's method entry points
 * (aswStereoMatch/methods/aswMethods.h, cited below as A.h:<line>, bodies in
 * aswMethods.cpp = A.cpp, enums in parametersStereo.h = P.h).  The reference
 * exposes free C++ functions over cv::Mat; this header exposes the same calls
 * over plain pointers and sizes so that any host (the C++ shim in
 * aswMethods_compat.h, ctypes, cgo, JNI ...) can bind them.  No OpenCV, torch or
 * CUDA types appear in a signature.
 *
 * Image convention = cv::Mat: row-major, `step` bytes per row, colour images are
 * BGR interleaved CV_8UC3 (channels == 3).  All image pointers are HOST pointers
 * owned by the caller; inputs are never modified (the reference takes cv::Mat by
 * value and only re-seats its own headers, A.cpp:2272, A.cpp:1847).  Disparity
 * maps are CV_32FC1 (integral values in [minDisparity, minDisparity+D_eval-1]).
 *
 * One asw_ctx = one CUDA device + one stream + its workspaces.  A ctx is not
 * thread-safe; distinct ctxs are.  Every computation runs in hand-written sm_100a
 * CUDA kernels; there is no CPU fallback: without a usable device asw_create fails
 * with ASW_ERR_CUDA.
 */
#ifndef ASW_ASW_H
#define ASW_ASW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums: integer values identical to P.h:4-24 ---- */
enum { ASW_DISPARITY_LEFT = 0, ASW_DISPARITY_RIGHT = 1 };
enum {
    ASW_ALG_BM = 0, ASW_ALG_SGBM = 1, ASW_ALG_ADAPTIVE_WEIGHT = 2, ASW_ALG_ADAPTIVE_WEIGHT_8DIRECT = 3,
    ASW_ALG_ADAPTIVE_WEIGHT_GEODESIC = 4, ASW_ALG_ADAPTIVE_WEIGHT_BILATERAL_GRID = 5,
    ASW_ALG_ADAPTIVE_WEIGHT_BLO1 = 6, ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER = 7,
    ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER_2 = 8, ASW_ALG_ADAPTIVE_WEIGHT_GUIDED_FILTER_3 = 9,
    ASW_ALG_ADAPTIVE_WEIGHT_MEDIAN = 10, ASW_ALG_NCC = 11
};

typedef enum {
    ASW_OK = 0,
    ASW_ERR_BAD_ARG = 1,        /* null pointer, even window, bad channels: reference returns Mat() */
    ASW_ERR_SIZE_MISMATCH = 2,  /* left/right (or map) sizes differ */
    ASW_ERR_CUDA = 3,           /* CUDA runtime / launch failure, or no device */
    ASW_ERR_UNSUPPORTED = 4,    /* reference behaviour undefined (throws / UB) or out of scope */
    ASW_ERR_NOMEM = 5
} asw_status;

typedef struct { const uint8_t* data; int rows, cols, channels; size_t step; } asw_u8_image;
typedef struct { float* data; int rows, cols; size_t step; } asw_f32_image;
typedef struct { uint8_t* data; int rows, cols; size_t step; } asw_mask_image;

typedef struct asw_ctx asw_ctx;

/* ---- lifetime ---- */
int asw_device_count(void);
asw_status asw_create(int device, asw_ctx** out);
void asw_destroy(asw_ctx* ctx);
const char* asw_last_error(const asw_ctx* ctx);
const char* asw_version(void);
asw_status asw_sync(asw_ctx* ctx);
void* asw_stream(asw_ctx* ctx);                       /* the ctx's cudaStream_t */

/* tuning knobs (0 restores the built-in choice): row bands per strip of the streaming guided filter, disparities per CTA
 * of the BLO(1) kernel, MiB of (a,b) workspace per chunk of the generic guided filter */
enum { ASW_TUNE_GFS_BANDS = 0, ASW_TUNE_BLO_DCH = 1, ASW_TUNE_GF_CHUNK_MB = 2, ASW_TUNE_COUNT = 3 };
asw_status asw_set_tuning(asw_ctx* ctx, int key, int value);

/* pinned host memory for callers that want full PCIe bandwidth */
void* asw_host_alloc(size_t bytes);
void asw_host_free(void* p);

/* ---- dispatcher: stereoMatching (A.h:91-92, A.cpp:46-88) with the dispatcher's literals ---- */
asw_status asw_stereo_matching(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                               asw_f32_image* disparity, int disparity_type, int algorithm_type,
                               int win_size, int min_disparity, int num_disparity);

/* number of candidate disparities the dispatcher's method scans for a named numDisparity: num_disparity + 1 for
 * traditional / geodesic / bilateral grid (their loops run to max_offset inclusive, A.cpp:1074, 1467, 2279),
 * num_disparity for the others; -1 for algorithms outside the hot path.  Needs no device. */
int asw_method_candidates(int algorithm_type, int num_disparity);

/* ---- per-method entry points (same argument order and meaning as A.h) ---- */
/* computeAdaptiveWeight (A.h:133-134, A.cpp:1016-1156) */
asw_status asw_adaptive_weight(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                               asw_f32_image* disparity, double gamma_c, double gamma_g, int disp_type,
                               int win_size, int min_disparity, int num_disparity);
/* computeAdaptiveWeight_direct8 (A.h:135-136, A.cpp:1167-1319): DISPARITY_LEFT only (the reference's RIGHT branch indexes its
 * weight lists out of bounds, A.cpp:1291) */
asw_status asw_adaptive_weight_direct8(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                       asw_f32_image* disparity, int disp_type, int win_size, int min_disparity,
                                       int num_disparity);
/* computeAdaptiveWeight_geodesic (A.h:141-142, A.cpp:1436-1534) */
asw_status asw_adaptive_weight_geodesic(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                        asw_f32_image* disparity, int disp_type, int win_size,
                                        int min_disparity, int num_disparity);
/* computeAdaptiveWeight_bilateralGrid (A.h:153-155, A.cpp:2253-2430) */
asw_status asw_adaptive_weight_bilateral_grid(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                              asw_f32_image* disparity, int disp_type, double sample_rate_s,
                                              double sample_rate_r, int min_disparity, int num_disparity);
/* computeAdaptiveWeight_BLO1 (A.h:158-160, A.cpp:2505-2725) */
asw_status asw_adaptive_weight_blo1(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                    asw_f32_image* disparity, int disp_type, double sample_rate_r,
                                    int win_size, int min_disparity, int num_disparity);
/* computeAdaptiveWeight_GuidedF (A.h:164-166, A.cpp:2867-2963) */
asw_status asw_adaptive_weight_guidedf(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                       asw_f32_image* disparity, int disp_type, double eps, int win_size,
                                       int min_disparity, int num_disparity);
/* computeAdaptiveWeight_GuidedF_2 (A.h:167-169, A.cpp:2976-3050) */
asw_status asw_adaptive_weight_guidedf_2(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                         asw_f32_image* disparity, int disp_type, double eps, int win_size,
                                         int min_disparity, int num_disparity);
/* computeAdaptiveWeight_GuidedF_3 (A.h:170-172, A.cpp:3063-3137): NCC cost, 6-channel guide (LEFT) */
asw_status asw_adaptive_weight_guidedf_3(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                         asw_f32_image* disparity, int disp_type, double eps, int win_size,
                                         int min_disparity, int num_disparity);
/* computeNCC, Mat overload (A.h:124-125, A.cpp:812-912; the dispatcher's NCC).  As in the reference, LEFT scans offsets
 * min .. max - 1 and keeps the minimum raw cost; RIGHT leaves the map at 0 */
asw_status asw_ncc(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right, asw_f32_image* disparity,
                   int disp_type, int win_size, int min_disparity, int num_disparity);
/* computeAdaptiveWeight_WeightedMedian (A.h:176-179, A.cpp:3228-3383) */
asw_status asw_adaptive_weight_weighted_median(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                               asw_f32_image* disparity, int disp_type, int win_size,
                                               double sample_rate_s, double sample_rate_r,
                                               int min_disparity, int num_disparity);

/* When set (non-NULL), the next per-method call also copies its aggregated cost volume,
 * [D_eval][rows][cols] float, to `host_volume` (capacity in floats).  Cleared after one call.
 * Parity-test hook for "aggregated float costs within 1e-4". */
asw_status asw_capture_aggregated(asw_ctx* ctx, float* host_volume, size_t capacity_floats);

/* ---- stage level ---- */
/* computeSimilarity 7-arg (A.h:109-111, A.cpp:415-487): volume [num_disparity][rows][cols] */
asw_status asw_cost_tad_cg(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                           float* host_volume, double regularity, double thres_c, double thres_g,
                           int disp_type, int min_disparity, int num_disparity);
/* computeSimilarity 8-arg (A.h:112-114, A.cpp:488-668): the same cost, every slice padded by win_size / 2 with BORDER_REFLECT
 * (what the weighted-median method consumes): volume [num_disparity][rows + 2h][cols + 2h] */
asw_status asw_cost_tad_cg_padded(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                  float* host_volume, double regularity, double thres_c, double thres_g,
                                  int disp_type, int win_size, int min_disparity, int num_disparity);
/* getCostSAD_d for every d (A.h:157, A.cpp:2442-2503 as called at A.cpp:2524-2536) */
asw_status asw_cost_sad_box(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                            float* host_volume, int disp_type, int win_size, int min_disparity,
                            int num_disparity);
/* computeNCC, vector overload (A.h:126-128, A.cpp:924-1013): [num_disparity][rows][cols], slices min-max normalised */
asw_status asw_cost_ncc(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right, float* host_volume,
                        int disp_type, int win_size, int min_disparity, int num_disparity);
/* the inlined WTA blocks (A.cpp:3032-3048 ...): strict <, ascending d, NaN never wins, untouched = 0 */
asw_status asw_wta(asw_ctx* ctx, const float* host_volume, int num_slices, int rows, int cols,
                   int min_disparity, asw_f32_image* disparity);
/* getGuidedFilter (A.h:163, A.cpp:2766-2854); guide channels 3 or 6 */
asw_status asw_guided_filter(asw_ctx* ctx, const asw_u8_image* guide, const asw_f32_image* input_p,
                             int r, double eps, asw_f32_image* out);
/* geodesic distance windows getGeodesicDist (A.h:140, A.cpp:1392-1424): [rows][cols][win*win] */
asw_status asw_geodesic_dist(asw_ctx* ctx, const asw_u8_image* img, int win_size, float* host_dist);

/* ---- stage 4 (not in the reference; specification in DESIGN.md / SURVEY 8 a-14) ---- */
asw_status asw_lr_check(asw_ctx* ctx, const asw_f32_image* disp_left, const asw_f32_image* disp_right,
                        float tol, asw_mask_image* valid);
asw_status asw_fill_invalid(asw_ctx* ctx, const asw_f32_image* disp, const asw_mask_image* valid,
                            asw_f32_image* out);
asw_status asw_wmedian_refine(asw_ctx* ctx, const asw_u8_image* img, const asw_f32_image* filled,
                              const asw_mask_image* valid, int win_size, double rate_s, double rate_r,
                              asw_f32_image* out);
/* full frame of configs 2/5: GuidedF_2 left + right view, LR check, fill, weighted-median refine.
 * Optional outputs (may be NULL): raw left/right maps and the validity mask. */
asw_status asw_guidedf2_lr_refine(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                  asw_f32_image* refined, double eps, int win_size, int min_disparity,
                                  int num_disparity, float lr_tol, double rate_s, double rate_r,
                                  asw_f32_image* raw_left, asw_f32_image* raw_right, asw_mask_image* valid);

/* driver post-processing (aswStereoMatch.cpp:97-98): disparityMap.convertTo(CV_8UC1) + normalize(0, 255, NORM_MINMAX) */
asw_status asw_disparity_to_u8(asw_ctx* ctx, const asw_f32_image* disparity, asw_mask_image* out_u8);

/* driver pre-processing (aswStereoMatch.cpp:30-31, 67-89): resize(src, Size(dst->cols, dst->rows)) [the driver: 640 x 360] +
 * cvtColor(BGR2HSV) + bilateralFilter(V, 7, 10, 3, BORDER_REFLECT) + V = V + 2 (V - blur) + cvtColor(HSV2BGR), CV_8UC3.
 * OpenCV 4.13's 8-bit arithmetic (resize / BGR2HSV bit-exact; HSV2BGR as cv2's vector body: see oracle/preproc.py) */
asw_status asw_preprocess(asw_ctx* ctx, const asw_u8_image* src, asw_u8_image* dst);

/* ---- device-resident batches (config 5; inputs live in HBM between upload and run) ---- */
typedef struct asw_batch asw_batch;
asw_status asw_batch_create(asw_ctx* ctx, int n_pairs, int rows, int cols, asw_batch** out);
void asw_batch_destroy(asw_batch* b);
/* the run calls below process pairs [0, n_active) (default: all n_pairs) */
asw_status asw_batch_set_active(asw_batch* b, int n_active);
asw_status asw_batch_upload(asw_batch* b, int index, const asw_u8_image* left, const asw_u8_image* right);
/* raw frames of any size: uploaded once, pre-processed on the device (asw_preprocess) into the pair's slot */
asw_status asw_batch_upload_raw(asw_batch* b, int index, const asw_u8_image* left_raw, const asw_u8_image* right_raw);
/* asynchronous on the ctx stream: every pair through asw_guidedf2_lr_refine's pipeline */
asw_status asw_batch_run_guidedf2_lr_refine(asw_batch* b, double eps, int win_size, int min_disparity,
                                            int num_disparity, float lr_tol, double rate_s, double rate_r);
/* asynchronous: every pair through one method of the dispatcher (left view only) */
asw_status asw_batch_run_method(asw_batch* b, int algorithm_type, int disp_type, int win_size,
                                int min_disparity, int num_disparity);
asw_status asw_batch_download(asw_batch* b, int index, asw_f32_image* disparity);

/* ---- disparity-range split of one pair (multi-GPU) ----
 * Each rank evaluates candidates [d_begin, d_end) and gets per-pixel 64-bit keys
 * (48-bit orderable(cost) << 16 | d) in a device buffer; a MIN all-reduce over ranks (NCCL via
 * torch.distributed, or asw_keys_min_merge for peer buffers) then asw_keys_to_disparity
 * reproduces strict-< / lowest-d / NaN-never-wins exactly.
 * Methods: GuidedF_2, traditional, 8-direction, geodesic, bilateral grid, BLO(1).  [d_begin, d_end) indexes the candidates the
 * method scans: num_disparity of them (GuidedF_2, BLO(1)) or num_disparity + 1 (traditional, 8-direction, geodesic, grid).
 * BLO(1) keeps the full range's normaliser (the reference takes it from the last disparity, A.cpp:2588). */
asw_status asw_split_local_keys(asw_ctx* ctx, const asw_u8_image* left, const asw_u8_image* right,
                                int algorithm_type, int disp_type, int win_size, int min_disparity,
                                int num_disparity, int d_begin, int d_end, void** device_keys);
/* a rows x cols key buffer in the ctx workspace, initialised to "empty" (all ones) */
asw_status asw_keys_alloc(asw_ctx* ctx, int rows, int cols, void** device_keys);
asw_status asw_keys_download(asw_ctx* ctx, const void* device_keys, int rows, int cols, uint64_t* host_keys);
asw_status asw_keys_upload(asw_ctx* ctx, const uint64_t* host_keys, int rows, int cols, void* device_keys);
asw_status asw_keys_min_merge(asw_ctx* ctx, void* device_keys_inout, const void* device_keys_other,
                              int rows, int cols);
asw_status asw_keys_to_disparity(asw_ctx* ctx, const void* device_keys, asw_f32_image* disparity);
/* in-place key ^= 1 << 63 on the device (asynchronous on the ctx stream): the order-preserving map u64 <-> i64, for
 * hosts whose collective only offers a signed 64-bit MIN (torch.distributed); apply before and after the all-reduce */
asw_status asw_keys_flip_sign(asw_ctx* ctx, void* device_keys, int rows, int cols);

/* ---- device pool: the multi-GPU face of stereoMatching (A.h:91-92) for C / C++ hosts ----
 * One asw_ctx per device, one host thread per device inside every call.  n_devices = 0 takes every device of the box. */
typedef struct asw_pool asw_pool;
asw_status asw_pool_create(int n_devices, asw_pool** out);
void asw_pool_destroy(asw_pool* pool);
int asw_pool_size(const asw_pool* pool);
const char* asw_pool_last_error(const asw_pool* pool);
asw_ctx* asw_pool_ctx(asw_pool* pool, int device);              /* the pool's ctx of a device (owned by the pool) */
/* pair sharding (config 5): pair i runs on device i % n through the dispatcher (A.cpp:46-88), no collective.  All pairs
 * of one call have the same size.  Arrays of n_pairs images; host pointers, caller-owned. */
asw_status asw_stereo_matching_batch(asw_pool* pool, int n_pairs, const asw_u8_image* left, const asw_u8_image* right,
                                     asw_f32_image* disparity, int disparity_type, int algorithm_type, int win_size,
                                     int min_disparity, int num_disparity);
/* the same sharding for the full frame of configs 2 / 5 (asw_guidedf2_lr_refine per pair) */
asw_status asw_guidedf2_lr_refine_batch(asw_pool* pool, int n_pairs, const asw_u8_image* left, const asw_u8_image* right,
                                        asw_f32_image* refined, double eps, int win_size, int min_disparity,
                                        int num_disparity, float lr_tol, double rate_s, double rate_r);
/* disparity-range split of ONE pair over the pool's devices: device g evaluates its share of the method's candidates
 * (asw_method_candidates), the u64 keys are MIN-all-reduced in device memory with ncclAllReduce(ncclMin, ncclUint64)
 * over NVLink (libnccl.so.2 is bound at run time; ASW_ERR_UNSUPPORTED without it), device 0 returns the map.
 * Bit-identical to the single-device call for every method that supports the split (see asw_split_local_keys). */
asw_status asw_stereo_matching_split(asw_pool* pool, const asw_u8_image* left, const asw_u8_image* right,
                                     asw_f32_image* disparity, int disparity_type, int algorithm_type, int win_size,
                                     int min_disparity, int num_disparity);
float asw_pool_last_allreduce_ms(const asw_pool* pool);        /* device time of the last split's all-reduce (device 0) */

/* ---- measurement hooks (CUDA events on the ctx stream) ---- */
asw_status asw_timer_start(asw_ctx* ctx);
asw_status asw_timer_stop(asw_ctx* ctx, float* elapsed_ms);      /* synchronises the stream */
asw_status asw_profile_enable(asw_ctx* ctx, int on);              /* per-kernel event timing */
asw_status asw_profile_reset(asw_ctx* ctx);
int asw_profile_count(asw_ctx* ctx);
asw_status asw_profile_entry(asw_ctx* ctx, int index, const char** kernel_name, double* total_ms,
                             long long* launches);
long long asw_launch_count(asw_ctx* ctx);                         /* kernels launched since create/reset */
asw_status asw_flush_l2(asw_ctx* ctx);                            /* writes a >L2-sized scratch buffer */

#ifdef __cplusplus
}
#endif
#endif /* ASW_ASW_H */
