// asw_pool.inl -- multi-device entry points of the C ABI (included at the end of asw_lib.cu).
//
//   asw_pool_create                 one asw_ctx per device of the box
//   asw_stereo_matching_batch       pair sharding: pair i -> device i % n, no collective on the data path
//   asw_guidedf2_lr_refine_batch    the config-5 pipeline, pair-sharded the same way
//   asw_stereo_matching_split       disparity-range split of ONE pair: device g evaluates its share of the candidates,
//                                   the per-pixel 64-bit WTA keys are MIN-all-reduced IN PLACE in device memory with
//                                   ncclAllReduce(ncclMin, ncclUint64) over NVLink, device 0 turns the keys into the map
// One host thread per device drives that device's ctx (a ctx is not thread-safe, distinct ctxs are).  NCCL is bound at run
// time (dlopen of libnccl.so.2) so that the library also loads on hosts without it; only the split needs it.
#include <dlfcn.h>
#include <nccl.h>

#include <thread>

struct asw_pool {
    int n = 0;
    std::vector<asw_ctx*> ctx;
    std::vector<asw_batch*> batch;          // cached per device, re-created when the geometry changes
    std::vector<ncclComm_t> comm;
    void* nccl = nullptr;
    ncclResult_t (*pCommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*pAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*pGetErrorString)(ncclResult_t) = nullptr;
    char err[512] = {0};
    float last_allreduce_ms = 0.f;
};

static asw_status pool_fail(asw_pool* p, asw_status st, const char* what, const char* detail = "") {
    if (p) snprintf(p->err, sizeof(p->err), "%s%s%s", what, detail[0] ? ": " : "", detail);
    return st;
}

extern "C" asw_status asw_pool_create(int n_devices, asw_pool** out) {
    if (!out) return ASW_ERR_BAD_ARG;
    *out = nullptr;
    const int have = asw_device_count();
    if (n_devices == 0) n_devices = have;
    if (n_devices < 0 || n_devices > have || have <= 0) return ASW_ERR_CUDA;
    asw_pool* p = new asw_pool();
    p->n = n_devices;
    p->ctx.assign(n_devices, nullptr);
    p->batch.assign(n_devices, nullptr);
    for (int g = 0; g < n_devices; g++) {
        if (asw_create(g, &p->ctx[g]) != ASW_OK) {
            for (int j = 0; j < g; j++) asw_destroy(p->ctx[j]);
            delete p;
            return ASW_ERR_CUDA;
        }
    }
    *out = p;
    return ASW_OK;
}
extern "C" void asw_pool_destroy(asw_pool* p) {
    if (!p) return;
    for (int g = 0; g < p->n; g++) {
        if (g < (int)p->comm.size() && p->comm[g] && p->pCommDestroy) { cudaSetDevice(g); p->pCommDestroy(p->comm[g]); }
        if (p->batch[g]) asw_batch_destroy(p->batch[g]);
        asw_destroy(p->ctx[g]);
    }
    if (p->nccl) dlclose(p->nccl);
    delete p;
}
extern "C" int asw_pool_size(const asw_pool* p) { return p ? p->n : 0; }
extern "C" const char* asw_pool_last_error(const asw_pool* p) { return p ? p->err : "null pool"; }
extern "C" asw_ctx* asw_pool_ctx(asw_pool* p, int device) { return (p && device >= 0 && device < p->n) ? p->ctx[device] : nullptr; }
extern "C" float asw_pool_last_allreduce_ms(const asw_pool* p) { return p ? p->last_allreduce_ms : 0.f; }

// run fn(g) on one host thread per device; the first failing status (by device order) is returned with that ctx's message
template <typename F> static asw_status pool_parallel(asw_pool* p, F fn) {
    std::vector<asw_status> st(p->n, ASW_OK);
    std::vector<std::thread> th;
    for (int g = 1; g < p->n; g++) th.emplace_back([&, g] { st[g] = fn(g); });
    st[0] = fn(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < p->n; g++)
        if (st[g] != ASW_OK) return pool_fail(p, st[g], "device call failed", asw_last_error(p->ctx[g]));
    return ASW_OK;
}

static asw_status pool_check_batch(asw_pool* p, int n_pairs, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp) {
    if (!p || n_pairs <= 0 || !L || !R || !disp) return pool_fail(p, ASW_ERR_BAD_ARG, "null batch arguments");
    for (int i = 0; i < n_pairs; i++) {
        asw_status st = check_pair(p->ctx[0], &L[i], &R[i], &disp[i]);
        if (st != ASW_OK) return pool_fail(p, st, "bad pair", asw_last_error(p->ctx[0]));
        if (L[i].rows != L[0].rows || L[i].cols != L[0].cols) return pool_fail(p, ASW_ERR_SIZE_MISMATCH, "the pairs of a batch must have one size");
    }
    return ASW_OK;
}
// device g owns pairs g, g + n, ...: upload (own stream), run, download (own stream) -- the copies of pair i +- 1 overlap
// the kernels of pair i exactly as in the single-device batch path
template <typename RUN> static asw_status pool_run_batch(asw_pool* p, int n_pairs, const asw_u8_image* L, const asw_u8_image* R,
                                                          asw_f32_image* disp, RUN run) {
    ASW_TRY(pool_check_batch(p, n_pairs, L, R, disp));
    const int H = L[0].rows, W = L[0].cols;
    return pool_parallel(p, [&](int g) -> asw_status {
        const int mine = (n_pairs - g + p->n - 1) / p->n;
        if (mine <= 0) return ASW_OK;
        asw_batch*& b = p->batch[g];
        if (b && (b->n < mine || b->H != H || b->W != W)) { asw_batch_destroy(b); b = nullptr; }
        if (!b) ASW_TRY(asw_batch_create(p->ctx[g], mine, H, W, &b));
        ASW_TRY(asw_batch_set_active(b, mine));
        for (int j = 0; j < mine; j++) ASW_TRY(asw_batch_upload(b, j, &L[g + j * p->n], &R[g + j * p->n]));
        ASW_TRY(run(b));
        for (int j = 0; j < mine; j++) ASW_TRY(asw_batch_download(b, j, &disp[g + j * p->n]));
        return asw_sync(p->ctx[g]);
    });
}
extern "C" asw_status asw_stereo_matching_batch(asw_pool* p, int n_pairs, const asw_u8_image* L, const asw_u8_image* R,
                                                asw_f32_image* disp, int disp_type, int algorithm, int win, int min_d, int num_d) {
    if (!p) return ASW_ERR_BAD_ARG;
    return pool_run_batch(p, n_pairs, L, R, disp, [&](asw_batch* b) { return asw_batch_run_method(b, algorithm, disp_type, win, min_d, num_d); });
}
extern "C" asw_status asw_guidedf2_lr_refine_batch(asw_pool* p, int n_pairs, const asw_u8_image* L, const asw_u8_image* R,
                                                   asw_f32_image* refined, double eps, int win, int min_d, int num_d, float tol,
                                                   double rate_s, double rate_r) {
    if (!p) return ASW_ERR_BAD_ARG;
    return pool_run_batch(p, n_pairs, L, R, refined, [&](asw_batch* b) {
        return asw_batch_run_guidedf2_lr_refine(b, eps, win, min_d, num_d, tol, rate_s, rate_r);
    });
}

static asw_status pool_nccl_init(asw_pool* p) {
    if (!p->comm.empty()) return ASW_OK;
    if (!p->nccl) {
        p->nccl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!p->nccl) return pool_fail(p, ASW_ERR_UNSUPPORTED, "libnccl.so.2 cannot be loaded", dlerror());
        p->pCommInitAll = (decltype(p->pCommInitAll))dlsym(p->nccl, "ncclCommInitAll");
        p->pCommDestroy = (decltype(p->pCommDestroy))dlsym(p->nccl, "ncclCommDestroy");
        p->pAllReduce = (decltype(p->pAllReduce))dlsym(p->nccl, "ncclAllReduce");
        p->pGetErrorString = (decltype(p->pGetErrorString))dlsym(p->nccl, "ncclGetErrorString");
        if (!p->pCommInitAll || !p->pCommDestroy || !p->pAllReduce || !p->pGetErrorString)
            return pool_fail(p, ASW_ERR_UNSUPPORTED, "libnccl.so.2 lacks the expected symbols");
    }
    std::vector<int> devs(p->n);
    for (int g = 0; g < p->n; g++) devs[g] = g;
    p->comm.assign(p->n, nullptr);
    ncclResult_t r = p->pCommInitAll(p->comm.data(), p->n, devs.data());
    if (r != ncclSuccess) { p->comm.clear(); return pool_fail(p, ASW_ERR_CUDA, "ncclCommInitAll", p->pGetErrorString(r)); }
    return ASW_OK;
}

// candidates [lo, hi) of rank r of n: ranges tile [0, total) exactly, early ranks take the remainder
static void split_range_c(int total, int r, int n, int* lo, int* hi) {
    const int base = total / n, rem = total % n;
    *lo = r * base + std::min(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

extern "C" asw_status asw_stereo_matching_split(asw_pool* p, const asw_u8_image* L, const asw_u8_image* R, asw_f32_image* disp,
                                                int disp_type, int algorithm, int win, int min_d, int num_d) {
    if (!p) return ASW_ERR_BAD_ARG;
    {
        asw_status st = check_pair(p->ctx[0], L, R, disp);
        if (st != ASW_OK) return pool_fail(p, st, "bad pair", asw_last_error(p->ctx[0]));
    }
    const int total = asw_method_candidates(algorithm, num_d);
    if (total < 0) return pool_fail(p, ASW_ERR_UNSUPPORTED, "algorithm outside the dense-matching hot path");
    if (p->n > 1) ASW_TRY(pool_nccl_init(p));
    const size_t n = (size_t)L->rows * L->cols;
    std::vector<void*> keys(p->n, nullptr);
    // 1. every device: its candidate range -> keys in device memory (asynchronous on the ctx stream)
    ASW_TRY(pool_parallel(p, [&](int g) -> asw_status {
        int lo, hi;
        split_range_c(total, g, p->n, &lo, &hi);
        return split_local_keys_dev(p->ctx[g], L, R, algorithm, disp_type, win, min_d, num_d, lo, hi, &keys[g]);
    }));
    // 2. MIN all-reduce of the u64 keys, in place, on each device's compute stream (stream-ordered after the kernels)
    if (p->n > 1) {
        asw_ctx* c0 = p->ctx[0];
        ASW_TRY(pool_parallel(p, [&](int g) -> asw_status {
            asw_ctx* c = p->ctx[g];
            ASW_CUDA(c, cudaSetDevice(c->device));
            if (g == 0) ASW_CUDA(c, cudaEventRecord(c->ev_t0, c->stream));
            ncclResult_t r = p->pAllReduce(keys[g], keys[g], n, ncclUint64, ncclMin, p->comm[g], c->stream);
            if (r != ncclSuccess) return asw_fail(c, ASW_ERR_CUDA, "ncclAllReduce: %s", p->pGetErrorString(r));
            if (g == 0) ASW_CUDA(c, cudaEventRecord(c->ev_t1, c->stream));
            ASW_CUDA(c, cudaStreamSynchronize(c->stream));
            return ASW_OK;
        }));
        cudaSetDevice(c0->device);
        cudaEventElapsedTime(&p->last_allreduce_ms, c0->ev_t0, c0->ev_t1);
    } else p->last_allreduce_ms = 0.f;
    // 3. device 0: keys -> disparity map -> host
    asw_status st = asw_keys_to_disparity(p->ctx[0], keys[0], disp);
    if (st != ASW_OK) return pool_fail(p, st, "keys_to_disparity", asw_last_error(p->ctx[0]));
    return ASW_OK;
}
