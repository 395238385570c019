// api.cu -- the C-ABI of libb200icp.so (include/b200icp.h): context, setMap, register, match,
// knn, transform.  Host-side orchestration only; all arithmetic is in index.cu / knn.cu / icp.cu.
// There is no CPU fallback anywhere in this library.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

using namespace b200;

struct b200icp_ctx {
    b200icp_config cfg{};
    IcpParams prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    GridIndex map;   // the LIVE index: what icp(input) / match search (icp.setMap's last argument)
    // Online mapping (Mapper isOnline, Mapper.cpp:280-283): between b200icp_map_begin_update and b200icp_map_end_update the update
    // steps build and refine a second index (`work`) from the store while registrations keep using `map`; end_update swaps them --
    // the reference's icp.setMap(localPointCloud) at the end of Map::updateLocalPointCloud (Map.cpp:527-529).
    GridIndex work;
    bool updating = false, work_valid = false;
    mutable std::recursive_mutex api_mutex;  // entry points are atomic with respect to each other (caller thread, update thread, window thread)
    bool has_map = false;
    int64_t map_n = 0;
    MapStore store;  // the device-resident map the index is built from
    bool index_stale = false;
    uint8_t* d_keep = nullptr;
    int64_t cap_keep = 0;
    GridIndex aux;  // index for b200icp_knn on arbitrary clouds
    IcpBuffers buf;
    VarTrimScratch var_scratch;
    float* d_stage_a = nullptr;  // uploads: features
    float* d_stage_b = nullptr;  // uploads: normals / queries
    size_t stage_a_bytes = 0, stage_b_bytes = 0;
    int32_t* d_out_ids = nullptr;
    float* d_out_d2 = nullptr;
    float4* d_q4 = nullptr;
    int64_t cap_out = 0, cap_q4 = 0;
    int* d_scalar_nq = nullptr;
    unsigned* d_bar_counter = nullptr;
    int n_sms = 148;
    int sm_share = 0;  // b200icp_set_sm_share: CTAs (= SMs) the persistent loop kernel may occupy; 0 = all of them
    // the device-resident scan slot (b200icp_scan_*): a whole DataPoints -- features + descriptors -- uploaded once; the steps of
    // Mapper::processInput work on it in place.  Two buffer sets: ordered compaction (input filters) goes from one to the other.
    struct ScanSlot {
        float *feat[2] = {nullptr, nullptr}, *nrm[2] = {nullptr, nullptr}, *prob[2] = {nullptr, nullptr}, *extra[2] = {nullptr, nullptr};
        int cur = 0;
        int64_t cap = 0, cap_extra = 0, n = 0;
        bool has_nrm = false, has_prob = false;
        int extra_rows = 0;
        int n_rot = 0, rot_row[4] = {0, 0, 0, 0};  // descriptors in `extra` that rotate with the cloud (observationDirections)
    } scan, scan_upd;  // scan_upd: the snapshot an asynchronous map update works on (b200icp_scan_snapshot) while the next scan arrives
    // self k-NN with staged tiles (selfknn.cu): the queries it could not prove exact, redone by the shell walk
    uint32_t* d_fb_list = nullptr;
    unsigned* d_fb_count = nullptr;
    float4* d_fb_q = nullptr;
    int32_t* d_fb_ids = nullptr;
    float* d_fb_d2 = nullptr;
    int64_t cap_fb = 0, cap_fb_rows = 0;
    int64_t last_selfknn_redone = -1;  // (introspection: -1 = the staged kernel was not used)
    float* d_kth = nullptr;      // incremental SurfaceNormal: squared k-th neighbour distance per store point
    int64_t cap_kth = 0;
    uint8_t* d_dirty = nullptr;  // ... dirty flags (store order, then index order)
    uint32_t* d_list = nullptr;  // ... index positions to recompute
    int64_t cap_dirty = 0;
    int64_t last_normals_recomputed = 0;
    uint64_t octree_calls = 0;  // advances the random sampler's seed from one b200icp_map_octree call to the next
    float margin3[3] = {3.0f, 0.002f, 0.25f};  // search margin of the loop kernel's match cache; B200ICP_MARGIN="gain,min[m],max[cells]"
    float win3[3] = {2.0f, 0.0015f, 0.01f};  // quantile-window policy of the one-barrier iteration (loop.cu); B200ICP_WINDOW="gain,floor,max"
    char* h_pinned = nullptr;  // [0, 1024): state image to upload, [1024, 2048): state read back, [2048..): ints
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_map0 = nullptr, ev_map1 = nullptr, ev_loop0 = nullptr, ev_loop1 = nullptr;
    std::vector<cudaEvent_t> nn_events;
    bool profiling = false;
    bool want_trace = false;
    std::vector<float> h_trace;  // T_iter after each iteration of the last registration (4x4 each)
    b200icp_timing timing{};
    std::string err;
};

namespace {

std::string g_create_error;
std::mutex g_create_mutex;

// the context whose snapshot slot (scan_upd) this thread's b200icp_scan_* calls work on: set by b200icp_map_begin_update, cleared by
// b200icp_map_end_update -- "the thread that runs the asynchronous update sees the scan it was handed"
thread_local const b200icp_ctx* t_update_ctx = nullptr;
b200icp_ctx::ScanSlot& scan_slot(b200icp_ctx* ctx) { return t_update_ctx == ctx ? ctx->scan_upd : ctx->scan; }
const b200icp_ctx::ScanSlot& scan_slot(const b200icp_ctx* ctx) { return t_update_ctx == ctx ? ctx->scan_upd : ctx->scan; }
// where commit builds (and whose sort scratch the update-side steps borrow) / the index that is in step with the store
GridIndex& build_index(b200icp_ctx* ctx) { return ctx->updating ? ctx->work : ctx->map; }
GridIndex& synced_index(b200icp_ctx* ctx) { return (ctx->updating && ctx->work_valid) ? ctx->work : ctx->map; }
#define B200_LOCK(ctx) std::lock_guard<std::recursive_mutex> api_guard_((ctx)->api_mutex)

int32_t fail(b200icp_ctx* ctx, int32_t code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
int32_t fail_cuda(b200icp_ctx* ctx, cudaError_t e, const char* what) {
    cudaGetLastError();  // clear sticky-less error state
    return fail(ctx, B200ICP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CK(expr)                                                      \
    do {                                                              \
        cudaError_t e_ = (expr);                                      \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #expr);      \
    } while (0)

template <typename T>
cudaError_t grow(T*& p, size_t& have_bytes, size_t need_bytes) {
    if (need_bytes <= have_bytes && p) return cudaSuccess;
    if (p) B200_CUDA_FREE(p);
    p = nullptr;
    have_bytes = 0;
    const size_t cap = need_bytes + need_bytes / 4 + 4096;
    cudaError_t e = B200_CUDA_MALLOC((void**)&p, cap);
    if (e == cudaSuccess) have_bytes = cap;
    return e;
}

void mat4_identity(float* M) {
    for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.f : 0.f;
}
void mat4_mul(const float* A, const float* B, float* C) {
    float tmp[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc += A[k * 4 + r] * B[c * 4 + k];
            tmp[c * 4 + r] = acc;
        }
    memcpy(C, tmp, sizeof(tmp));
}
// (dim+1)^2 column-major <-> embedded 4x4 column-major
void embed(const float* T, int dim, float* M) {
    mat4_identity(M);
    if (!T) return;
    const int n = dim + 1;
    for (int r = 0; r < dim; ++r) {
        for (int c = 0; c < dim; ++c) M[c * 4 + r] = T[c * n + r];
        M[12 + r] = T[dim * n + r];
    }
}
void extract(const float* M, int dim, float* T) {
    const int n = dim + 1;
    for (int i = 0; i < n * n; ++i) T[i] = 0.f;
    for (int r = 0; r < dim; ++r) {
        for (int c = 0; c < dim; ++c) T[c * n + r] = M[c * 4 + r];
        T[dim * n + r] = M[12 + r];
    }
    T[n * n - 1] = 1.f;
}
float det3(const float* M) {
    return M[0] * (M[5] * M[10] - M[9] * M[6]) - M[4] * (M[1] * M[10] - M[9] * M[2]) + M[8] * (M[1] * M[6] - M[5] * M[2]);
}
void quat_from_M(const float* T, float* q) {
    float R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = T[c * 4 + r];
    float t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0.f) {
        t = std::sqrt(t + 1.f);
        q[0] = 0.5f * t;
        t = 0.5f / t;
        q[1] = (R[2][1] - R[1][2]) * t;
        q[2] = (R[0][2] - R[2][0]) * t;
        q[3] = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.f);
        q[1 + i] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (R[k][j] - R[j][k]) * t;
        q[1 + j] = (R[j][i] + R[i][j]) * t;
        q[1 + k] = (R[k][i] + R[i][k]) * t;
    }
}

int32_t validate_config(const b200icp_config* c, std::string& why) {
    if (!c) return why = "null config", B200ICP_ERR_INVALID_ARG;
    if (c->dim != 2 && c->dim != 3) return why = "dim must be 2 or 3", B200ICP_ERR_INVALID_ARG;
    if (c->knn < 1 || c->knn > 32) return why = "knn must be in [1, 32]", B200ICP_ERR_INVALID_ARG;
    if (!(c->max_dist > 0.f)) return why = "maxDist must be positive", B200ICP_ERR_INVALID_ARG;
    if (c->n_outlier < 0 || c->n_outlier > B200ICP_MAX_OUTLIER_FILTERS) return why = "too many outlier filters", B200ICP_ERR_INVALID_ARG;
    int quant = 0, n_robust = 0;
    for (int f = 0; f < c->n_outlier; ++f) {
        const int kd = c->outlier_kind[f];
        if (kd == B200ICP_OUTLIER_ROBUST && ++n_robust > 1) return why = "at most one RobustOutlierFilter per chain", B200ICP_ERR_NOT_IMPLEMENTED;
        if (kd < B200ICP_OUTLIER_TRIMMED_DIST || kd > B200ICP_OUTLIER_ROBUST) return why = "unknown outlier filter", B200ICP_ERR_INVALID_ARG;
        if (kd == B200ICP_OUTLIER_ROBUST) {
            const int mode = c->outlier_mode[f];
            if ((mode & 255) > B200ICP_ROBUST_STUDENT) return why = "RobustOutlierFilter: invalid robust function name", B200ICP_ERR_INVALID_ARG;
            if (((mode >> 8) & 15) > B200ICP_SCALE_STD) return why = "RobustOutlierFilter: invalid scale estimator name", B200ICP_ERR_INVALID_ARG;
            if (mode < 0) return why = "RobustOutlierFilter: invalid mode bits", B200ICP_ERR_INVALID_ARG;
        }
        if (kd == B200ICP_OUTLIER_TRIMMED_DIST || kd == B200ICP_OUTLIER_MEDIAN_DIST || kd == B200ICP_OUTLIER_VAR_TRIMMED_DIST) ++quant;
        if (kd == B200ICP_OUTLIER_VAR_TRIMMED_DIST &&
            !(c->outlier_param[f] >= 0.f && c->outlier_param[f] < c->outlier_param2[f] && c->outlier_param2[f] <= 1.f && c->outlier_param3[f] > 0.f))
            return why = "VarTrimmedDist: need 0 <= minRatio < maxRatio <= 1 and lambda > 0", B200ICP_ERR_INVALID_ARG;
        if (kd == B200ICP_OUTLIER_TRIMMED_DIST && !(c->outlier_param[f] >= 0.f && c->outlier_param[f] <= 1.f))
            return why = "quantile must be between 0 and 1", B200ICP_ERR_INVALID_ARG;
    }
    if (quant > 1) return why = "at most one quantile-based outlier filter (Trimmed, VarTrimmed or Median) per chain", B200ICP_ERR_NOT_IMPLEMENTED;
    if (c->minimizer < B200ICP_MIN_POINT_TO_PLANE || c->minimizer > B200ICP_MIN_IDENTITY) return why = "unknown error minimizer", B200ICP_ERR_INVALID_ARG;
    if (c->use_differential && (c->smooth_length < 1 || c->smooth_length > 7)) return why = "smoothLength must be in [1, 7]", B200ICP_ERR_INVALID_ARG;
    if ((c->minimizer_flags & 3) == 3) return why = "Force 2D cannot be used together with force4DOF.", B200ICP_ERR_INVALID_ARG;  // LPM ConfigurationError
    if (c->minimizer_flags & ~3) return why = "unknown minimizer flag", B200ICP_ERR_INVALID_ARG;
    if (c->conventions & ~3) return why = "unknown conventions bit", B200ICP_ERR_INVALID_ARG;
    return B200ICP_OK;
}

int32_t ensure_icp_buffers(b200icp_ctx* ctx, int64_t nq) {
    IcpBuffers& b = ctx->buf;
    const int rows = ctx->cfg.dim + 1;
    const int K = ctx->cfg.knn;
    if (!b.state) {
        CK(B200_CUDA_MALLOC((void**)&b.state, kStateBytes));
        CK(B200_CUDA_MALLOC((void**)&b.hist, kHistWords * sizeof(uint32_t)));
        CK(cudaMemsetAsync(b.hist, 0, kHistWords * sizeof(uint32_t), ctx->stream));
        CK(B200_CUDA_MALLOC((void**)&b.partials, (size_t)kMaxAccBlocks * kAccSlots * sizeof(double)));
        CK(B200_CUDA_MALLOC((void**)&b.fastws, icp_loop_workspace_bytes()));
        CK(cudaMemsetAsync(b.fastws, 0, icp_loop_workspace_bytes(), ctx->stream));
    }
    if (nq > b.cap_nq) {
        const int64_t cap = grow_capacity(nq);
        B200_CUDA_FREE(b.reading_in);
        B200_CUDA_FREE(b.reading);
        B200_CUDA_FREE(b.reading_tmp);
        B200_CUDA_FREE(b.match_pos);
        B200_CUDA_FREE(b.match_d2);
        B200_CUDA_FREE(b.spill_pp);
        B200_CUDA_FREE(b.spill_nv);
        b.spill_pp = b.spill_nv = nullptr;
        b.reading_in = nullptr;
        b.reading = b.reading_tmp = nullptr;
        b.match_pos = nullptr;
        b.match_d2 = nullptr;
        b.cap_nq = 0;
        CK(B200_CUDA_MALLOC((void**)&b.reading_in, (size_t)cap * rows * sizeof(float)));
        CK(B200_CUDA_MALLOC((void**)&b.reading, (size_t)cap * sizeof(float4)));
        CK(B200_CUDA_MALLOC((void**)&b.reading_tmp, (size_t)cap * sizeof(float4)));
        CK(B200_CUDA_MALLOC((void**)&b.match_pos, (size_t)cap * K * sizeof(int32_t)));
        CK(B200_CUDA_MALLOC((void**)&b.match_d2, (size_t)cap * K * sizeof(float)));
        CK(B200_CUDA_MALLOC((void**)&b.spill_pp, (size_t)cap * K * sizeof(float4)));  // one row per (reading point, neighbour) pair
        CK(B200_CUDA_MALLOC((void**)&b.spill_nv, (size_t)cap * K * sizeof(float4)));
        b.cap_nq = cap;
    }
    return B200ICP_OK;
}

int32_t ensure_query_buffers(b200icp_ctx* ctx, int64_t nq, int k) {
    if (nq * k > ctx->cap_out) {
        const int64_t cap = grow_capacity(nq * k);
        B200_CUDA_FREE(ctx->d_out_ids);
        B200_CUDA_FREE(ctx->d_out_d2);
        ctx->d_out_ids = nullptr;
        ctx->d_out_d2 = nullptr;
        ctx->cap_out = 0;
        CK(B200_CUDA_MALLOC((void**)&ctx->d_out_ids, (size_t)cap * sizeof(int32_t)));
        CK(B200_CUDA_MALLOC((void**)&ctx->d_out_d2, (size_t)cap * sizeof(float)));
        ctx->cap_out = cap;
    }
    if (nq > ctx->cap_q4) {
        const int64_t cap = grow_capacity(nq);
        B200_CUDA_FREE(ctx->d_q4);
        ctx->d_q4 = nullptr;
        ctx->cap_q4 = 0;
        CK(B200_CUDA_MALLOC((void**)&ctx->d_q4, (size_t)cap * sizeof(float4)));
        ctx->cap_q4 = cap;
    }
    return B200ICP_OK;
}

int bits_for(uint64_t v) {
    int b = 1;
    while (b < 32 && (1ull << b) < v) ++b;
    return b;
}

// icp.setMap(localPointCloud): rebuild the spatial index from the loaded points of the store.
int32_t commit_index(b200icp_ctx* ctx) {
    MapStore& st = ctx->store;
    cudaStream_t s = ctx->stream;
    GridIndex& idx = build_index(ctx);
    if (!st.all_loaded || st.n_active != st.n) CK(store_compact_active(st, idx, s));
    ctx->index_stale = false;
    if (st.n_active == 0) return B200ICP_OK;  // LPM: "Ignoring attempt to create a map from an empty cloud"
    float cell_hint = 0.f;
    if (const char* env = getenv("B200ICP_CELL_EDGE")) cell_hint = (float)atof(env);
    CK(cudaEventRecord(ctx->ev_map0, s));
    // (normals of the points that already had some are carried along even when the cloud formally lost the descriptor to a
    //  concatenate: the incremental SurfaceNormal pass reuses them; has_normals below keeps the formal truth)
    const bool carry_normals = st.has_normals || (st.nrm_epoch_ok && st.nrm != nullptr);
    CK(grid_build(idx, reinterpret_cast<const float*>(st.feat), 4, ctx->cfg.dim, carry_normals ? st.nrm : nullptr, st.n_active,
                  /*centre=*/true, cell_hint, s, st.all_loaded ? nullptr : st.active));
    idx.has_normals = st.has_normals;
    CK(cudaEventRecord(ctx->ev_map1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_map0, ctx->ev_map1);
    ctx->timing.setmap_ms = ms;
    if (ctx->updating) {
        ctx->work_valid = true;  // (published by b200icp_map_end_update)
    } else {
        ctx->has_map = true;
        ctx->map_n = st.n_active;
    }
    return B200ICP_OK;
}

}  // namespace

extern "C" {

int32_t b200icp_abi_version(void) { return B200ICP_ABI_VERSION; }

void b200icp_config_default(b200icp_config* cfg, int32_t dim) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->dim = dim;
    cfg->knn = 1;
    cfg->max_dist = INFINITY;
    cfg->epsilon = 0.f;
    cfg->n_outlier = 1;
    cfg->outlier_kind[0] = B200ICP_OUTLIER_TRIMMED_DIST;
    cfg->outlier_param[0] = 0.85f;
    cfg->minimizer = B200ICP_MIN_POINT_TO_PLANE;
    cfg->max_iteration_count = 40;
    cfg->use_differential = 1;
    cfg->min_diff_rot_err = 1e-3f;
    cfg->min_diff_trans_err = 1e-3f;
    cfg->smooth_length = 3;
    cfg->use_bound = 0;
    cfg->max_rotation_norm = 1.f;
    cfg->max_translation_norm = 1.f;
    cfg->sort_reading = 1;
    cfg->use_graph = 1;
    cfg->nn_variant = 0;
}

const char* b200icp_last_error(const b200icp_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    return g_create_error.c_str();
}

int32_t b200icp_create(const b200icp_config* cfg, int32_t device, b200icp_ctx** out) {
    std::lock_guard<std::mutex> lock(g_create_mutex);
    if (!out) return g_create_error = "null out pointer", B200ICP_ERR_INVALID_ARG;
    *out = nullptr;
    std::string why;
    const int32_t vc = validate_config(cfg, why);
    if (vc != B200ICP_OK) return g_create_error = why, vc;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
        return B200ICP_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) return g_create_error = "bad device ordinal", B200ICP_ERR_INVALID_ARG;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return g_create_error = cudaGetErrorString(e), B200ICP_ERR_CUDA;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return g_create_error = cudaGetErrorString(e), B200ICP_ERR_CUDA;
    if (prop.major != 10) {
        g_create_error = "libb200icp.so is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return B200ICP_ERR_CUDA;
    }
    b200icp_ctx* ctx = new b200icp_ctx();
    ctx->cfg = *cfg;
    ctx->device = device;
    IcpParams& p = ctx->prm;
    p.dim = cfg->dim;
    p.knn = cfg->knn;
    p.max_r2 = std::isinf(cfg->max_dist) ? INFINITY : cfg->max_dist * cfg->max_dist;
    p.n_outlier = cfg->n_outlier;
    p.quantile_filter = -1;
    p.quantile = 0.f;
    for (int f = 0; f < cfg->n_outlier; ++f) {
        p.outlier_kind[f] = cfg->outlier_kind[f];
        p.outlier_param[f] = cfg->outlier_param[f];
        p.outlier_param2[f] = cfg->outlier_param2[f];
        p.outlier_param3[f] = cfg->outlier_param3[f];
        p.outlier_mode[f] = cfg->outlier_mode[f];
        if (cfg->outlier_kind[f] == B200ICP_OUTLIER_VAR_TRIMMED_DIST) {
            p.quantile_filter = f;
            p.quantile = -1.f;  // tuned per iteration on the device (outlier.cu)
        }
        if (cfg->outlier_kind[f] == B200ICP_OUTLIER_TRIMMED_DIST) {
            p.quantile_filter = f;
            p.quantile = cfg->outlier_param[f];
        } else if (cfg->outlier_kind[f] == B200ICP_OUTLIER_MEDIAN_DIST) {
            p.quantile_filter = f;
            p.quantile = 0.5f;
        }
    }
    p.minimizer = cfg->minimizer;
    p.max_iteration_count = cfg->max_iteration_count;
    p.use_differential = cfg->use_differential;
    p.min_diff_rot_err = cfg->min_diff_rot_err;
    p.min_diff_trans_err = cfg->min_diff_trans_err;
    p.smooth_length = cfg->smooth_length;
    p.use_bound = cfg->use_bound;
    p.max_rotation_norm = cfg->max_rotation_norm;
    p.max_translation_norm = cfg->max_translation_norm;
    p.counter_after = cfg->checker_order & 3;
    p.min_flags = (cfg->dim == 3 && cfg->minimizer == B200ICP_MIN_POINT_TO_PLANE) ? (cfg->minimizer_flags & 3) : 0;
    // conventions (SURVEY App. A "(?)" items): strict '<' at maxDist == '<=' at the next float below; Median factor on the distance
    if ((cfg->conventions & 1) && !std::isinf(cfg->max_dist)) p.max_r2 = std::nextafterf(p.max_r2, 0.f);
    if (cfg->conventions & 2)
        for (int f = 0; f < cfg->n_outlier; ++f)
            if (cfg->outlier_kind[f] == B200ICP_OUTLIER_MEDIAN_DIST) p.outlier_param[f] = cfg->outlier_param[f] * cfg->outlier_param[f];
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev_begin) == cudaSuccess && cudaEventCreate(&ctx->ev_end) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev_map0) == cudaSuccess && cudaEventCreate(&ctx->ev_map1) == cudaSuccess;
    ok = ok && cudaEventCreate(&ctx->ev_loop0) == cudaSuccess && cudaEventCreate(&ctx->ev_loop1) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&ctx->h_pinned, 4096) == cudaSuccess;
    ok = ok && B200_CUDA_MALLOC((void**)&ctx->d_scalar_nq, 64) == cudaSuccess;
    ok = ok && icp_device_setup() == cudaSuccess;
    ok = ok && B200_CUDA_MALLOC((void**)&ctx->d_bar_counter, 64) == cudaSuccess;
    ctx->n_sms = prop.multiProcessorCount;
    if (const char* env = getenv("B200ICP_MARGIN")) {
        float a, f, m;
        if (sscanf(env, "%f,%f,%f", &a, &f, &m) == 3 && a >= 0.f && f >= 0.f && m >= 0.f) {
            ctx->margin3[0] = a;
            ctx->margin3[1] = f;
            ctx->margin3[2] = m;
        }
    }
    if (const char* env = getenv("B200ICP_WINDOW")) {
        float a, f, m;
        if (sscanf(env, "%f,%f,%f", &a, &f, &m) == 3 && a >= 0.f && f > 0.f && m >= f) {
            ctx->win3[0] = a;
            ctx->win3[1] = f;
            ctx->win3[2] = m;
        }
    }
    if (!ok) {
        g_create_error = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        b200icp_destroy(ctx);
        return B200ICP_ERR_CUDA;
    }
    *out = ctx;
    return B200ICP_OK;
}

void b200icp_destroy(b200icp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    grid_free(ctx->map);
    grid_free(ctx->work);
    grid_free(ctx->aux);
    store_free(ctx->store);
    B200_CUDA_FREE(ctx->d_keep);
    IcpBuffers& b = ctx->buf;
    B200_CUDA_FREE(b.reading_in);
    B200_CUDA_FREE(b.reading);
    B200_CUDA_FREE(b.reading_tmp);
    B200_CUDA_FREE(b.match_pos);
    B200_CUDA_FREE(b.match_d2);
    B200_CUDA_FREE(b.rnrm_in);
    B200_CUDA_FREE(b.rnrm);
    B200_CUDA_FREE(b.rnrm_tmp);
    B200_CUDA_FREE(b.rmax_in);
    B200_CUDA_FREE(b.hist);
    B200_CUDA_FREE(b.partials);
    B200_CUDA_FREE(b.state);
    B200_CUDA_FREE(b.trace);
    var_trimmed_free(ctx->var_scratch);
    for (int i = 0; i < 2; ++i) {
        for (b200icp_ctx::ScanSlot* sl : {&ctx->scan, &ctx->scan_upd}) {
            B200_CUDA_FREE(sl->feat[i]);
            B200_CUDA_FREE(sl->nrm[i]);
            B200_CUDA_FREE(sl->prob[i]);
            B200_CUDA_FREE(sl->extra[i]);
        }
    }
    B200_CUDA_FREE(ctx->d_fb_list);
    B200_CUDA_FREE(ctx->d_fb_count);
    B200_CUDA_FREE(ctx->d_fb_q);
    B200_CUDA_FREE(ctx->d_fb_ids);
    B200_CUDA_FREE(ctx->d_fb_d2);
    B200_CUDA_FREE(ctx->d_kth);
    B200_CUDA_FREE(ctx->d_dirty);
    B200_CUDA_FREE(ctx->d_list);
    B200_CUDA_FREE(b.fastws);
    B200_CUDA_FREE(b.spill_pp);
    B200_CUDA_FREE(b.spill_nv);
    B200_CUDA_FREE(ctx->d_stage_a);
    B200_CUDA_FREE(ctx->d_stage_b);
    B200_CUDA_FREE(ctx->d_out_ids);
    B200_CUDA_FREE(ctx->d_out_d2);
    B200_CUDA_FREE(ctx->d_q4);
    B200_CUDA_FREE(ctx->d_scalar_nq);
    B200_CUDA_FREE(ctx->d_bar_counter);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (cudaEvent_t ev : ctx->nn_events) cudaEventDestroy(ev);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_map0) cudaEventDestroy(ctx->ev_map0);
    if (ctx->ev_map1) cudaEventDestroy(ctx->ev_map1);
    if (ctx->ev_loop0) cudaEventDestroy(ctx->ev_loop0);
    if (ctx->ev_loop1) cudaEventDestroy(ctx->ev_loop1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* b200icp_stream(b200icp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t b200icp_set_profiling(b200icp_ctx* ctx, int32_t on) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    ctx->profiling = on != 0;
    return B200ICP_OK;
}

int32_t b200icp_set_sm_share(b200icp_ctx* ctx, int32_t n_sms) {
    if (!ctx || n_sms < 0) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    ctx->sm_share = n_sms;
    return B200ICP_OK;
}

int32_t b200icp_get_timing(const b200icp_ctx* ctx, b200icp_timing* out) {
    if (!ctx || !out) return B200ICP_ERR_INVALID_ARG;
    *out = ctx->timing;
    return B200ICP_OK;
}

int32_t b200icp_set_trace(b200icp_ctx* ctx, int32_t on) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    ctx->want_trace = on != 0;
    return B200ICP_OK;
}

int32_t b200icp_get_trace(const b200icp_ctx* ctx, float* out, int32_t max_iterations) {
    if (!ctx || (max_iterations > 0 && !out)) return -1;
    const int dim = ctx->cfg.dim, n = dim + 1;
    const int have = (int)(ctx->h_trace.size() / 16);
    const int cnt = std::min(have, std::max(max_iterations, 0));
    for (int i = 0; i < cnt; ++i) extract(ctx->h_trace.data() + (size_t)i * 16, dim, out + (size_t)i * n * n);
    return have;
}

/* development aid (not in the public header): globaltimer stamps of the last iteration, see B200_STAMP */
int32_t b200icp_debug_stamps(const b200icp_ctx* ctx, unsigned long long* out32) {
    if (!ctx || !out32) return B200ICP_ERR_INVALID_ARG;
    memcpy(out32, ctx->h_pinned + kStateBytes + kDebugOffset, 32 * sizeof(unsigned long long));
    return B200ICP_OK;
}

/* development aid (not in the public header): the loop kernel's per-iteration record of the last registration,
 * 8 words per iteration: {path 0 general / 1 one-barrier / 2 failed attempt, limit bits, candidates, pairs below the
 * window, queries searched so far (all CTAs), next window lo, hi, iteration time in ns on CTA 0} */
int32_t b200icp_debug_loop_record(b200icp_ctx* ctx, uint32_t* out, int32_t iterations) {
    if (!ctx || !out || iterations < 0 || iterations > 256 || !ctx->buf.hist) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->buf.hist + 12288, (size_t)iterations * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

/* development aid (not in the public header; meaningful in the stamped build only): 32 x uint64 %globaltimer stamps per CTA of
 * the loop kernel's last iteration */
int32_t b200icp_debug_cta_stamps(b200icp_ctx* ctx, unsigned long long* out, int32_t n_ctas) {
    if (!ctx || !out || n_ctas < 0 || n_ctas > kMaxAccBlocks || !ctx->buf.partials) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->buf.partials, (size_t)n_ctas * kAccSlots * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

int64_t b200icp_map_size(const b200icp_ctx* ctx) { return (ctx && ctx->has_map) ? ctx->map_n : 0; }

int32_t b200icp_get_map_mean(const b200icp_ctx* ctx, float* mean3) {
    if (!ctx || !mean3) return B200ICP_ERR_INVALID_ARG;
    for (int d = 0; d < 3; ++d) mean3[d] = ctx->map.mean[d];
    return B200ICP_OK;
}

int32_t b200icp_get_grid_info(const b200icp_ctx* ctx, float* cell_edge, int32_t* dims3) {
    if (!ctx || !ctx->has_map) return B200ICP_ERR_NO_MAP;
    if (cell_edge) *cell_edge = ctx->map.view.h;
    if (dims3) {
        dims3[0] = ctx->map.view.nx;
        dims3[1] = ctx->map.view.ny;
        dims3[2] = ctx->map.view.nz;
    }
    return B200ICP_OK;
}

int32_t b200icp_set_map_device(b200icp_ctx* ctx, const float* d_features, int32_t feature_rows, const float* d_normals,
                               int64_t n) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (feature_rows != ctx->cfg.dim + 1) return fail(ctx, B200ICP_ERR_INVALID_ARG, "feature_rows must be dim + 1");
    if (n < 0 || n > (int64_t)INT32_MAX - 1024) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad point count");
    if (n == 0) return B200ICP_OK;  // LPM: "Ignoring attempt to create a map from an empty cloud"
    if (!d_features) return fail(ctx, B200ICP_ERR_INVALID_ARG, "null features");
    CK(cudaSetDevice(ctx->device));
    DevCloud c;
    c.feat = d_features;
    c.rows = feature_rows;
    c.n = n;
    c.nrm = d_normals;
    CK(store_set(ctx->store, c, ctx->cfg.dim, ctx->stream));
    return commit_index(ctx);
}

int32_t b200icp_set_map(b200icp_ctx* ctx, const float* features, int32_t feature_rows, const float* normals, int64_t n) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (feature_rows != ctx->cfg.dim + 1) return fail(ctx, B200ICP_ERR_INVALID_ARG, "feature_rows must be dim + 1");
    if (n < 0 || n > (int64_t)INT32_MAX - 1024) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad point count");
    if (n == 0) return B200ICP_OK;
    if (!features) return fail(ctx, B200ICP_ERR_INVALID_ARG, "null features");
    CK(cudaSetDevice(ctx->device));
    const size_t fb = (size_t)n * feature_rows * sizeof(float);
    const size_t nb = (size_t)n * ctx->cfg.dim * sizeof(float);
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, fb));
    CK(cudaMemcpyAsync(ctx->d_stage_a, features, fb, cudaMemcpyHostToDevice, ctx->stream));
    if (normals) {
        CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, nb));
        CK(cudaMemcpyAsync(ctx->d_stage_b, normals, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    return b200icp_set_map_device(ctx, ctx->d_stage_a, feature_rows, normals ? ctx->d_stage_b : nullptr, n);
}

// The ICP loop on device-resident reading points.
static int32_t register_on_device(b200icp_ctx* ctx, const float* d_reading, int32_t rows, int64_t nq, const float* T_init,
                                  float* T_out, b200icp_result* result, const float* d_reading_normals = nullptr,
                                  const float* d_reading_max_dist = nullptr) {
    const int dim = ctx->cfg.dim;
    ctx->prm.rnrm = nullptr;
    const IcpParams& p = ctx->prm;
    IcpBuffers& b = ctx->buf;
    cudaStream_t s = ctx->stream;

    float Tinit4[16], Tmean[16], Tmean_inv[16], Tpre[16];
    embed(T_init, dim, Tinit4);
    mat4_identity(Tmean);
    mat4_identity(Tmean_inv);
    for (int d = 0; d < dim; ++d) {
        Tmean[12 + d] = ctx->map.mean[d];
        Tmean_inv[12 + d] = -ctx->map.mean[d];
    }
    mat4_mul(Tmean_inv, Tinit4, Tpre);
    if (std::fabs(1.f - det3(Tpre)) > 1e-3f)
        return fail(ctx, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
    if (p.minimizer == B200ICP_MIN_POINT_TO_PLANE && !ctx->map.has_normals)
        return fail(ctx, B200ICP_ERR_INVALID_FIELD, "Cannot find descriptor normals in reference (PointToPlaneErrorMinimizer)");
    for (int f = 0; f < p.n_outlier; ++f)
        if (p.outlier_kind[f] == B200ICP_OUTLIER_ROBUST && ((p.outlier_mode[f] >> 12) & 1) && !ctx->map.has_normals)
            return fail(ctx, B200ICP_ERR_INVALID_FIELD, "Cannot find descriptor normals in reference (RobustOutlierFilter, distanceType point2plane)");

    // state image
    IcpState* hs = reinterpret_cast<IcpState*>(ctx->h_pinned);
    memset(ctx->h_pinned, 0, kStateBytes);
    mat4_identity(hs->T);
    hs->nq = (int)nq;
    quat_from_M(hs->T, hs->dq[0]);
    hs->dcount = 1;
    quat_from_M(hs->T, hs->bq0);
    CK(cudaEventRecord(ctx->ev_begin, s));
    CK(cudaMemcpyAsync(b.state, ctx->h_pinned, kStateBytes, cudaMemcpyHostToDevice, s));
    int launches = 0;

    // reading -> refMean frame (+ optional cell sort for locality; results do not depend on order
    // except for the summation order of the error terms)
    const bool do_sort = ctx->cfg.sort_reading != 0 && nq > 1024;
    if (do_sort) {
        CK(ensure_scratch(ctx->map, nq));
        const GridView& v = ctx->map.view;
        // sort key: blocks of 8 x 8 x 4 cells (shift 3) -- half the radix passes of the full cell id and the loop is no
        // slower for it; nn_variant bits 12..14 = shift + 1 override it (1 = the full cell id)
        const int cs_bits = (ctx->cfg.nn_variant >> 12) & 7;
        const int cs = cs_bits ? cs_bits - 1 : 3;
        CK(launch_prep_reading(d_reading, rows, dim, Tpre, b.reading_tmp, &ctx->map.view, ctx->map.keys_in, ctx->map.vals_in, nq, s, cs, &b.state->pmax2_bits,
                               d_reading_max_dist, ctx->cfg.conventions & 1));
        const uint64_t n_keys = cs ? (uint64_t)(((v.nx - 1) >> cs) + 1) * (((v.ny - 1) >> cs) + 1) * (((v.nz - 1) >> std::max(cs - 1, 0)) + 1)
                                   : (uint64_t)v.nx * v.ny * v.nz;
        CK(sort_pairs(ctx->map, ctx->map.keys_in, ctx->map.keys_out, ctx->map.vals_in, ctx->map.vals_out, nq, bits_for(n_keys), s));
        CK(launch_gather_reading(b.reading_tmp, ctx->map.vals_out, b.reading, nq, s));
        launches += 4;
    } else {
        CK(launch_prep_reading(d_reading, rows, dim, Tpre, b.reading, nullptr, nullptr, nullptr, nq, s, 0, &b.state->pmax2_bits, d_reading_max_dist,
                               ctx->cfg.conventions & 1));
        launches += 1;
    }
    if (d_reading_normals) {  // the reading's `normals` descriptor follows the reading: same rotation, same order
        CK(launch_prep_normals(d_reading_normals, dim, Tpre, do_sort ? b.rnrm_tmp : b.rnrm, nq, s));
        if (do_sort) CK(launch_gather_reading(b.rnrm_tmp, ctx->map.vals_out, b.rnrm, nq, s));
        launches += do_sort ? 2 : 1;
        ctx->prm.rnrm = b.rnrm;
    }

    const bool fixed_count = p.max_iteration_count > 0 && !p.use_differential && !p.use_bound;
    const int hard_cap = p.max_iteration_count > 0 ? p.max_iteration_count : 10000;
    const int chunk = fixed_count ? hard_cap : 8;
    if (ctx->profiling && (int)ctx->nn_events.size() < 4 * hard_cap) {
        const size_t want = (size_t)4 * std::min(hard_cap, 512);
        while (ctx->nn_events.size() < want) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            ctx->nn_events.push_back(ev);
        }
    }
    constexpr int kTraceCap = 512;
    if (ctx->want_trace && !b.trace) CK(B200_CUDA_MALLOC((void**)&b.trace, (size_t)kTraceCap * 16 * sizeof(float)));
    float* const trace_keep = b.trace;
    if (!ctx->want_trace || hard_cap > kTraceCap) b.trace = nullptr;  // kernels see null -> no trace writes
    struct RestoreTrace {
        IcpBuffers& b;
        float* keep;
        ~RestoreTrace() { b.trace = keep; }
    } restore_trace{b, trace_keep};
    IcpState* out_state = reinterpret_cast<IcpState*>(ctx->h_pinned + kStateBytes);
    int issued = 0, nn_timed = 0;
    // Cold search for iteration 0, then the whole loop in one persistent cooperative kernel.
    // (Per-kernel profiling and nn_variant bit 2 select the kernel-per-step path below instead.)
    const bool var_trimmed = p.quantile_filter >= 0 && p.outlier_kind[p.quantile_filter] == B200ICP_OUTLIER_VAR_TRIMMED_DIST;
    // RobustOutlierFilter runs inside the loop kernel (its mad / berg scale through exact selects between grid barriers) unless the
    // scale is the standard deviation, or the chain also holds a quantile filter while the robust distance is point2plane under another
    // minimiser (the candidate tuples of the windowed iterations then carry no normal to recompute the weight from)
    bool robust = false;
    for (int f = 0; f < p.n_outlier; ++f)
        if (p.outlier_kind[f] == B200ICP_OUTLIER_ROBUST)
            robust = robust || ((p.outlier_mode[f] >> 8) & 15) == B200ICP_SCALE_STD || (ctx->cfg.nn_variant & 0x4000000) ||
                     (p.quantile_filter >= 0 && ((p.outlier_mode[f] >> 12) & 1) && p.minimizer != B200ICP_MIN_POINT_TO_PLANE);
    // (VarTrimmed -- and Robust in those cases -- estimate their ratio / scale with device-wide sorts between the steps: kernel-per-step path)
    // (a reading with a `maxSearchDist` descriptor: per-point radii ride in the reading's .w, which only the stand-alone search
    //  kernels read -- the sentinel max_r2 = -1 tells them to)
    const float search_r2 = d_reading_max_dist ? -1.f : p.max_r2;
    const bool persistent = !ctx->profiling && !(ctx->cfg.nn_variant & 4) && !var_trimmed && !robust && !d_reading_max_dist;
    if (persistent) {
        CK(launch_knn(ctx->map.view, b.reading, &b.state->nq, (int)nq, b.state, p.knn, p.max_r2, b.match_pos, b.match_d2,
                      /*want_original_ids=*/0, ctx->cfg.nn_variant & 0xffff, s));
        CK(cudaMemsetAsync(ctx->d_bar_counter, 0, sizeof(unsigned), s));
        {
            size_t zoff = 0, zbytes = 0;
            icp_loop_workspace_zero_range(&zoff, &zbytes);  // the fixed-point accumulators start from zero
            CK(cudaMemsetAsync(b.fastws + zoff, 0, zbytes, s));
        }
        CK(cudaEventRecord(ctx->ev_loop0, s));
        CK(launch_icp_loop(p, ctx->map, b, ctx->d_bar_counter, hard_cap, ctx->sm_share > 0 ? std::min(ctx->sm_share, ctx->n_sms) : ctx->n_sms,
                           ctx->cfg.nn_variant, ctx->win3, ctx->margin3, nq, s));
        CK(cudaEventRecord(ctx->ev_loop1, s));
        launches += 2;
        CK(cudaMemcpyAsync(out_state, b.state, kStateBytes, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->ev_end, s));
        CK(cudaStreamSynchronize(s));
    }
    while (!persistent) {
        const int upto = std::min(hard_cap, issued + chunk);
        for (; issued < upto; ++issued) {
            const bool time_it = ctx->profiling && (size_t)(4 * nn_timed + 3) < ctx->nn_events.size();
            cudaEvent_t* ev = time_it ? &ctx->nn_events[4 * nn_timed] : nullptr;
            if (time_it) CK(cudaEventRecord(ev[0], s));
            if (issued > 0 && p.knn == 1 && !(ctx->cfg.nn_variant & 2))  // warm: previous matches bound the search
                CK(launch_nn1_warm(ctx->map.view, b.reading, (int)nq, b.state, search_r2, b.match_pos, b.match_d2, ctx->cfg.nn_variant, s));
            else  // k > 1 from iteration 1 on: the previous matches bound the search (variant bit 16)
                CK(launch_knn(ctx->map.view, b.reading, &b.state->nq, (int)nq, b.state, p.knn, search_r2, b.match_pos, b.match_d2,
                              /*want_original_ids=*/0, (ctx->cfg.nn_variant & 0xffff) | ((issued > 0 && !(ctx->cfg.nn_variant & 2)) ? 0x10000 : 0), s));
            if (time_it) CK(cudaEventRecord(ev[1], s));
            ++launches;
            CK(launch_iteration_tail(p, ctx->map, b, issued, s, &launches, time_it ? ev[2] : nullptr, &ctx->var_scratch));
            if (time_it) {
                CK(cudaEventRecord(ev[3], s));
                ++nn_timed;
            }
        }
        CK(cudaMemcpyAsync(out_state, b.state, kStateBytes, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->ev_end, s));
        CK(cudaStreamSynchronize(s));
        if (out_state->done || issued >= hard_cap) break;
    }

    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end);
    ctx->timing.total_ms = ms;
    ctx->timing.kernel_launches = launches;
    ctx->timing.nn_ms_sum = ctx->timing.select_ms_sum = ctx->timing.acc_ms_sum = 0.f;
    ctx->timing.loop_iterations = persistent ? out_state->loop_iters_timed : 0;
    ctx->timing.loop_search_ms_sum = persistent ? (float)(1e-6 * (double)out_state->loop_search_ns) : 0.f;
    ctx->timing.loop_total_ms = persistent ? (float)(1e-6 * (double)out_state->loop_total_ns) : 0.f;
    ctx->timing.loop_kernel_ms = 0.f;
    if (persistent) cudaEventElapsedTime(&ctx->timing.loop_kernel_ms, ctx->ev_loop0, ctx->ev_loop1);
    ctx->timing.loop_fast_iterations = persistent ? out_state->fast_iters : 0;
    ctx->timing.loop_searched_queries = persistent ? out_state->searched_queries : 0;
    ctx->timing.loop_two_barrier_iterations = persistent ? out_state->hist_iters : 0;
    ctx->timing.nn_launches = 0;
    const int executed = out_state->iter;
    for (int i = 0; i < nn_timed && i < std::max(executed, 1); ++i) {
        float t = 0.f;
        cudaEvent_t* ev = &ctx->nn_events[4 * i];
        if (cudaEventElapsedTime(&t, ev[0], ev[1]) == cudaSuccess) {
            ctx->timing.nn_ms_sum += t;
            ctx->timing.nn_launches += 1;
        }
        if (cudaEventElapsedTime(&t, ev[1], ev[2]) == cudaSuccess) ctx->timing.select_ms_sum += t;
        if (cudaEventElapsedTime(&t, ev[2], ev[3]) == cudaSuccess) ctx->timing.acc_ms_sum += t;
    }
    ctx->h_trace.clear();
    if (b.trace && executed > 0) {
        ctx->h_trace.resize((size_t)executed * 16);
        CK(cudaMemcpyAsync(ctx->h_trace.data(), b.trace, ctx->h_trace.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    if (result) {
        result->overlap = out_state->overlap;
        result->point_used_ratio = out_state->used_ratio;
        result->iterations = out_state->iter;
        result->max_iter_reached = out_state->max_iter_reached;
        result->pairs_last_iter = out_state->pairs;
    }
    // return T_refIn_refMean * T_iter * T_refMean_dataIn
    float tmp[16], Tfull[16];
    mat4_mul(out_state->T, Tpre, tmp);
    mat4_mul(Tmean, tmp, Tfull);
    extract(Tfull, dim, T_out);
    switch (out_state->status) {
        case B200ICP_OK: return B200ICP_OK;
        case B200ICP_ERR_CONVERGENCE: return fail(ctx, B200ICP_ERR_CONVERGENCE, "ConvergenceError: no point to minimize / no outlier to filter");
        case B200ICP_ERR_BOUND: return fail(ctx, B200ICP_ERR_BOUND, "ConvergenceError: limit out of bounds");
        case B200ICP_ERR_NAN: return fail(ctx, B200ICP_ERR_NAN, "ConvergenceError: abs rotation/translation norm not a number");
        case B200ICP_ERR_TRANSFORM: return fail(ctx, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
        case B200ICP_ERR_NOT_IMPLEMENTED:
            return fail(ctx, B200ICP_ERR_NOT_IMPLEMENTED, "the error sums left the fixed-point range of the loop kernel (coordinates or normals of absurd magnitude)");
        default: return fail(ctx, out_state->status, "device reported an error");
    }
}

static int32_t ensure_reading_normals(b200icp_ctx* ctx, int64_t nq) {
    IcpBuffers& b = ctx->buf;
    if (nq <= b.cap_rnrm) return B200ICP_OK;
    B200_CUDA_FREE(b.rnrm_in);
    B200_CUDA_FREE(b.rnrm);
    B200_CUDA_FREE(b.rnrm_tmp);
    b.rnrm_in = nullptr;
    b.rnrm = b.rnrm_tmp = nullptr;
    b.cap_rnrm = 0;
    const int64_t cap = grow_capacity(nq);
    CK(B200_CUDA_MALLOC((void**)&b.rnrm_in, (size_t)cap * ctx->cfg.dim * sizeof(float)));
    CK(B200_CUDA_MALLOC((void**)&b.rnrm, (size_t)cap * sizeof(float4)));
    CK(B200_CUDA_MALLOC((void**)&b.rnrm_tmp, (size_t)cap * sizeof(float4)));
    b.cap_rnrm = cap;
    return B200ICP_OK;
}

static int32_t register_checks(b200icp_ctx* ctx, const float* reading, int32_t rows, int64_t nq, float* T_out) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    if (!T_out) return fail(ctx, B200ICP_ERR_INVALID_ARG, "null T_out");
    if (rows != ctx->cfg.dim + 1) return fail(ctx, B200ICP_ERR_INVALID_ARG, "feature_rows must be dim + 1");
    if (nq < 0 || nq > (int64_t)INT32_MAX / 64) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad point count");
    if (nq > 0 && !reading) return fail(ctx, B200ICP_ERR_INVALID_ARG, "null reading");
    if (!ctx->has_map) {  // LPM ICPSequence::operator(): no map -> identity
        float I4[16];
        mat4_identity(I4);
        extract(I4, ctx->cfg.dim, T_out);
        return fail(ctx, B200ICP_ERR_NO_MAP, "no map: call b200icp_set_map first");
    }
    if (nq == 0) return fail(ctx, B200ICP_ERR_CONVERGENCE, "ConvergenceError: empty reading (no point to minimize)");
    return B200ICP_OK;
}

int32_t b200icp_register_device(b200icp_ctx* ctx, const float* d_reading, int32_t feature_rows, int64_t nq, const float* T_init,
                                float* T_out, b200icp_result* result) {
    if (result) memset(result, 0, sizeof(*result));
    const int32_t rc = register_checks(ctx, d_reading, feature_rows, nq, T_out);
    if (rc != B200ICP_OK) return rc;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    const int32_t eb = ensure_icp_buffers(ctx, nq);
    if (eb != B200ICP_OK) return eb;
    return register_on_device(ctx, d_reading, feature_rows, nq, T_init, T_out, result);
}

int32_t b200icp_register_normals(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq, const float* reading_normals,
                                 const float* T_init, float* T_out, b200icp_result* result) {
    if (!reading_normals) return b200icp_register(ctx, reading, feature_rows, nq, T_init, T_out, result);
    if (result) memset(result, 0, sizeof(*result));
    const int32_t rc = register_checks(ctx, reading, feature_rows, nq, T_out);
    if (rc != B200ICP_OK) return rc;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    const int32_t eb = ensure_icp_buffers(ctx, nq);
    if (eb != B200ICP_OK) return eb;
    IcpBuffers& b = ctx->buf;
    const int dim = ctx->cfg.dim;
    const int32_t en = ensure_reading_normals(ctx, nq);
    if (en != B200ICP_OK) return en;
    CK(cudaMemcpyAsync(b.reading_in, reading, (size_t)nq * feature_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.rnrm_in, reading_normals, (size_t)nq * dim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return register_on_device(ctx, b.reading_in, feature_rows, nq, T_init, T_out, result, b.rnrm_in);
}

int32_t b200icp_register_descriptors(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq, const float* reading_normals,
                                     const float* reading_max_search_dist, const float* T_init, float* T_out, b200icp_result* result) {
    if (!reading_max_search_dist) return b200icp_register_normals(ctx, reading, feature_rows, nq, reading_normals, T_init, T_out, result);
    if (result) memset(result, 0, sizeof(*result));
    const int32_t rc = register_checks(ctx, reading, feature_rows, nq, T_out);
    if (rc != B200ICP_OK) return rc;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    const int32_t eb = ensure_icp_buffers(ctx, nq);
    if (eb != B200ICP_OK) return eb;
    IcpBuffers& b = ctx->buf;
    const int dim = ctx->cfg.dim;
    if (nq > b.cap_rmax) {
        B200_CUDA_FREE(b.rmax_in);
        b.rmax_in = nullptr;
        b.cap_rmax = 0;
        CK(B200_CUDA_MALLOC((void**)&b.rmax_in, (size_t)grow_capacity(nq) * sizeof(float)));
        b.cap_rmax = grow_capacity(nq);
    }
    if (reading_normals) {
        const int32_t en = ensure_reading_normals(ctx, nq);
        if (en != B200ICP_OK) return en;
    }
    CK(cudaMemcpyAsync(b.reading_in, reading, (size_t)nq * feature_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.rmax_in, reading_max_search_dist, (size_t)nq * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (reading_normals) CK(cudaMemcpyAsync(b.rnrm_in, reading_normals, (size_t)nq * dim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return register_on_device(ctx, b.reading_in, feature_rows, nq, T_init, T_out, result, reading_normals ? b.rnrm_in : nullptr, b.rmax_in);
}

int32_t b200icp_register(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq, const float* T_init,
                         float* T_out, b200icp_result* result) {
    if (result) memset(result, 0, sizeof(*result));
    const int32_t rc = register_checks(ctx, reading, feature_rows, nq, T_out);
    if (rc != B200ICP_OK) return rc;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    const int32_t eb = ensure_icp_buffers(ctx, nq);
    if (eb != B200ICP_OK) return eb;
    CK(cudaMemcpyAsync(ctx->buf.reading_in, reading, (size_t)nq * feature_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return register_on_device(ctx, ctx->buf.reading_in, feature_rows, nq, T_init, T_out, result);
}

int32_t b200icp_register_batch(b200icp_ctx* const* ctxs, int32_t n_ctx, const b200icp_pair* pairs, int64_t n_pairs,
                               b200icp_pair_result* out) {
    if (!ctxs || n_ctx < 1 || n_pairs < 0 || (n_pairs > 0 && (!pairs || !out))) return B200ICP_ERR_INVALID_ARG;
    for (int c = 0; c < n_ctx; ++c) {
        if (!ctxs[c]) return B200ICP_ERR_INVALID_ARG;
        if (ctxs[c]->cfg.dim != ctxs[0]->cfg.dim) return fail(ctxs[0], B200ICP_ERR_INVALID_ARG, "register_batch: contexts differ in dim");
        for (int d = 0; d < c; ++d)
            if (ctxs[d] == ctxs[c]) return fail(ctxs[0], B200ICP_ERR_INVALID_ARG, "register_batch: the same context listed twice");
    }
    const int dim = ctxs[0]->cfg.dim, rows = dim + 1;
    // Contexts that share a GPU split its SMs: each registration loop is a persistent kernel with one CTA per SM it may use, so
    // two half-size loops run side by side (their barrier waits and serial phases overlap) instead of one after the other.
    std::vector<int> saved_share(n_ctx);
    for (int c = 0; c < n_ctx; ++c) {
        saved_share[c] = ctxs[c]->sm_share;
        int on_device = 0;
        for (int d = 0; d < n_ctx; ++d) on_device += ctxs[d]->device == ctxs[c]->device;
        if (on_device > 1 && n_pairs > 1 && saved_share[c] == 0) ctxs[c]->sm_share = std::max(1, ctxs[c]->n_sms / on_device);
    }
    auto work = [&](int c) {
        b200icp_ctx* ctx = ctxs[c];
        for (int64_t j = c; j < n_pairs; j += n_ctx) {
            const b200icp_pair& pr = pairs[j];
            b200icp_pair_result& r = out[j];
            memset(&r, 0, sizeof(r));
            int32_t rc = B200ICP_OK;
            if (pr.map_features) {
                rc = b200icp_set_map(ctx, pr.map_features, rows, pr.map_normals, pr.n_map);
                r.setmap_ms = ctx->timing.setmap_ms;
            }
            if (rc == B200ICP_OK) {
                rc = b200icp_register(ctx, pr.reading, rows, pr.n_reading, pr.T_init, r.T, &r.result);
                r.register_ms = ctx->timing.total_ms;
            }
            r.status = rc;
        }
    };
    if (n_ctx == 1 || n_pairs <= 1) {
        for (int c = 0; c < n_ctx; ++c) work(c);
    } else {
        std::vector<std::thread> th;
        th.reserve(n_ctx - 1);
        for (int c = 1; c < n_ctx; ++c) th.emplace_back(work, c);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int c = 0; c < n_ctx; ++c) ctxs[c]->sm_share = saved_share[c];
    for (int64_t j = 0; j < n_pairs; ++j)
        if (out[j].status != B200ICP_OK) return out[j].status;
    return B200ICP_OK;
}

static int32_t run_queries(b200icp_ctx* ctx, GridIndex& g, const float* queries, int32_t rows, int64_t nq, int dim, int k,
                           float max_r2, bool centre, int32_t* ids, float* dists2) {
    const int32_t eb = ensure_query_buffers(ctx, nq, k);
    if (eb != B200ICP_OK) return eb;
    cudaStream_t s = ctx->stream;
    const size_t qb = (size_t)nq * rows * sizeof(float);
    CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, qb));
    CK(cudaMemcpyAsync(ctx->d_stage_b, queries, qb, cudaMemcpyHostToDevice, s));
    float Tpre[16];
    mat4_identity(Tpre);
    if (centre)
        for (int d = 0; d < dim; ++d) Tpre[12 + d] = -g.mean[d];
    CK(launch_prep_reading(ctx->d_stage_b, rows, dim, Tpre, ctx->d_q4, nullptr, nullptr, nullptr, nq, s));
    int* h_nq = reinterpret_cast<int*>(ctx->h_pinned + 2 * kStateBytes);
    *h_nq = (int)nq;
    CK(cudaMemcpyAsync(ctx->d_scalar_nq, h_nq, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(launch_knn(g.view, ctx->d_q4, ctx->d_scalar_nq, (int)nq, nullptr, k, max_r2, ctx->d_out_ids, ctx->d_out_d2,
                  /*want_original_ids=*/1, ctx->cfg.nn_variant, s));
    CK(cudaMemcpyAsync(ids, ctx->d_out_ids, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(dists2, ctx->d_out_d2, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int32_t b200icp_match(b200icp_ctx* ctx, const float* queries, int32_t feature_rows, int64_t nq, int32_t* ids, float* dists2) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (feature_rows != ctx->cfg.dim + 1) return fail(ctx, B200ICP_ERR_INVALID_ARG, "feature_rows must be dim + 1");
    if (nq < 0 || (nq > 0 && (!queries || !ids || !dists2))) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad arguments");
    if (!ctx->has_map) return fail(ctx, B200ICP_ERR_NO_MAP, "no map: call b200icp_set_map first");
    if (nq == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    return run_queries(ctx, ctx->map, queries, feature_rows, nq, ctx->cfg.dim, ctx->cfg.knn, ctx->prm.max_r2, true, ids, dists2);
}

int32_t b200icp_knn(b200icp_ctx* ctx, const float* ref, int32_t ref_rows, int64_t nref, const float* queries, int32_t query_rows,
                    int64_t nq, int32_t dim, int32_t k, float max_radius, int32_t* ids, float* dists2) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if ((dim != 2 && dim != 3) || ref_rows < dim || query_rows < dim) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad dim / rows");
    if (k < 1 || k > 32) return fail(ctx, B200ICP_ERR_INVALID_ARG, "k must be in [1, 32]");
    if (nref <= 0 || nref > (int64_t)INT32_MAX - 1024 || !ref) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad reference cloud");
    if (nq < 0 || (nq > 0 && (!queries || !ids || !dists2))) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad queries");
    if (!(max_radius > 0.f)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "max_radius must be positive (inf allowed)");
    if (nq == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t rb = (size_t)nref * ref_rows * sizeof(float);
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, rb));
    CK(cudaMemcpyAsync(ctx->d_stage_a, ref, rb, cudaMemcpyHostToDevice, ctx->stream));
    CK(grid_build(ctx->aux, ctx->d_stage_a, ref_rows, dim, nullptr, nref, /*centre=*/false, 0.f, ctx->stream));
    const float max_r2 = std::isinf(max_radius) ? INFINITY : max_radius * max_radius;
    return run_queries(ctx, ctx->aux, queries, query_rows, nq, dim, k, max_r2, false, ids, dists2);
}

int32_t b200icp_transform(b200icp_ctx* ctx, float* features, int32_t feature_rows, float* normals, int64_t n, const float* T) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = feature_rows - 1;
    if ((dim != 2 && dim != 3) || !T || n < 0 || (n > 0 && !features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad arguments");
    float M[16];
    embed(T, dim, M);
    if (std::fabs(1.f - det3(M)) > 1e-3f)
        return fail(ctx, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
    if (n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t fb = (size_t)n * feature_rows * sizeof(float), nb = (size_t)n * dim * sizeof(float);
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, fb));
    CK(cudaMemcpyAsync(ctx->d_stage_a, features, fb, cudaMemcpyHostToDevice, s));
    if (normals) {
        CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, nb));
        CK(cudaMemcpyAsync(ctx->d_stage_b, normals, nb, cudaMemcpyHostToDevice, s));
    }
    CK(launch_transform(ctx->d_stage_a, feature_rows, dim, normals ? ctx->d_stage_b : nullptr, n, M, s));
    CK(cudaMemcpyAsync(features, ctx->d_stage_a, fb, cudaMemcpyDeviceToHost, s));
    if (normals) CK(cudaMemcpyAsync(normals, ctx->d_stage_b, nb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int32_t b200icp_transform_device(b200icp_ctx* ctx, float* d_features, int32_t feature_rows, float* d_normals, int64_t n,
                                 const float* T) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = feature_rows - 1;
    if ((dim != 2 && dim != 3) || !T || n < 0 || (n > 0 && !d_features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad arguments");
    float M[16];
    embed(T, dim, M);
    if (std::fabs(1.f - det3(M)) > 1e-3f)
        return fail(ctx, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
    CK(cudaSetDevice(ctx->device));
    CK(launch_transform(d_features, feature_rows, dim, d_normals, n, M, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

/* ---- device-resident map: Map::updateLocalPointCloud / updatePose pieces -------------------------- */

int32_t b200icp_map_reserve(b200icp_ctx* ctx, int64_t n_points, int32_t normals_knn) {
    if (!ctx || n_points < 0 || n_points > (int64_t)INT32_MAX / 2) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    const int dim = ctx->cfg.dim;
    CK(store_reserve(ctx->store, dim, n_points, ctx->stream));
    CK(store_reserve_scratch(ctx->store, n_points));
    // Index buffers: the map's own grid, and the auxiliary grid the update steps build over the changed points (a window move
    // changes up to a quarter of the local map at once).  Releasing any of them later is a cudaFree on the update path, and
    // a cudaFree next to gigabytes of reserved buffers was measured at 70-870 ms (B200ICP_TRACE_ALLOC=1).
    auto reserve_grid = [&](GridIndex& g, int64_t n, bool with_normals) -> int32_t {
        CK(ensure_scratch(g, n));
        if (n > g.cap_pts) {
            B200_CUDA_FREE(g.pts);
            g.pts = nullptr;
            g.cap_pts = 0;
            CK(B200_CUDA_MALLOC((void**)&g.pts, (size_t)grow_capacity(n) * sizeof(float4)));
            g.cap_pts = grow_capacity(n);
        }
        if (with_normals && n > g.cap_normals) {
            B200_CUDA_FREE(g.normals);
            g.normals = nullptr;
            g.cap_normals = 0;
            CK(B200_CUDA_MALLOC((void**)&g.normals, (size_t)grow_capacity(n) * sizeof(float4)));
            g.cap_normals = grow_capacity(n);
        }
        const int64_t want_cells = std::min<int64_t>(4 * n + 2, (int64_t)kMaxCells + 2);  // about 4 cells per point, capped
        if (want_cells > g.cap_cells) {
            B200_CUDA_FREE(g.cell_start);
            g.cell_start = nullptr;
            g.cap_cells = 0;
            CK(B200_CUDA_MALLOC((void**)&g.cell_start, (size_t)want_cells * sizeof(uint32_t)));
            g.cap_cells = want_cells;
        }
        return B200ICP_OK;
    };
    int32_t rg = reserve_grid(ctx->map, n_points, true);
    if (rg != B200ICP_OK) return rg;
    rg = reserve_grid(ctx->aux, std::max<int64_t>(n_points / 4, 1), false);
    if (rg != B200ICP_OK) return rg;
    if (normals_knn > 0) {
        const int32_t eb = ensure_query_buffers(ctx, n_points, normals_knn);
        if (eb != B200ICP_OK) return eb;
        // bookkeeping of the incremental SurfaceNormal pass (b200icp_map_surface_normals)
        if (n_points > ctx->cap_kth && !ctx->d_kth) {
            CK(B200_CUDA_MALLOC((void**)&ctx->d_kth, (size_t)n_points * sizeof(float)));
            ctx->cap_kth = n_points;
        }
        if (2 * n_points > ctx->cap_dirty && !ctx->d_dirty && !ctx->d_list) {
            CK(B200_CUDA_MALLOC((void**)&ctx->d_dirty, (size_t)(2 * n_points)));
            CK(B200_CUDA_MALLOC((void**)&ctx->d_list, (size_t)(2 * n_points) * sizeof(uint32_t)));
            ctx->cap_dirty = 2 * n_points;
        }
    }
    return B200ICP_OK;
}

int32_t b200icp_map_commit(b200icp_ctx* ctx) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    return commit_index(ctx);
}

/* Online mapping: see the comment on b200icp_ctx::work. */
int32_t b200icp_map_begin_update(b200icp_ctx* ctx) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (ctx->updating) return fail(ctx, B200ICP_ERR_INVALID_ARG, "a map update is already in progress on this context");
    ctx->updating = true;
    ctx->work_valid = false;
    t_update_ctx = ctx;
    return B200ICP_OK;
}

int32_t b200icp_map_end_update(b200icp_ctx* ctx) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (!ctx->updating) return fail(ctx, B200ICP_ERR_INVALID_ARG, "no map update in progress");
    if (t_update_ctx == ctx) t_update_ctx = nullptr;
    if (ctx->index_stale && ctx->store.n_active > 0) {  // (the steps left the index behind the store: the published map is the final one)
        const int32_t rc = commit_index(ctx);
        if (rc != B200ICP_OK) {
            ctx->updating = false;
            ctx->work_valid = false;
            return rc;
        }
    }
    if (ctx->work_valid) {
        CK(cudaStreamSynchronize(ctx->stream));
        std::swap(ctx->map, ctx->work);  // icp.setMap(localPointCloud): registrations see the new map from here on
        ctx->has_map = true;
        ctx->map_n = ctx->store.n_active;
    }
    ctx->updating = false;
    ctx->work_valid = false;
    return B200ICP_OK;
}

int32_t b200icp_map_update_in_progress(const b200icp_ctx* ctx) { return (ctx && ctx->updating) ? 1 : 0; }

/* The scan in the slot becomes the input of an asynchronous map update (the reference copies `currentInput` into the std::async
 * call, Mapper.cpp:282): the thread that calls b200icp_map_begin_update works on it through the b200icp_scan_* module entry
 * points while the caller thread is free to upload the next scan.  O(1): the two slots trade places; the caller's slot is empty
 * afterwards. */
int32_t b200icp_scan_snapshot(b200icp_ctx* ctx) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (ctx->updating) return fail(ctx, B200ICP_ERR_INVALID_ARG, "the previous snapshot is still in use by a map update");
    CK(cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->scan, ctx->scan_upd);
    ctx->scan.n = 0;
    ctx->scan.has_nrm = ctx->scan.has_prob = false;
    ctx->scan.extra_rows = 0;
    ctx->scan.n_rot = 0;
    return B200ICP_OK;
}

int32_t b200icp_map_counts(const b200icp_ctx* ctx, int64_t* n_local, int64_t* n_global) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    if (n_local) *n_local = ctx->store.n_active;
    if (n_global) *n_global = ctx->store.n;
    return B200ICP_OK;
}

int32_t b200icp_map_has_normals(const b200icp_ctx* ctx) { return (ctx && ctx->store.has_normals) ? 1 : 0; }

// PointDistance insert of a cloud that is already on the device (map frame)
static int32_t insert_point_distance_dev(b200icp_ctx* ctx, const DevCloud& in, float min_dist_new_point, int64_t* n_added, uint8_t* keep_out) {
    const int dim = ctx->cfg.dim;
    cudaStream_t s = ctx->stream;
    MapStore& st = ctx->store;
    const int64_t n_in = in.n;
    const int32_t eb = ensure_query_buffers(ctx, n_in, 1);
    if (eb != B200ICP_OK) return eb;
    if (st.n_active > 0) {
        // Nabo::NNS::knn(input, k = 1, eps = 0, no radius) against the map -- on the live index
        float Tpre[16];
        mat4_identity(Tpre);
        for (int d = 0; d < dim; ++d) Tpre[12 + d] = -synced_index(ctx).mean[d];
        CK(launch_prep_reading(in.feat, in.rows, dim, Tpre, ctx->d_q4, nullptr, nullptr, nullptr, n_in, s));
        int* h_nq = reinterpret_cast<int*>(ctx->h_pinned + 2 * kStateBytes);
        *h_nq = (int)n_in;
        CK(cudaMemcpyAsync(ctx->d_scalar_nq, h_nq, sizeof(int), cudaMemcpyHostToDevice, s));
        // Only "is there a map point closer than minDistNewPoint" matters (the keep test below re-evaluates the distance to the point
        // found, in map-frame coordinates): the search is bounded by that radius, 0.1 % wider than the threshold so that the rounding
        // of the centred frame cannot hide a point the exact test would reject.  No neighbour inside = keep, as with an unbounded search.
        const float r2 = min_dist_new_point > 0.f ? min_dist_new_point * min_dist_new_point * 1.001f + 1e-30f : 0.f;
        CK(launch_knn(synced_index(ctx).view, ctx->d_q4, ctx->d_scalar_nq, (int)n_in, nullptr, 1, (ctx->cfg.nn_variant & 0x2000000) ? INFINITY : r2,
                      ctx->d_out_ids, ctx->d_out_d2, /*want_original_ids=*/1, ctx->cfg.nn_variant & 0xffff, s));
    } else {  // createMap: the first cloud is taken as it is (PointDistanceMapperModule.cpp:9-19)
        CK(cudaMemsetAsync(ctx->d_out_ids, 0xff, (size_t)n_in * sizeof(int32_t), s));
    }
    if (keep_out && n_in > ctx->cap_keep) {
        B200_CUDA_FREE(ctx->d_keep);
        ctx->d_keep = nullptr;
        ctx->cap_keep = 0;
        CK(B200_CUDA_MALLOC((void**)&ctx->d_keep, (size_t)(n_in + n_in / 4 + 1024)));
        ctx->cap_keep = n_in + n_in / 4 + 1024;
    }
    int64_t kept = 0;
    CK(store_insert_point_distance(st, build_index(ctx), in, dim, ctx->d_out_ids, min_dist_new_point, &kept, keep_out ? ctx->d_keep : nullptr, s));
    if (keep_out) CK(cudaMemcpyAsync(keep_out, ctx->d_keep, (size_t)n_in, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (kept > 0) ctx->index_stale = true;
    if (n_added) *n_added = kept;
    return B200ICP_OK;
}

// host cloud (features + the two descriptors the host-pointer entry points know) -> the upload staging buffer
static int32_t upload_input(b200icp_ctx* ctx, const float* input, int rows, int64_t n_in, const float* nrm, const float* prob, DevCloud* out) {
    const int dim = ctx->cfg.dim;
    const size_t fb = (((size_t)n_in * rows * sizeof(float)) + 255) / 256 * 256;
    const size_t nb = (((size_t)n_in * dim * sizeof(float)) + 255) / 256 * 256;
    const size_t pb = (((size_t)n_in * sizeof(float)) + 255) / 256 * 256;
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, fb + nb + pb + 256));
    char* base = reinterpret_cast<char*>(ctx->d_stage_a);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_nrm = nrm ? reinterpret_cast<float*>(base + fb) : nullptr;
    float* d_prob = prob ? reinterpret_cast<float*>(base + fb + nb) : nullptr;
    CK(cudaMemcpyAsync(d_in, input, (size_t)n_in * rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (nrm) CK(cudaMemcpyAsync(d_nrm, nrm, (size_t)n_in * dim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (prob) CK(cudaMemcpyAsync(d_prob, prob, (size_t)n_in * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    *out = DevCloud();
    out->feat = d_in;
    out->rows = rows;
    out->n = n_in;
    out->nrm = d_nrm;
    out->prob = d_prob;
    return B200ICP_OK;
}

int32_t b200icp_map_insert_point_distance(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                                          const float* input_normals, float min_dist_new_point, int64_t* n_added, uint8_t* keep_out) {
    return b200icp_map_insert_point_distance_prob(ctx, input, feature_rows, n_in, input_normals, nullptr, min_dist_new_point, n_added, keep_out);
}

int32_t b200icp_map_insert_point_distance_prob(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                                               const float* input_normals, const float* input_prob, float min_dist_new_point,
                                               int64_t* n_added, uint8_t* keep_out) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1) return fail(ctx, B200ICP_ERR_INVALID_ARG, "feature_rows must be dim + 1");
    if (n_in < 0 || (n_in > 0 && !input)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad input cloud");
    if (n_added) *n_added = 0;
    if (n_in == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->store.n_active > 0 && ctx->index_stale) {
        const int32_t rc = commit_index(ctx);
        if (rc != B200ICP_OK) return rc;
    }
    DevCloud in;
    const int32_t rc = upload_input(ctx, input, feature_rows, n_in, input_normals, input_prob, &in);
    if (rc != B200ICP_OK) return rc;
    return insert_point_distance_dev(ctx, in, min_dist_new_point, n_added, keep_out);
}

/* ---- the device-resident scan slot: the raw scan is uploaded ONCE with its descriptors; the `input:` filter
 * chain (Mapper.cpp:187-191), the rigid transforms (Mapper.cpp:197,221), icp(input) (:213) and the MapperModules
 * (Map.cpp:502-534) then run on that copy ---- */
static int32_t scan_reserve(b200icp_ctx* ctx, int64_t n, int extra_rows) {
    auto& sc = scan_slot(ctx);
    const int dim = ctx->cfg.dim;
    if (n > sc.cap) {
        const int64_t cap = grow_capacity(n);
        for (int i = 0; i < 2; ++i) {
            B200_CUDA_FREE(sc.feat[i]);
            B200_CUDA_FREE(sc.nrm[i]);
            B200_CUDA_FREE(sc.prob[i]);
            B200_CUDA_FREE(sc.extra[i]);
            sc.feat[i] = sc.nrm[i] = sc.prob[i] = sc.extra[i] = nullptr;
        }
        sc.cap = 0;
        sc.cap_extra = 0;
        for (int i = 0; i < 2; ++i) {
            CK(B200_CUDA_MALLOC((void**)&sc.feat[i], (size_t)cap * (dim + 1) * sizeof(float)));
            CK(B200_CUDA_MALLOC((void**)&sc.nrm[i], (size_t)cap * dim * sizeof(float)));
            CK(B200_CUDA_MALLOC((void**)&sc.prob[i], (size_t)cap * sizeof(float)));
        }
        sc.cap = cap;
    }
    if (extra_rows > 0 && sc.cap * extra_rows > sc.cap_extra) {
        for (int i = 0; i < 2; ++i) {
            B200_CUDA_FREE(sc.extra[i]);
            sc.extra[i] = nullptr;
        }
        sc.cap_extra = 0;
        for (int i = 0; i < 2; ++i) CK(B200_CUDA_MALLOC((void**)&sc.extra[i], (size_t)sc.cap * extra_rows * sizeof(float)));
        sc.cap_extra = sc.cap * extra_rows;
    }
    return B200ICP_OK;
}

static DevCloud scan_cloud(const b200icp_ctx* ctx) {
    const auto& sc = scan_slot(ctx);
    DevCloud c;
    c.feat = sc.feat[sc.cur];
    c.rows = ctx->cfg.dim + 1;
    c.n = sc.n;
    c.nrm = sc.has_nrm ? sc.nrm[sc.cur] : nullptr;
    c.prob = sc.has_prob ? sc.prob[sc.cur] : nullptr;
    c.extra = sc.extra_rows > 0 ? sc.extra[sc.cur] : nullptr;
    c.extra_rows = sc.extra_rows;
    return c;
}

int32_t b200icp_scan_upload(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (feature_rows != ctx->cfg.dim + 1 || n < 0 || (n > 0 && !features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad scan");
    CK(cudaSetDevice(ctx->device));
    const int32_t rc = scan_reserve(ctx, n, 0);
    if (rc != B200ICP_OK) return rc;
    auto& sc = scan_slot(ctx);
    if (n > 0) CK(cudaMemcpyAsync(sc.feat[sc.cur], features, (size_t)n * feature_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    sc.n = n;
    sc.has_nrm = sc.has_prob = false;  // a new cloud: the previous scan's descriptors are gone
    sc.extra_rows = 0;
    sc.n_rot = 0;
    return B200ICP_OK;
}

int32_t b200icp_scan_set_descriptors(b200icp_ctx* ctx, const float* normals, const float* prob, const float* extra, int32_t extra_rows,
                                     const int32_t* rotating_rows, int32_t n_rotating) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (extra && (extra_rows < 1 || extra_rows > B200ICP_MAX_EXTRA_ROWS)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "extra_rows out of range");
    if (n_rotating < 0 || n_rotating > 4 || (n_rotating > 0 && (!rotating_rows || !extra)))
        return fail(ctx, B200ICP_ERR_INVALID_ARG, "at most 4 rotating descriptors, inside `extra`");
    for (int i = 0; i < n_rotating; ++i)
        if (rotating_rows[i] < 0 || rotating_rows[i] + dim > extra_rows) return fail(ctx, B200ICP_ERR_INVALID_ARG, "rotating descriptor outside `extra`");
    CK(cudaSetDevice(ctx->device));
    auto& sc = scan_slot(ctx);
    const int32_t rc = scan_reserve(ctx, sc.n, extra ? extra_rows : 0);
    if (rc != B200ICP_OK) return rc;
    cudaStream_t s = ctx->stream;
    const int64_t n = sc.n;
    if (normals && n > 0) CK(cudaMemcpyAsync(sc.nrm[sc.cur], normals, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice, s));
    if (prob && n > 0) CK(cudaMemcpyAsync(sc.prob[sc.cur], prob, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
    if (extra && n > 0) CK(cudaMemcpyAsync(sc.extra[sc.cur], extra, (size_t)n * extra_rows * sizeof(float), cudaMemcpyHostToDevice, s));
    sc.has_nrm = normals != nullptr;
    sc.has_prob = prob != nullptr;
    sc.extra_rows = extra ? extra_rows : 0;
    sc.n_rot = extra ? n_rotating : 0;
    for (int i = 0; i < sc.n_rot; ++i) sc.rot_row[i] = rotating_rows[i];
    return B200ICP_OK;
}

int64_t b200icp_scan_size(const b200icp_ctx* ctx) { return ctx ? scan_slot(ctx).n : 0; }

int32_t b200icp_scan_info(const b200icp_ctx* ctx, int32_t* has_normals, int32_t* has_prob, int32_t* extra_rows) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    if (has_normals) *has_normals = scan_slot(ctx).has_nrm ? 1 : 0;
    if (has_prob) *has_prob = scan_slot(ctx).has_prob ? 1 : 0;
    if (extra_rows) *extra_rows = scan_slot(ctx).extra_rows;
    return B200ICP_OK;
}

int32_t b200icp_scan_transform(b200icp_ctx* ctx, const float* T) {
    if (!ctx || !T) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    float M[16];
    embed(T, dim, M);
    if (std::fabs(1.f - det3(M)) > 1e-3f)
        return fail(ctx, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
    auto& sc = scan_slot(ctx);
    if (sc.n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    // stream-ordered: no sync needed here
    CK(launch_transform(sc.feat[sc.cur], dim + 1, dim, sc.has_nrm ? sc.nrm[sc.cur] : nullptr, sc.n, M, ctx->stream));
    // descriptors named `observationDirections` rotate like normals (LPM TransformationsImpl.cpp); the others are left alone
    for (int i = 0; i < sc.n_rot; ++i) CK(launch_rotate_rows(sc.extra[sc.cur], sc.extra_rows, sc.rot_row[i], dim, sc.n, M, ctx->stream));
    return B200ICP_OK;
}

int32_t b200icp_scan_register(b200icp_ctx* ctx, const float* T_init, float* T_out, b200icp_result* result) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    auto& sc = scan_slot(ctx);
    if (!sc.has_nrm) return b200icp_register_device(ctx, sc.feat[sc.cur], ctx->cfg.dim + 1, sc.n, T_init, T_out, result);
    // the reading carries `normals` (SurfaceNormalOutlierFilter compares them with the map's)
    if (result) memset(result, 0, sizeof(*result));
    const int32_t rc = register_checks(ctx, sc.feat[sc.cur], ctx->cfg.dim + 1, sc.n, T_out);
    if (rc != B200ICP_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    int32_t eb = ensure_icp_buffers(ctx, sc.n);
    if (eb != B200ICP_OK) return eb;
    eb = ensure_reading_normals(ctx, sc.n);
    if (eb != B200ICP_OK) return eb;
    return register_on_device(ctx, sc.feat[sc.cur], ctx->cfg.dim + 1, sc.n, T_init, T_out, result, sc.nrm[sc.cur]);
}

static int32_t check_filter_chain(b200icp_ctx* ctx, const b200icp_filter* chain, int32_t n_filters) {
    if (n_filters < 0 || n_filters > 8 || (n_filters > 0 && !chain)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad filter chain (at most 8 entries)");
    for (int i = 0; i < n_filters; ++i)
        if (chain[i].kind != B200ICP_FILTER_BOUNDING_BOX && chain[i].kind != B200ICP_FILTER_DISTANCE_LIMIT &&
            chain[i].kind != B200ICP_FILTER_RANDOM_SAMPLING)
            return fail(ctx, B200ICP_ERR_INVALID_ARG, "unknown input filter");
    return B200ICP_OK;
}

int32_t b200icp_scan_filter(b200icp_ctx* ctx, const b200icp_filter* chain, int32_t n_filters, int64_t* n_out) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int32_t rc = check_filter_chain(ctx, chain, n_filters);
    if (rc != B200ICP_OK) return rc;
    auto& sc = scan_slot(ctx);
    if (n_out) *n_out = sc.n;
    if (sc.n == 0 || n_filters == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    const int o = sc.cur ^ 1;
    int64_t kept = 0;
    CK(filter_cloud_device(ctx->store, build_index(ctx), scan_cloud(ctx), ctx->cfg.dim, chain, n_filters, sc.feat[o], sc.nrm[o], sc.prob[o],
                           sc.extra_rows > 0 ? sc.extra[o] : nullptr, &kept, ctx->stream));
    sc.cur = o;
    sc.n = kept;
    if (n_out) *n_out = kept;
    return B200ICP_OK;
}

int32_t b200icp_scan_add_prob(b200icp_ctx* ctx, float constant) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    auto& sc = scan_slot(ctx);
    CK(cudaSetDevice(ctx->device));
    if (sc.n > 0) CK(launch_fill(sc.prob[sc.cur], constant, sc.n, ctx->stream));
    sc.has_prob = true;
    return B200ICP_OK;
}

// k-NN of SurfaceNormal's neighbour search: `n` queries that are points of the indexed cloud itself (all of them, cell-sorted: the
// full pass; or a gathered subset: the incremental pass); row i of d_out_ids / d_out_d2 (sized by ensure_query_buffers(n, knn)
// beforehand) = query i.  The shell walk of knn.cu with a speculative bound of one cell edge (cells are sized for ~4 per point, so
// the k nearest of a surface point lie well inside it): far fewer insertions into the sorted lists; the few queries that find
// fewer than k neighbours there are rerun without the bound.  nn_variant bit 20 (0x100000), full pass only: TMA-staged tiles
// (selfknn.cu) instead -- measured 2.3x SLOWER on a 2 M-point surface map (profiles/r2_tma_ab.md), kept as the tracked A/B of
// north_star's "TMA-staged point tiles".  nn_variant bit 21 (0x200000): no speculative bound.
static int32_t self_knn(b200icp_ctx* ctx, const GridView& view, const float4* d_queries, int64_t n, int knn, bool queries_are_all_points) {
    cudaStream_t s = ctx->stream;
    int* h_nq = reinterpret_cast<int*>(ctx->h_pinned + 2 * kStateBytes);
    unsigned* h_cnt = reinterpret_cast<unsigned*>(ctx->h_pinned + 2 * kStateBytes + 64);
    ctx->last_selfknn_redone = -1;
    const bool staged = queries_are_all_points && knn <= 16 && (ctx->cfg.nn_variant & 0x100000) && n >= 4096;
    const bool spec = !staged && knn > 1 && !(ctx->cfg.nn_variant & 0x200000) && n >= 4096;
    bool redo_all = !staged && !spec;
    if (staged || spec) {
        const int64_t cap = std::max<int64_t>(n / 4, 4096);
        if (cap > ctx->cap_fb) {
            B200_CUDA_FREE(ctx->d_fb_list);
            B200_CUDA_FREE(ctx->d_fb_q);
            ctx->d_fb_list = nullptr;
            ctx->d_fb_q = nullptr;
            ctx->cap_fb = 0;
            const int64_t c2 = grow_capacity(cap);
            CK(B200_CUDA_MALLOC((void**)&ctx->d_fb_list, (size_t)c2 * sizeof(uint32_t)));
            CK(B200_CUDA_MALLOC((void**)&ctx->d_fb_q, (size_t)c2 * sizeof(float4)));
            ctx->cap_fb = c2;
        }
        if (!ctx->d_fb_count) CK(B200_CUDA_MALLOC((void**)&ctx->d_fb_count, 64));
        if (staged) {
            CK(launch_selfknn_tiles(view, knn, ctx->d_out_ids, ctx->d_out_d2, ctx->d_fb_list, ctx->d_fb_count, (unsigned)ctx->cap_fb, s));
        } else {
            *h_nq = (int)n;
            CK(cudaMemcpyAsync(ctx->d_scalar_nq, h_nq, sizeof(int), cudaMemcpyHostToDevice, s));
            CK(cudaMemsetAsync(ctx->d_fb_count, 0, sizeof(unsigned), s));
            KnnSpec ks;
            // the bound: from the density around each query (at most one cell edge); B200ICP_SPEC_BOUND=<cell edges> fixes it instead
            // (development sweep).  Gathered queries of the incremental pass bring the k-th distance they had in .w.
            ks.bound2 = 0.f;
            if (const char* env = getenv("B200ICP_SPEC_BOUND")) ks.bound2 = (float)atof(env) * view.h * (float)atof(env) * view.h;
            ks.per_query = queries_are_all_points ? 0 : 1;
            ks.list = ctx->d_fb_list;
            ks.count = ctx->d_fb_count;
            ks.capacity = (unsigned)ctx->cap_fb;
            CK(launch_knn(view, d_queries, ctx->d_scalar_nq, (int)n, nullptr, knn, INFINITY, ctx->d_out_ids, ctx->d_out_d2, /*want_original_ids=*/0,
                          ctx->cfg.nn_variant & 0xffff, s, ks));
        }
        CK(cudaMemcpyAsync(h_cnt, ctx->d_fb_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const int64_t redo = *h_cnt;
        if (redo > ctx->cap_fb) {
            redo_all = true;  // (a very sparse cloud for its grid: the unbounded shell walk does everything)
        } else if (redo > 0) {
            if (redo * knn > ctx->cap_fb_rows) {
                B200_CUDA_FREE(ctx->d_fb_ids);
                B200_CUDA_FREE(ctx->d_fb_d2);
                ctx->d_fb_ids = nullptr;
                ctx->d_fb_d2 = nullptr;
                ctx->cap_fb_rows = 0;
                const int64_t c2 = grow_capacity(redo * knn);
                CK(B200_CUDA_MALLOC((void**)&ctx->d_fb_ids, (size_t)c2 * sizeof(int32_t)));
                CK(B200_CUDA_MALLOC((void**)&ctx->d_fb_d2, (size_t)c2 * sizeof(float)));
                ctx->cap_fb_rows = c2;
            }
            // (the list holds query indices: positions for the staged tiles, rows of d_queries for the shell walk -- the same thing
            //  in the full pass)
            CK(launch_gather_reading(d_queries, ctx->d_fb_list, ctx->d_fb_q, redo, s));
            CK(launch_knn(view, ctx->d_fb_q, reinterpret_cast<const int*>(ctx->d_fb_count), (int)redo, nullptr, knn, INFINITY, ctx->d_fb_ids, ctx->d_fb_d2,
                          /*want_original_ids=*/0, ctx->cfg.nn_variant & 0xffff, s));
            CK(launch_selfknn_scatter(ctx->d_fb_list, ctx->d_fb_count, (unsigned)redo, knn, ctx->d_fb_ids, ctx->d_fb_d2, ctx->d_out_ids, ctx->d_out_d2, s));
        }
        if (!redo_all) ctx->last_selfknn_redone = redo;
    }
    if (redo_all) {
        *h_nq = (int)n;
        CK(cudaMemcpyAsync(ctx->d_scalar_nq, h_nq, sizeof(int), cudaMemcpyHostToDevice, s));
        CK(launch_knn(view, d_queries, ctx->d_scalar_nq, (int)n, nullptr, knn, INFINITY, ctx->d_out_ids, ctx->d_out_d2, /*want_original_ids=*/0,
                      ctx->cfg.nn_variant & 0xffff, s));
    }
    return B200ICP_OK;
}

// SurfaceNormalDataPointsFilter{knn} on a device cloud: grid over the cloud, self k-NN, covariance, smallest eigenvector -> d_nrm (cloud order)
static int32_t cloud_normals_dev(b200icp_ctx* ctx, const float* d_in, int rows, int64_t n, int knn, float* d_nrm) {
    const int dim = ctx->cfg.dim;
    cudaStream_t s = ctx->stream;
    CK(grid_build(ctx->aux, d_in, rows, dim, nullptr, n, /*centre=*/false, 0.f, s));
    const int32_t eb = ensure_query_buffers(ctx, n, knn);
    if (eb != B200ICP_OK) return eb;
    const int32_t sk = self_knn(ctx, ctx->aux.view, ctx->aux.pts, n, knn, true);
    if (sk != B200ICP_OK) return sk;
    // d_q4 (n float4, sized by ensure_query_buffers) receives the cell-sorted copy nobody needs; d_nrm the cloud-order normals
    CK(launch_normals(ctx->aux.view, dim, knn, ctx->d_out_ids, ctx->d_out_d2, nullptr, nullptr, 0, ctx->d_q4, d_nrm, nullptr, s));
    return B200ICP_OK;
}

int32_t b200icp_scan_surface_normals(b200icp_ctx* ctx, int32_t knn) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (knn < 1 || knn > 32) return fail(ctx, B200ICP_ERR_INVALID_ARG, "knn must be in [1, 32]");
    auto& sc = scan_slot(ctx);
    CK(cudaSetDevice(ctx->device));
    if (sc.n > 0) {
        const int32_t rc = cloud_normals_dev(ctx, sc.feat[sc.cur], ctx->cfg.dim + 1, sc.n, knn, sc.nrm[sc.cur]);
        if (rc != B200ICP_OK) return rc;
        CK(cudaStreamSynchronize(ctx->stream));  // (h_nq is reused by the next call)
    }
    sc.has_nrm = true;
    return B200ICP_OK;
}

int32_t b200icp_scan_select_extra(b200icp_ctx* ctx, const int32_t* rows, int32_t n_rows) {
    if (!ctx || n_rows < 0 || n_rows > B200ICP_MAX_EXTRA_ROWS || (n_rows > 0 && !rows)) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    auto& sc = scan_slot(ctx);
    for (int i = 0; i < n_rows; ++i)
        if (rows[i] < 0 || rows[i] >= sc.extra_rows) return fail(ctx, B200ICP_ERR_INVALID_ARG, "descriptor row outside `extra`");
    CK(cudaSetDevice(ctx->device));
    if (n_rows > 0 && sc.n > 0) {
        // (the other buffer set may hold a different row count: both were sized for the larger, original one)
        CK(launch_select_rows(sc.extra[sc.cur], sc.extra_rows, sc.n, sc.extra[sc.cur ^ 1], n_rows, rows, ctx->stream));
        std::swap(sc.extra[0], sc.extra[1]);
    }
    // rotating descriptors keep rotating if all their rows survive contiguously
    int nr = 0, rr[4];
    for (int i = 0; i < sc.n_rot; ++i)
        for (int j = 0; j + ctx->cfg.dim <= n_rows; ++j) {
            bool ok = true;
            for (int c = 0; c < ctx->cfg.dim; ++c) ok = ok && rows[j + c] == sc.rot_row[i] + c;
            if (ok) {
                rr[nr++] = j;
                break;
            }
        }
    sc.n_rot = nr;
    for (int i = 0; i < nr; ++i) sc.rot_row[i] = rr[i];
    sc.extra_rows = n_rows;
    return B200ICP_OK;
}

int32_t b200icp_scan_insert_point_distance(b200icp_ctx* ctx, float min_dist_new_point, int64_t* n_added) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (n_added) *n_added = 0;
    if (scan_slot(ctx).n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    if (ctx->store.n_active > 0 && ctx->index_stale) {
        const int32_t rc = commit_index(ctx);
        if (rc != B200ICP_OK) return rc;
    }
    return insert_point_distance_dev(ctx, scan_cloud(ctx), min_dist_new_point, n_added, nullptr);
}

int32_t b200icp_scan_download(b200icp_ctx* ctx, float* features, int64_t capacity, int64_t* n_out) {
    if (!ctx || !n_out) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    auto& sc = scan_slot(ctx);
    *n_out = sc.n;
    if (!features) return B200ICP_OK;
    if (capacity < sc.n) return fail(ctx, B200ICP_ERR_INVALID_ARG, "capacity too small");
    CK(cudaSetDevice(ctx->device));
    if (sc.n > 0)
        CK(cudaMemcpyAsync(features, sc.feat[sc.cur], (size_t)sc.n * (ctx->cfg.dim + 1) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

int32_t b200icp_scan_download_descriptors(b200icp_ctx* ctx, float* normals, float* prob, float* extra, int64_t capacity) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    auto& sc = scan_slot(ctx);
    if (capacity < sc.n) return fail(ctx, B200ICP_ERR_INVALID_ARG, "capacity too small");
    if ((normals && !sc.has_nrm) || (prob && !sc.has_prob) || (extra && sc.extra_rows == 0))
        return fail(ctx, B200ICP_ERR_INVALID_FIELD, "the scan does not carry that descriptor");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (sc.n > 0) {
        if (normals) CK(cudaMemcpyAsync(normals, sc.nrm[sc.cur], (size_t)sc.n * ctx->cfg.dim * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (prob) CK(cudaMemcpyAsync(prob, sc.prob[sc.cur], (size_t)sc.n * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (extra) CK(cudaMemcpyAsync(extra, sc.extra[sc.cur], (size_t)sc.n * sc.extra_rows * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int32_t b200icp_map_surface_normals(b200icp_ctx* ctx, int32_t knn) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (knn < 1 || knn > 32) return fail(ctx, B200ICP_ERR_INVALID_ARG, "knn must be in [1, 32]");
    CK(cudaSetDevice(ctx->device));
    MapStore& st = ctx->store;
    if (st.n_active == 0) return B200ICP_OK;
    if (ctx->index_stale || !(ctx->updating ? ctx->work_valid : ctx->has_map)) {
        const int32_t rc = commit_index(ctx);
        if (rc != B200ICP_OK) return rc;
    }
    cudaStream_t s = ctx->stream;
    GridIndex& idx = synced_index(ctx);  // (the commit above built it)
    const int64_t n = idx.view.n;
    if (n > idx.cap_normals) {
        // (the index was built without normals: nothing to preserve)
        B200_CUDA_FREE(idx.normals);
        idx.normals = nullptr;
        idx.cap_normals = 0;
        CK(B200_CUDA_MALLOC((void**)&idx.normals, (size_t)grow_capacity(n) * sizeof(float4)));
        idx.cap_normals = grow_capacity(n);
    }
    // k-th neighbour distance per store point (incremental bookkeeping); grows with the store, content preserved
    if (st.n > ctx->cap_kth) {
        const int64_t cap = grow_capacity(st.n);
        float* nk = nullptr;
        CK(B200_CUDA_MALLOC((void**)&nk, (size_t)cap * sizeof(float)));
        if (ctx->d_kth && st.nrm_epoch_ok && st.nrm_epoch_n > 0)
            CK(cudaMemcpyAsync(nk, ctx->d_kth, (size_t)std::min<int64_t>(st.nrm_epoch_n, ctx->cap_kth) * sizeof(float), cudaMemcpyDeviceToDevice, s));
        CK(cudaStreamSynchronize(s));
        B200_CUDA_FREE(ctx->d_kth);
        ctx->d_kth = nk;
        ctx->cap_kth = cap;
    }

    // ---- incremental pass: only appends since the last pass -> recompute the new points and the old points that have a new
    //      point within their k-th neighbour distance; every other point keeps the neighbours, hence the normal, it had ----
    const int64_t n_new = st.n - st.nrm_epoch_n;
    bool incremental = st.nrm_epoch_ok && st.nrm_epoch_k == knn && n_new >= 0 && st.nrm_epoch_n > 0 && n_new * 4 <= st.n &&
                       !getenv("B200ICP_FULL_NORMALS");
    ctx->last_normals_recomputed = n;
    if (incremental && (n_new > 0 || st.nrm_touched) && st.n + n > ctx->cap_dirty) {
        B200_CUDA_FREE(ctx->d_dirty);
        B200_CUDA_FREE(ctx->d_list);
        ctx->d_dirty = nullptr;
        ctx->d_list = nullptr;
        ctx->cap_dirty = 0;
        const int64_t cap = grow_capacity(st.n + n);
        CK(B200_CUDA_MALLOC((void**)&ctx->d_dirty, (size_t)cap));
        CK(B200_CUDA_MALLOC((void**)&ctx->d_list, (size_t)cap * sizeof(uint32_t)));
        ctx->cap_dirty = cap;
    }
    unsigned int* d_count = reinterpret_cast<unsigned int*>(ctx->d_scalar_nq) + 4;
    int64_t n_changed = n_new;
    if (incremental && st.nrm_touched) {
        // the window moved since the last pass: the changed set = appended points + points whose loaded flag flipped
        CK(launch_normals_changed(st, st.nrm_epoch_n, ctx->d_dirty, s));
        size_t need = 0;
        thrust::counting_iterator<uint32_t> counting(0u);
        cub::DeviceSelect::Flagged(nullptr, need, counting, ctx->d_dirty, ctx->d_list, d_count, (int)st.n);
        if (need > idx.cub_tmp_bytes) {
            B200_CUDA_FREE(idx.cub_tmp);
            idx.cub_tmp = nullptr;
            idx.cub_tmp_bytes = 0;
            CK(B200_CUDA_MALLOC(&idx.cub_tmp, need + 256));
            idx.cub_tmp_bytes = need + 256;
        }
        size_t bytes = idx.cub_tmp_bytes;
        CK(cub::DeviceSelect::Flagged(idx.cub_tmp, bytes, counting, ctx->d_dirty, ctx->d_list, d_count, (int)st.n, s));
        unsigned int c = 0;
        CK(cudaMemcpyAsync(&c, d_count, sizeof(c), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        n_changed = c;
        if (n_changed * 4 > st.n) incremental = false;  // too much changed: the full pass is cheaper
    }
    if (incremental && n_changed == 0) {
        ctx->last_normals_recomputed = 0;
    } else if (incremental) {
        // grid over the changed points alone (map frame, not centred)
        if (st.nrm_touched) {
            // (d_list holds the subset for the duration of the build only; it is reused for the index positions below)
            CK(grid_build(ctx->aux, reinterpret_cast<const float*>(st.feat), 4, ctx->cfg.dim, nullptr, n_changed, /*centre=*/false, 0.f, s,
                          ctx->d_list));
        } else {
            CK(grid_build(ctx->aux, reinterpret_cast<const float*>(st.feat + st.nrm_epoch_n), 4, ctx->cfg.dim, nullptr, n_new, /*centre=*/false, 0.f, s));
        }
        uint8_t* d_dirty = ctx->d_dirty;         // per store index
        uint8_t* d_flag = ctx->d_dirty + st.n;   // per cell-sorted position of the live index
        CK(launch_normals_dirty(ctx->aux.view, st, ctx->d_kth, st.nrm_epoch_n, d_dirty, s));
        CK(launch_normals_positions(idx.view, d_dirty, d_flag, s));
        size_t need = 0;
        thrust::counting_iterator<uint32_t> counting(0u);
        cub::DeviceSelect::Flagged(nullptr, need, counting, d_flag, ctx->d_list, d_count, (int)n);
        if (need > idx.cub_tmp_bytes) {
            B200_CUDA_FREE(idx.cub_tmp);
            idx.cub_tmp = nullptr;
            idx.cub_tmp_bytes = 0;
            CK(B200_CUDA_MALLOC(&idx.cub_tmp, need + 256));
            idx.cub_tmp_bytes = need + 256;
        }
        size_t bytes = idx.cub_tmp_bytes;
        CK(cub::DeviceSelect::Flagged(idx.cub_tmp, bytes, counting, d_flag, ctx->d_list, d_count, (int)n, s));
        unsigned int m = 0;
        CK(cudaMemcpyAsync(&m, d_count, sizeof(m), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ctx->last_normals_recomputed = m;
        if (m > 0) {
            const int32_t eb = ensure_query_buffers(ctx, m, knn);
            if (eb != B200ICP_OK) return eb;
            CK(launch_normals_gather(idx.view, ctx->d_list, d_count, m, ctx->d_q4, s, ctx->d_kth, st.nrm_epoch_ok ? st.nrm_epoch_n : 0));
            const int32_t sk = self_knn(ctx, idx.view, ctx->d_q4, m, knn, false);
            if (sk != B200ICP_OK) return sk;
            CK(launch_normals(idx.view, ctx->cfg.dim, knn, ctx->d_out_ids, ctx->d_out_d2, ctx->d_list, d_count, m, idx.normals, st.nrm,
                              ctx->d_kth, s));
        }
    } else {
        const int32_t eb = ensure_query_buffers(ctx, n, knn);
        if (eb != B200ICP_OK) return eb;
        // self k-NN: the queries are the cell-sorted map points themselves (every point of a tile shares its candidates)
        const int32_t sk = self_knn(ctx, idx.view, idx.pts, n, knn, true);
        if (sk != B200ICP_OK) return sk;
        CK(launch_normals(idx.view, ctx->cfg.dim, knn, ctx->d_out_ids, ctx->d_out_d2, nullptr, nullptr, 0, idx.normals, st.nrm, ctx->d_kth, s));
    }
    CK(store_clear_touched(st, s));
    CK(cudaStreamSynchronize(s));
    idx.has_normals = true;
    st.has_normals = true;
    // bookkeeping valid when every store point is in the index (parked points keep what they had and are recomputed
    // wholesale when the window moves: store_window invalidates)
    st.nrm_epoch_ok = true;
    st.nrm_epoch_n = st.n;
    st.nrm_epoch_k = knn;
    return B200ICP_OK;
}

/* SurfaceNormalDataPointsFilter{knn} on an arbitrary host cloud (the `input:` chain of a configuration that wants normals on
 * the reading, e.g. for SurfaceNormalOutlierFilter): grid over the cloud, self k-NN, covariance, smallest eigenvector. */
int32_t b200icp_cloud_surface_normals(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, int32_t knn, float* normals_out) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n < 0 || (n > 0 && (!features || !normals_out))) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad cloud");
    if (knn < 1 || knn > 32) return fail(ctx, B200ICP_ERR_INVALID_ARG, "knn must be in [1, 32]");
    if (n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t fb = (((size_t)n * feature_rows * sizeof(float)) + 255) / 256 * 256, nb = (size_t)n * dim * sizeof(float);
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, fb + nb + 256));
    float* d_in = ctx->d_stage_a;
    float* d_nrm = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->d_stage_a) + fb);
    CK(cudaMemcpyAsync(d_in, features, (size_t)n * feature_rows * sizeof(float), cudaMemcpyHostToDevice, s));
    const int32_t rc = cloud_normals_dev(ctx, d_in, feature_rows, n, knn, d_nrm);
    if (rc != B200ICP_OK) return rc;
    CK(cudaMemcpyAsync(normals_out, d_nrm, nb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

/* development aid / tests (not in the public header): points whose normal the last b200icp_map_surface_normals recomputed */
int64_t b200icp_debug_normals_recomputed(const b200icp_ctx* ctx) { return ctx ? ctx->last_normals_recomputed : -1; }
/* ... and how many queries of the last staged self k-NN were redone by the shell walk (-1: the staged kernel was not used) */
int64_t b200icp_debug_selfknn_redone(const b200icp_ctx* ctx) { return ctx ? ctx->last_selfknn_redone : -1; }

int32_t b200icp_map_window(b200icp_ctx* ctx, int32_t load, const int32_t* slab6, int64_t* n_changed) {
    if (!ctx || !slab6) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    int32_t slab[6];
    memcpy(slab, slab6, sizeof(slab));
    if (ctx->cfg.dim == 2) slab[4] = slab[5] = 0;  // Map.cpp:73-77,142-146
    int64_t changed = 0;
    CK(store_window(ctx->store, load, slab, &changed, ctx->stream));
    if (changed > 0) {
        ctx->store.all_loaded = false;
        ctx->store.n_active += load ? changed : -changed;
        ctx->index_stale = true;
    }
    if (n_changed) *n_changed = changed;
    return B200ICP_OK;
}

int32_t b200icp_map_download(b200icp_ctx* ctx, int32_t global, float* features, float* normals, int64_t capacity, int64_t* n_out) {
    if (!ctx || !n_out) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    MapStore& st = ctx->store;
    const int dim = ctx->cfg.dim, rows = dim + 1;
    const int64_t want = global ? st.n : st.n_active;
    *n_out = want;
    if (!features || want == 0) return B200ICP_OK;
    if (capacity < want) return fail(ctx, B200ICP_ERR_INVALID_ARG, "capacity too small");
    std::vector<float4> f((size_t)st.n);
    std::vector<uint8_t> l((size_t)st.n);
    std::vector<float> nr;
    // (copies on the context's own non-blocking stream: the legacy default stream does not order against it)
    CK(cudaMemcpyAsync(f.data(), st.feat, f.size() * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(l.data(), st.loaded, l.size(), cudaMemcpyDeviceToHost, ctx->stream));
    const bool want_n = normals && st.has_normals;
    if (want_n) {
        nr.resize((size_t)st.n * dim);
        CK(cudaMemcpyAsync(nr.data(), st.nrm, nr.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t o = 0;
    for (int64_t i = 0; i < st.n; ++i) {
        if (!global && !l[i]) continue;
        features[o * rows + 0] = f[i].x;
        features[o * rows + 1] = f[i].y;
        if (dim == 3) features[o * rows + 2] = f[i].z;
        features[o * rows + dim] = 1.f;
        if (want_n)
            for (int c = 0; c < dim; ++c) normals[o * dim + c] = nr[i * dim + c];
        ++o;
    }
    *n_out = o;
    return B200ICP_OK;
}

int32_t b200icp_filter_cloud(b200icp_ctx* ctx, float* features, int32_t feature_rows, int64_t* n, const b200icp_filter* chain,
                             int32_t n_filters) {
    if (!ctx || !n) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || *n < 0 || (*n > 0 && !features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad cloud");
    const int32_t rc = check_filter_chain(ctx, chain, n_filters);
    if (rc != B200ICP_OK) return rc;
    if (*n == 0 || n_filters == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t fb = (((size_t)*n * feature_rows * sizeof(float)) + 255) / 256 * 256;
    CK(grow(ctx->d_stage_a, ctx->stage_a_bytes, 2 * fb + 256));
    float* d_in = ctx->d_stage_a;
    float* d_out = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->d_stage_a) + fb);
    CK(cudaMemcpyAsync(d_in, features, (size_t)*n * feature_rows * sizeof(float), cudaMemcpyHostToDevice, s));
    int64_t kept = 0;
    DevCloud in;
    in.feat = d_in;
    in.rows = feature_rows;
    in.n = *n;
    CK(filter_cloud_device(ctx->store, build_index(ctx), in, dim, chain, n_filters, d_out, nullptr, nullptr, nullptr, &kept, s));
    if (kept > 0) CK(cudaMemcpyAsync(features, d_out, (size_t)kept * feature_rows * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *n = kept;
    return B200ICP_OK;
}

int32_t b200icp_map_set_prob(b200icp_ctx* ctx, const float* prob, float constant) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    MapStore& st = ctx->store;
    if (st.n == 0) return B200ICP_OK;
    if (prob) {
        CK(cudaMemcpyAsync(st.prob, prob, (size_t)st.n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        CK(launch_fill(st.prob, constant, st.n, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    st.has_prob = true;
    return B200ICP_OK;
}

// host copy of a per-point store array, compacted to the loaded points unless `global`
static int32_t download_rows(b200icp_ctx* ctx, const float* d_src, int rows, int32_t global, float* out, int64_t capacity) {
    MapStore& st = ctx->store;
    std::vector<float> p((size_t)st.n * rows);
    std::vector<uint8_t> l((size_t)st.n);
    // (copies on the context's own non-blocking stream: the legacy default stream does not order against it)
    CK(cudaMemcpyAsync(p.data(), d_src, p.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(l.data(), st.loaded, l.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t o = 0;
    for (int64_t i = 0; i < st.n; ++i) {
        if (!global && !l[i]) continue;
        if (o >= capacity) return fail(ctx, B200ICP_ERR_INVALID_ARG, "capacity too small");
        for (int c = 0; c < rows; ++c) out[o * rows + c] = p[i * rows + c];
        ++o;
    }
    return B200ICP_OK;
}

int32_t b200icp_map_download_prob(b200icp_ctx* ctx, int32_t global, float* prob, int64_t capacity) {
    if (!ctx || !prob) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->store.has_prob) return fail(ctx, B200ICP_ERR_INVALID_FIELD, "the map has no probabilityDynamic descriptor");
    return download_rows(ctx, ctx->store.prob, 1, global, prob, capacity);
}

int32_t b200icp_map_has_prob(const b200icp_ctx* ctx) { return (ctx && ctx->store.has_prob) ? 1 : 0; }

/* the map's other descriptors (everything but normals / probabilityDynamic), extra_rows floats per point, insertion order */
int32_t b200icp_map_set_extra(b200icp_ctx* ctx, const float* extra, int32_t extra_rows) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (extra_rows < 0 || extra_rows > B200ICP_MAX_EXTRA_ROWS || (extra_rows > 0 && !extra)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad descriptor block");
    CK(cudaSetDevice(ctx->device));
    MapStore& st = ctx->store;
    CK(store_set_extra_rows(st, extra_rows, ctx->stream));
    if (extra_rows > 0 && st.n > 0) CK(cudaMemcpyAsync(st.extra, extra, (size_t)st.n * extra_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

int32_t b200icp_map_extra_rows(const b200icp_ctx* ctx) { return ctx ? ctx->store.extra_rows : 0; }

int32_t b200icp_map_select_extra(b200icp_ctx* ctx, const int32_t* rows, int32_t n_rows) {
    if (!ctx || n_rows < 0 || (n_rows > 0 && !rows)) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    CK(store_select_extra(ctx->store, rows, n_rows, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ICP_OK;
}

int32_t b200icp_map_download_extra(b200icp_ctx* ctx, int32_t global, float* extra, int64_t capacity) {
    if (!ctx || !extra) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    CK(cudaSetDevice(ctx->device));
    if (ctx->store.extra_rows == 0) return fail(ctx, B200ICP_ERR_INVALID_FIELD, "the map carries no other descriptor");
    return download_rows(ctx, ctx->store.extra, ctx->store.extra_rows, global, extra, capacity);
}

static int32_t append_dev(b200icp_ctx* ctx, const DevCloud& in, int64_t* n_added) {
    CK(store_append_all(ctx->store, in, ctx->cfg.dim, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->index_stale = true;
    if (n_added) *n_added = in.n;
    return B200ICP_OK;
}

int32_t b200icp_map_append(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in, const float* input_normals,
                           const float* input_prob, int64_t* n_added) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n_in < 0 || (n_in > 0 && !input)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad input cloud");
    if (n_added) *n_added = 0;
    if (n_in == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    DevCloud in;
    const int32_t rc = upload_input(ctx, input, feature_rows, n_in, input_normals, input_prob, &in);
    if (rc != B200ICP_OK) return rc;
    return append_dev(ctx, in, n_added);
}

int32_t b200icp_map_replace_local(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                                  const float* prob, const float* extra, int32_t extra_rows) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n < 0 || (n > 0 && !features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad cloud");
    if (extra && (extra_rows < 1 || extra_rows > B200ICP_MAX_EXTRA_ROWS)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "extra_rows out of range");
    CK(cudaSetDevice(ctx->device));
    DevCloud in;
    if (n > 0) {
        const int32_t rc = upload_input(ctx, features, feature_rows, n, normals, prob, &in);
        if (rc != B200ICP_OK) return rc;
        if (extra) {
            CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, (size_t)n * extra_rows * sizeof(float)));
            CK(cudaMemcpyAsync(ctx->d_stage_b, extra, (size_t)n * extra_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            in.extra = ctx->d_stage_b;
            in.extra_rows = extra_rows;
        }
    }
    MapStore& st = ctx->store;
    if (st.n_active == st.n) {  // nothing parked: the cloud's descriptor set becomes the map's
        st.n = 0;
        st.n_active = 0;
    }
    CK(store_replace_loaded(st, build_index(ctx), in, dim, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->index_stale = true;
    return B200ICP_OK;
}

/* ---- spill tier under the device grid (the CellManager seam, CellManager.h:15-18) ---- */
int32_t b200icp_map_evict_parked(b200icp_ctx* ctx, float* features, float* normals, float* prob, float* extra, int64_t capacity, int64_t* n_out) {
    if (!ctx || !n_out) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    MapStore& st = ctx->store;
    *n_out = st.n - st.n_active;
    if (!features || *n_out == 0) return B200ICP_OK;  // (count query)
    if (capacity < *n_out) return fail(ctx, B200ICP_ERR_INVALID_ARG, "capacity too small");
    if ((normals && !st.has_normals) || (prob && !st.has_prob) || (extra && st.extra_rows == 0))
        return fail(ctx, B200ICP_ERR_INVALID_FIELD, "the map does not carry that descriptor");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int dim = ctx->cfg.dim, rows = dim + 1;
    int64_t np = 0;
    CK(store_extract_parked(st, build_index(ctx), dim, &np, s));
    std::vector<float4> f((size_t)np);
    CK(cudaMemcpyAsync(f.data(), st.feat2, f.size() * sizeof(float4), cudaMemcpyDeviceToHost, s));
    if (normals) CK(cudaMemcpyAsync(normals, st.nrm2, (size_t)np * dim * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (prob) CK(cudaMemcpyAsync(prob, st.prob2, (size_t)np * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (extra) CK(cudaMemcpyAsync(extra, st.extra2, (size_t)np * st.extra_rows * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < np; ++i) {
        features[i * rows + 0] = f[i].x;
        features[i * rows + 1] = f[i].y;
        if (dim == 3) features[i * rows + 2] = f[i].z;
        features[i * rows + dim] = 1.f;
    }
    CK(store_remove_parked(st, build_index(ctx), dim, s));
    *n_out = np;
    return B200ICP_OK;
}

int32_t b200icp_map_append_cloud(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                                 const float* prob, const float* extra, int32_t extra_rows, int64_t* n_added) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n < 0 || (n > 0 && !features)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad cloud");
    if (extra && (extra_rows < 1 || extra_rows > B200ICP_MAX_EXTRA_ROWS)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "extra_rows out of range");
    if (n_added) *n_added = 0;
    if (n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    DevCloud in;
    const int32_t rc = upload_input(ctx, features, feature_rows, n, normals, prob, &in);
    if (rc != B200ICP_OK) return rc;
    if (extra) {
        CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, (size_t)n * extra_rows * sizeof(float)));
        CK(cudaMemcpyAsync(ctx->d_stage_b, extra, (size_t)n * extra_rows * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        in.extra = ctx->d_stage_b;
        in.extra_rows = extra_rows;
    }
    return append_dev(ctx, in, n_added);
}

int32_t b200icp_scan_append(b200icp_ctx* ctx, int64_t* n_added) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (n_added) *n_added = 0;
    if (scan_slot(ctx).n == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    return append_dev(ctx, scan_cloud(ctx), n_added);
}

static int32_t octree_checks(b200icp_ctx* ctx, float max_size_by_node, int32_t max_point_by_node, int32_t sampling_method) {
    if (!(max_size_by_node > 0.f)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "maxSizeByNode must be positive");
    if (max_point_by_node != 1) return fail(ctx, B200ICP_ERR_NOT_IMPLEMENTED, "OctreeGrid: only maxPointByNode = 1 (the LPM default) is implemented");
    if (sampling_method < 0 || sampling_method > 3)
        return fail(ctx, B200ICP_ERR_INVALID_ARG, "OctreeGrid: samplingMethod must be 0 (first), 1 (random), 2 (centroid) or 3 (medoid)");
    return B200ICP_OK;
}

static int32_t octree_dev(b200icp_ctx* ctx, const DevCloud* in, float max_size_by_node, int32_t sampling_method, int64_t* n_after) {
    cudaStream_t s = ctx->stream;
    MapStore& st = ctx->store;
    if (in && in->n > 0) CK(store_append_all(st, *in, ctx->cfg.dim, s));  // map.concatenate(input)
    int64_t removed = 0;
    // the random sampler is reproducible: same base seed + same call count -> same survivors (B200ICP_OCTREE_SEED sets the base)
    uint64_t seed = 0x0c7ee5eedull;
    if (const char* env = getenv("B200ICP_OCTREE_SEED")) seed = strtoull(env, nullptr, 0);
    seed += ctx->octree_calls++;
    CK(store_octree_filter(st, build_index(ctx), ctx->cfg.dim, max_size_by_node, sampling_method, seed, &removed, s));
    CK(cudaStreamSynchronize(s));
    ctx->index_stale = true;
    if (n_after) *n_after = st.n_active;
    return B200ICP_OK;
}

int32_t b200icp_map_octree(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in, const float* input_normals,
                           const float* input_prob, float max_size_by_node, int32_t max_point_by_node, int32_t sampling_method,
                           int64_t* n_after) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n_in < 0 || (n_in > 0 && !input)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad input cloud");
    int32_t rc = octree_checks(ctx, max_size_by_node, max_point_by_node, sampling_method);
    if (rc != B200ICP_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    DevCloud in;
    if (n_in > 0) {
        rc = upload_input(ctx, input, feature_rows, n_in, input_normals, input_prob, &in);
        if (rc != B200ICP_OK) return rc;
    }
    return octree_dev(ctx, n_in > 0 ? &in : nullptr, max_size_by_node, sampling_method, n_after);
}

int32_t b200icp_scan_octree(b200icp_ctx* ctx, float max_size_by_node, int32_t max_point_by_node, int32_t sampling_method, int64_t* n_after) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int32_t rc = octree_checks(ctx, max_size_by_node, max_point_by_node, sampling_method);
    if (rc != B200ICP_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const DevCloud in = scan_cloud(ctx);
    return octree_dev(ctx, &in, max_size_by_node, sampling_method, n_after);
}

int32_t b200icp_map_cut_at_threshold(b200icp_ctx* ctx, float threshold, int32_t use_larger_than, int64_t* n_removed) {
    if (!ctx) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    if (!ctx->store.has_prob) return fail(ctx, B200ICP_ERR_INVALID_FIELD, "CutAtDescriptorThreshold: descriptor probabilityDynamic not found");
    CK(cudaSetDevice(ctx->device));
    int64_t removed = 0;
    CK(store_cut_prob(ctx->store, build_index(ctx), ctx->cfg.dim, threshold, use_larger_than, &removed, ctx->stream));
    if (removed > 0) ctx->index_stale = true;
    if (n_removed) *n_removed = removed;
    return B200ICP_OK;
}

static int32_t dynamic_points_checks(b200icp_ctx* ctx, bool input_has_prob) {
    MapStore& st = ctx->store;
    if (!input_has_prob)
        return fail(ctx, B200ICP_ERR_INVALID_FIELD, "Missing field 'probabilityDynamic' in input point cloud. You can add it with the AddDescriptorDataPointsFilter in your input filters.");
    if (!st.has_normals)
        return fail(ctx, B200ICP_ERR_INVALID_FIELD, "Missing field 'normals' in map point cloud. You can add it with the SurfaceNormalDataPointsFilter in your post filters.");
    if (!st.has_prob) return fail(ctx, B200ICP_ERR_INVALID_FIELD, "Missing field 'probabilityDynamic' in map point cloud.");
    return B200ICP_OK;
}

// `in`: the scan in the map frame, on the device
static int32_t dynamic_points_dev(b200icp_ctx* ctx, const DevCloud& in, const float* pose, const b200icp_dynamic_params* prm) {
    const int dim = ctx->cfg.dim;
    MapStore& st = ctx->store;
    const int64_t n_in = in.n;
    cudaStream_t s = ctx->stream;
    // pose.inverse(): rigid inverse, computed in double
    float P[16], Tinv[16];
    embed(pose, dim, P);
    mat4_identity(Tinv);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Tinv[c * 4 + r] = P[r * 4 + c];
    for (int r = 0; r < 3; ++r) {
        double acc = 0.0;
        for (int c = 0; c < 3; ++c) acc += (double)P[r * 4 + c] * (double)P[12 + c];
        Tinv[12 + r] = (float)(-acc);
    }
    if (!st.all_loaded || st.n_active != st.n) CK(store_compact_active(st, build_index(ctx), s));
    // scratch: input in the sensor frame (float4) + its angles (2 floats)
    CK(grow(ctx->d_stage_b, ctx->stage_b_bytes, (size_t)n_in * (sizeof(float4) + 2 * sizeof(float)) + 512));
    float4* d_in_sensor = reinterpret_cast<float4*>(ctx->d_stage_b);
    float* d_angles = reinterpret_cast<float*>(reinterpret_cast<char*>(ctx->d_stage_b) + (((size_t)n_in * sizeof(float4)) + 255) / 256 * 256);
    CK(launch_dyn_input(in.feat, in.rows, dim, Tinv, n_in, d_in_sensor, d_angles, s));
    // Nabo::NNS::create(inputInSensorFrameAngles) + knn(map angles, 1, 0, ALLOW_SELF_MATCH, 2 * beamHalfAngle) -- :75-78
    CK(grid_build(ctx->aux, d_angles, 2, 2, nullptr, n_in, /*centre=*/false, 0.f, s));
    const int64_t na = st.n_active;
    const int32_t eb = ensure_query_buffers(ctx, na, 1);
    if (eb != B200ICP_OK) return eb;
    CK(launch_dyn_queries(st, dim, Tinv, prm->sensor_max_range, ctx->d_q4, s));
    int* h_nq = reinterpret_cast<int*>(ctx->h_pinned + 2 * kStateBytes);
    *h_nq = (int)na;
    CK(cudaMemcpyAsync(ctx->d_scalar_nq, h_nq, sizeof(int), cudaMemcpyHostToDevice, s));
    const float r = 2.f * prm->beam_half_angle;
    CK(launch_knn(ctx->aux.view, ctx->d_q4, ctx->d_scalar_nq, (int)na, nullptr, 1, r * r, ctx->d_out_ids, ctx->d_out_d2,
                  /*want_original_ids=*/1, ctx->cfg.nn_variant, s));
    DynParams dp{prm->threshold_dynamic, prm->alpha, prm->beta, prm->beam_half_angle, prm->epsilon_a, prm->epsilon_d, prm->sensor_max_range};
    CK(launch_dyn_update(st, dim, Tinv, dp, d_in_sensor, ctx->d_out_ids, ctx->d_out_d2, s));
    CK(cudaStreamSynchronize(s));
    return B200ICP_OK;
}

int32_t b200icp_map_dynamic_points(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in, const float* input_prob,
                                   const float* pose, const b200icp_dynamic_params* prm) {
    if (!ctx || !prm || !pose) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int dim = ctx->cfg.dim;
    if (feature_rows != dim + 1 || n_in < 0 || (n_in > 0 && !input)) return fail(ctx, B200ICP_ERR_INVALID_ARG, "bad input cloud");
    int32_t rc = dynamic_points_checks(ctx, input_prob != nullptr);
    if (rc != B200ICP_OK) return rc;
    if (n_in == 0 || ctx->store.n_active == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    DevCloud in;
    rc = upload_input(ctx, input, feature_rows, n_in, nullptr, nullptr, &in);
    if (rc != B200ICP_OK) return rc;
    return dynamic_points_dev(ctx, in, pose, prm);
}

int32_t b200icp_scan_dynamic_points(b200icp_ctx* ctx, const float* pose, const b200icp_dynamic_params* prm) {
    if (!ctx || !prm || !pose) return B200ICP_ERR_INVALID_ARG;
    B200_LOCK(ctx);
    const int32_t rc = dynamic_points_checks(ctx, scan_slot(ctx).has_prob);
    if (rc != B200ICP_OK) return rc;
    if (scan_slot(ctx).n == 0 || ctx->store.n_active == 0) return B200ICP_OK;
    CK(cudaSetDevice(ctx->device));
    return dynamic_points_dev(ctx, scan_cloud(ctx), pose, prm);
}

}  // extern "C"
