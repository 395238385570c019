// common.cuh -- types shared by the translation units of libb200icp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <chrono>

#include "b200icp.h"

namespace b200 {

constexpr int kSMs = 148;        // B200
// Capacity policy of every growing device buffer: double, never below 1 Mi elements.  cudaMalloc /
// cudaFree of multi-hundred-MB buffers cost 100+ ms and synchronise the device, so a growing online
// map must hit them O(log N) times, not every few scans.
// B200ICP_TRACE_ALLOC=1: every device allocation / release of the library is reported on stderr with its size, call site
// and host time (development aid: where an online map update stalls).
inline bool trace_alloc_enabled() {
    static const bool on = [] { const char* e = getenv("B200ICP_TRACE_ALLOC"); return e && e[0] == '1'; }();
    return on;
}
inline cudaError_t traced_malloc(void** p, size_t bytes, const char* file, int line) {
    if (!trace_alloc_enabled()) return cudaMalloc(p, bytes);
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaMalloc(p, bytes);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[b200icp alloc] cudaMalloc %10.2f MB %8.3f ms  %s:%d\n", bytes / 1048576.0, ms, file, line);
    return e;
}
inline cudaError_t traced_free(void* p, const char* file, int line) {
    if (!trace_alloc_enabled() || !p) return cudaFree(p);
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = cudaFree(p);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[b200icp alloc] cudaFree                 %8.3f ms  %s:%d\n", ms, file, line);
    return e;
}
// every device allocation / release of the library goes through these two (call site recorded for the trace)
#define B200_CUDA_MALLOC(p, bytes) ::b200::traced_malloc((void**)(p), (bytes), __FILE__, __LINE__)
#define B200_CUDA_FREE(p) ::b200::traced_free((void*)(p), __FILE__, __LINE__)

inline int64_t grow_capacity(int64_t n) { return n * 2 > (int64_t(1) << 20) ? n * 2 : (int64_t(1) << 20); }
constexpr int kMaxCells = 1 << 25;  // dense cell table cap (uint32 per cell -> 128 MB)
constexpr int kHistBins = 2048;  // radix-select: 11 + 11 + 10 bits
constexpr int kHistWords = 36864; // uint32 words of the histogram / candidate-list buffer (layout in loop.cu)
constexpr int kAccSlots = 32;    // doubles per block partial (29 used by point-to-plane)
constexpr int kAccBlocks = 128;      // accumulate-kernel grid (one partial each, summed in fixed order)
constexpr int kLoopMaxBlocks = 192;  // persistent loop kernel: one CTA per SM
constexpr int kMaxAccBlocks = kLoopMaxBlocks;  // partials buffer size

// Uniform grid over a cloud; points sorted by linear cell id (x fastest), so the cells
// [x0..x1] of one (y, z) row are one contiguous run of `pts`.
struct GridView {
    const float4* __restrict__ pts;          // (x, y, z, bit-cast original index), cell-sorted
    const uint32_t* __restrict__ cell_start; // n_cells + 1 exclusive prefix
    float ox, oy, oz;                        // grid origin (min corner)
    float h, inv_h;                          // cell edge
    float slack;                             // safety margin in grid units for boundary distances
    int nx, ny, nz;
    int n;                                   // points
};

// Device-resident state of one registration (one `icp(input)` call). 4x4 column-major always;
// the 2-D case is embedded (z row/column = identity) so one set of kernels serves both.
struct IcpState {
    float T[16];  // T_iter (refMean frame)
    int nq;       // reading points (device-side so captured graphs do not depend on it)
    int iter;     // iterations completed
    int done;     // loop finished (checker said stop, or error)
    int status;   // b200icp_status
    int max_iter_reached;
    float overlap, used_ratio;
    long long pairs;
    float limit;       // last quantile limit (diagnostic)
    unsigned int ticket;  // last-block election in the accumulate kernel
    int counter;          // CounterTransformationChecker
    int dcount;           // DifferentialTransformationChecker ring fill
    float dq[8][4];
    float dt[8][3];
    float bq0[4];
    float bt0[3];
    // instrumentation of the persistent loop kernel (CTA 0's view, %globaltimer ns)
    unsigned long long loop_search_ns;  // sum over iterations of the correspondence-search phase
    unsigned long long loop_total_ns;   // first iteration start -> last iteration end
    int loop_iters_timed;
    int fast_iters;  // iterations that took the one-barrier path
    // one-barrier iteration (loop.cu): window [win_lo, win_hi] of dist2 bit patterns predicted to hold the next quantile limit
    uint32_t win_lo, win_hi;
    int win_valid, have_limit;
    int searched_queries;  // queries that went through the search phase (the rest were verified against their bound)
    int hist_iters;        // iterations that took the two-barrier path (window from a level-0 histogram)
    float dyn_quantile;    // VarTrimmedDist: the ratio tuned for this iteration (outlier.cu)
    float robust_scale;    // RobustOutlierFilter: the current scale estimate (outlier.cu); kept across iterations
    unsigned int pmax2_bits;  // float bits of max |p|^2 over the finite reading points (refMean frame; prep_reading_kernel): scales the
                              // loop kernel's fixed-point error sums
    int sum_overflow;         // loop kernel: a partial sum left the fixed-point range (the registration then fails loudly)
};

static_assert(sizeof(IcpState) <= 512, "IcpState must fit its 512-byte slot");

// Radix-select bookkeeping; lives kSelectOffset bytes after IcpState in the same allocation.
struct SelectState {
    uint32_t bin[3];
    uint32_t rank[3];
    unsigned int ticket[3];
    uint32_t pad;
};
constexpr int kSelectOffset = 512;
constexpr int kDebugOffset = 640;  // 32 x uint64 globaltimer stamps of the last iteration (development aid)
#ifdef B200ICP_STAMPS
#define B200_STAMP(st, i)                                                                         \
    do {                                                                                          \
        unsigned long long t_;                                                                    \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                    \
        reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(st) + b200::kDebugOffset)[i] = t_; \
    } while (0)
#else
#define B200_STAMP(st, i) do { } while (0)
#endif
constexpr int kStateBytes = 1024;

struct IcpParams {  // by-value kernel argument, constant for the life of a context
    int dim;
    int knn;
    float max_r2;  // maxDist^2 (inf allowed)
    int n_outlier;
    int outlier_kind[B200ICP_MAX_OUTLIER_FILTERS];
    float outlier_param[B200ICP_MAX_OUTLIER_FILTERS];
    float outlier_param2[B200ICP_MAX_OUTLIER_FILTERS];
    float outlier_param3[B200ICP_MAX_OUTLIER_FILTERS];
    int outlier_mode[B200ICP_MAX_OUTLIER_FILTERS];  // RobustOutlierFilter: B200ICP_ROBUST_MODE bits
    const float4* rnrm;   // reading normals (refMean frame, same order as the reading) for SurfaceNormalOutlierFilter, or null:
                          // set per registration, the only field that is not constant for the life of a context
    int quantile_filter;  // index of the Trimmed/Median filter or -1
    float quantile;       // ratio (Trimmed), 0.5 (Median), or < 0: read IcpState::dyn_quantile (VarTrimmed)
    int minimizer;
    int max_iteration_count;
    int use_differential;
    float min_diff_rot_err, min_diff_trans_err;
    int smooth_length;
    int use_bound;
    float max_rotation_norm, max_translation_norm;
    int counter_after;  // b200icp_config::checker_order
    int min_flags;      // b200icp_config::minimizer_flags: bit 0 force2D, bit 1 force4DOF (PointToPlaneErrorMinimizer, 3-D clouds)
};

// ---- index.cu ------------------------------------------------------------------------------
struct GridIndex {
    GridView view{};
    float4* pts = nullptr;
    float4* normals = nullptr;  // cell-sorted (nx, ny, nz, 0) or null
    uint32_t* cell_start = nullptr;
    int64_t cap_pts = 0, cap_normals = 0, cap_cells = 0;
    bool has_normals = false;
    // scratch
    float4* tmp_pts = nullptr;
    uint32_t *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    unsigned long long* d_reduce = nullptr;  // bbox + fixed-point sums
    int64_t cap_scratch = 0;
    float mean[3] = {0, 0, 0};
};

// Build the grid over `n` points given as `rows` floats each (device pointer, first dim floats
// are coordinates).  centre=true subtracts the fixed-point mean first (ICPSequence::setMap).
// cell_hint <= 0 lets the builder choose the cell edge.
// d_subset (optional): n indices into the cloud; the grid is then built over those points only and
// pts[].w / the normals gather keep referring to positions in the full cloud.
cudaError_t grid_build(GridIndex& g, const float* d_feat, int rows, int dim, const float* d_normals,
                       int64_t n, bool centre, float cell_hint, cudaStream_t s, const uint32_t* d_subset = nullptr);
void grid_free(GridIndex& g);
cudaError_t ensure_scratch(GridIndex& g, int64_t n);  // sort scratch for at least n pairs
cudaError_t cloud_bounds(GridIndex& g, const float* d_feat, int rows, int dim, int64_t n, const uint32_t* d_subset, float* lo3,
                         float* hi3, cudaStream_t s);

// Stable radix sort of (key, value) pairs on `s` using g's scratch (also used to cell-sort readings).
cudaError_t sort_pairs(GridIndex& scratch_owner, uint32_t* keys_in, uint32_t* keys_out,
                       uint32_t* vals_in, uint32_t* vals_out, int64_t n, int end_bit, cudaStream_t s);

// ---- mapupd.cu -----------------------------------------------------------------------------
// Device-resident map: every point ever inserted, in insertion order, map frame. `loaded` marks
// membership of Map::localPointCloud; the rest is what the reference parks in its CellManager.
// A cloud on the device, by pointers (PM::DataPoints): features `rows` floats per point, optional descriptors.  `extra` stacks
// every descriptor other than `normals` / `probabilityDynamic` (intensity, t, ring, observationDirections ...): extra_rows
// floats per point, point-major (== the reference's column-major extra_rows x n block); their names live on the host.
struct DevCloud {
    const float* feat = nullptr;
    int rows = 0;
    int64_t n = 0;
    const float* nrm = nullptr;
    const float* prob = nullptr;
    const float* extra = nullptr;
    int extra_rows = 0;
};

struct MapStore {
    float4* feat = nullptr;      // (x, y, z, 1)
    float* nrm = nullptr;        // dim floats per point
    float* prob = nullptr;       // `probabilityDynamic` descriptor (DynamicPointsMapperModule), 1 float per point
    float* extra = nullptr;      // the other descriptors, extra_rows floats per point (DataPoints::concatenate keeps the common ones:
    float* extra2 = nullptr;     //   the host mirror maps labels to rows and keeps both sides' layouts equal)
    int extra_rows = 0;
    int64_t cap_extra = 0;       // floats
    uint8_t* loaded = nullptr;
    uint8_t* touched = nullptr;  // loaded flag flipped since the last SurfaceNormal pass (incremental normals)
    uint32_t* active = nullptr;  // indices of the loaded points (valid after store_compact_active)
    uint32_t *tmp_u32a = nullptr, *tmp_u32b = nullptr;
    unsigned long long* d_counter = nullptr;
    int64_t n = 0, cap = 0, n_active = 0, cap_tmp = 0;
    bool has_normals = false;
    bool has_prob = false;
    bool all_loaded = true;
    // incremental SurfaceNormal: store points [0, nrm_epoch_n) carry normals (+ k-th neighbour distances) computed with
    // knn = nrm_epoch_k and nothing but appends happened since
    bool nrm_epoch_ok = false;
    bool nrm_touched = false;  // some `touched` flags are set
    int64_t nrm_epoch_n = 0;
    int nrm_epoch_k = 0;
    // double buffers for compaction
    float4* feat2 = nullptr;
    float* nrm2 = nullptr;
    float* prob2 = nullptr;
    uint8_t* loaded2 = nullptr;
    unsigned long long* keys64_a = nullptr;
    unsigned long long* keys64_b = nullptr;
    int64_t cap_keys64 = 0;
};
void store_free(MapStore& m);
cudaError_t store_reserve(MapStore& m, int dim, int64_t n, cudaStream_t s);
cudaError_t store_reserve_scratch(MapStore& m, int64_t n);
cudaError_t store_set(MapStore& m, const DevCloud& in, int dim, cudaStream_t s);
// resize / (re)allocate the `extra` block for `rows` floats per point (content dropped when the row count changes)
cudaError_t store_set_extra_rows(MapStore& m, int rows, cudaStream_t s);
// keep only the listed rows of `extra`, in that order (DataPoints::concatenate with a cloud that lacks some descriptors)
cudaError_t store_select_extra(MapStore& m, const int* rows, int n_rows, cudaStream_t s);
cudaError_t store_compact_active(MapStore& m, GridIndex& scratch, cudaStream_t s);
cudaError_t store_window(MapStore& m, int load, const int32_t* slab6, int64_t* changed, cudaStream_t s);
cudaError_t store_insert_point_distance(MapStore& m, GridIndex& scratch, const DevCloud& in, int dim, const int32_t* d_nn_id, float min_dist,
                                        int64_t* n_kept, uint8_t* d_keep_out, cudaStream_t s);
// map.concatenate(input): append every input point (descriptors survive only if both clouds have them)
cudaError_t store_append_all(MapStore& m, const DevCloud& in, int dim, cudaStream_t s);
// OctreeGridDataPointsFilter{maxPointByNode 1, maxSizeByNode, samplingMethod 0 first | 1 random | 2 centroid | 3 medoid} over the loaded points
cudaError_t store_octree_filter(MapStore& m, GridIndex& scratch, int dim, float max_size_by_node, int sampling_method, uint64_t seed,
                                int64_t* n_removed, cudaStream_t s);
// CutAtDescriptorThresholdDataPointsFilter{probabilityDynamic, useLargerThan, threshold} over the loaded points
cudaError_t store_cut_prob(MapStore& m, GridIndex& scratch, int dim, float threshold, int use_larger_than, int64_t* n_removed, cudaStream_t s);
// input filter chain on a device cloud (rows floats per point): keep flags -> ordered compaction
// `in` and the `out_*` buffers must not alias; descriptors of `in` that have an output buffer are compacted with the points
cudaError_t filter_cloud_device(MapStore& tmp, GridIndex& scratch, const DevCloud& in, int dim, const b200icp_filter* chain, int n_filters,
                                float* out_feat, float* out_nrm, float* out_prob, float* out_extra, int64_t* n_out, cudaStream_t s);
cudaError_t store_extract_parked(MapStore& m, GridIndex& scratch, int dim, int64_t* n_parked, cudaStream_t s);
cudaError_t store_remove_parked(MapStore& m, GridIndex& scratch, int dim, cudaStream_t s);
cudaError_t store_replace_loaded(MapStore& m, GridIndex& scratch, const DevCloud& in, int dim, cudaStream_t s);
// out[i * out_rows + c] = in[i * in_rows + rows[c]]
cudaError_t launch_select_rows(const float* d_in, int in_rows, int64_t n, float* d_out, int out_rows, const int* rows, cudaStream_t s);
cudaError_t launch_fill(float* d_out, float value, int64_t n, cudaStream_t s);
struct DynParams {  // DynamicPointsMapperModule parameters (DynamicPointsMapperModule.h:33-44)
    float thresholdDynamic, alpha, beta, beamHalfAngle, epsilonA, epsilonD, sensorMaxRange;
};
// stage 1: input -> sensor frame (x, y, z, |p|) and (elevation, azimuth); stage 2: angles of the loaded
// map points within sensorMaxRange as k-NN queries (NaN otherwise); stage 3: the Bayesian update.
cudaError_t launch_dyn_input(const float* d_in, int rows, int dim, const float* Tinv16, int64_t n_in, float4* d_in_sensor, float* d_in_angles,
                             cudaStream_t s);
cudaError_t launch_dyn_queries(const MapStore& m, int dim, const float* Tinv16, float sensor_max_range, float4* d_q4, cudaStream_t s);
cudaError_t launch_dyn_update(MapStore& m, int dim, const float* Tinv16, const DynParams& prm, const float4* d_in_sensor,
                              const int32_t* d_ids, const float* d_d2, cudaStream_t s);
cudaError_t launch_normals(const GridView& g, int dim, int knn, const int32_t* d_nn_pos, const float* d_nn_d2, const uint32_t* d_list,
                           const unsigned int* d_n_list, long long list_capacity, float4* d_nrm_sorted, float* d_store_nrm, float* d_kth,
                           cudaStream_t s);
cudaError_t launch_normals_dirty(const GridView& new_points, const MapStore& m, const float* d_kth, int64_t n_old, uint8_t* d_dirty, cudaStream_t s);
cudaError_t launch_normals_changed(const MapStore& m, int64_t n_old, uint8_t* d_flag, cudaStream_t s);
cudaError_t store_clear_touched(MapStore& m, cudaStream_t s);
cudaError_t launch_normals_positions(const GridView& g, const uint8_t* d_dirty, uint8_t* d_flag, cudaStream_t s);
cudaError_t launch_normals_gather(const GridView& g, const uint32_t* d_list, const unsigned int* d_n_list, long long capacity, float4* d_q, cudaStream_t s,
                                  const float* d_kth = nullptr, long long n_old = 0);

// ---- knn.cu --------------------------------------------------------------------------------
// queries: float4 (x, y, z, *) in the grid's frame, optionally moved by state->T first.
// out_ids: cell-sorted positions (want_original_ids = 0) or original indices (1); -1 = none.
// spec (k > 1 only): search inside sqrt(bound2) first; queries with fewer than k neighbours there are appended to `list`
// (*count of them, at most `capacity` stored) for the caller to rerun without the bound.  Their rows hold what was found.
// bound2 > 0: that bound for every query.  bound2 == 0: a bound from the density around the query (the radius expected to hold 2.5 k
// points of a surface sampled like the query's own cell).  per_query: a query whose .w is positive brings its own bound in it.
struct KnnSpec {
    float bound2 = 0.f;
    int per_query = 0;
    uint32_t* list = nullptr;
    unsigned* count = nullptr;
    unsigned capacity = 0;
};
cudaError_t launch_knn(const GridView& g, const float4* d_queries, const int* d_nq, int nq_capacity,
                       const IcpState* d_state_or_null, int k, float max_r2, int32_t* out_ids,
                       float* out_d2, int want_original_ids, int variant, cudaStream_t s, KnnSpec spec = KnnSpec());

// ---- selfknn.cu ----------------------------------------------------------------------------
// Self k-NN of the cell-sorted cloud (queries = g.pts, row i of the output = position i), k <= 16, with TMA-staged candidate
// tiles.  Queries whose k-th distance leaves their tile's halo are listed in fb_list (*fb_count of them, possibly more than
// fb_capacity: then the list is truncated and the caller redoes everything with launch_knn); their rows are redone with
// launch_knn and put back with launch_selfknn_scatter.
cudaError_t launch_selfknn_tiles(const GridView& g, int k, int32_t* out_ids, float* out_d2, uint32_t* fb_list, unsigned* fb_count,
                                 unsigned fb_capacity, cudaStream_t s);
cudaError_t launch_selfknn_scatter(const uint32_t* list, const unsigned* n_list, unsigned capacity, int k, const int32_t* ids, const float* d2,
                                   int32_t* out_ids, float* out_d2, cudaStream_t s);

// Warm k = 1 search for ICP iterations >= 1: match_pos holds the previous matches on entry.
cudaError_t launch_nn1_warm(const GridView& g, const float4* d_reading, int nq_capacity, const IcpState* st,
                            float max_r2, int32_t* match_pos, float* match_d2, int variant, cudaStream_t s);

// ---- icp.cu --------------------------------------------------------------------------------
struct IcpBuffers {
    float* reading_in = nullptr;   // raw (dim+1) x N upload
    float4* reading = nullptr;     // reading in the refMean frame (T_refMean_dataIn applied)
    float4* reading_tmp = nullptr; // pre-sort
    float* rnrm_in = nullptr;      // raw dim x N upload of the reading's normals (b200icp_register_normals)
    float4* rnrm = nullptr;        // ... rotated into the refMean frame, same order as `reading`
    float4* rnrm_tmp = nullptr;    // ... pre-sort
    int64_t cap_rnrm = 0;
    float* rmax_in = nullptr;      // raw upload of the reading's `maxSearchDist` descriptor (b200icp_register_descriptors)
    int64_t cap_rmax = 0;
    int32_t* match_pos = nullptr;  // knn x N
    float* match_d2 = nullptr;
    uint32_t* hist = nullptr;      // max_iter x 3 x kHistBins
    double* partials = nullptr;    // kMaxAccBlocks x kAccSlots
    IcpState* state = nullptr;
    float* trace = nullptr;        // max_iter x 16
    char* fastws = nullptr;        // loop.cu one-barrier iteration workspace (icp_loop_workspace_bytes())
    float4* spill_pp = nullptr;    // loop.cu match cache for the queries beyond the shared-memory capacity (cap_nq each)
    float4* spill_nv = nullptr;
    int64_t cap_nq = 0;
    int cap_iter = 0;
};

// outlier.cu: VarTrimmedDist's tuned ratio -> IcpState::dyn_quantile (sort + fp64 scan + argmin, all on the device)
struct VarTrimScratch {
    float *keys_in = nullptr, *keys_out = nullptr;
    double* cums = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    unsigned int* d_count = nullptr;
    long long cap = 0;
};
void var_trimmed_free(VarTrimScratch& v);
cudaError_t launch_var_trimmed_ratio(VarTrimScratch& v, const IcpParams& p, int filter_index, IcpBuffers& b, cudaStream_t s, int* launches);
// RobustOutlierFilter's scale estimate for iteration `it` (0-based) -> IcpState::robust_scale
cudaError_t launch_robust_scale(VarTrimScratch& v, const IcpParams& p, int filter_index, IcpBuffers& b, int it, cudaStream_t s, int* launches);

cudaError_t launch_prep_reading(const float* d_in, int rows, int dim, const float* Tpre16 /*host*/,
                                float4* d_out, const GridView* g_for_keys, uint32_t* d_keys,
                                uint32_t* d_vals, int64_t nq, cudaStream_t s, int coarse_shift = 0, unsigned int* d_pmax2_bits = nullptr,
                                const float* d_max_dist_desc = nullptr /* nq radii -> .w = r^2 */, int strict = 0);
// dim x N column-major normals -> float4, rotated by the rotation block of Tpre16 (descriptors named `normals` rotate with the cloud)
cudaError_t launch_prep_normals(const float* d_in, int dim, const float* Tpre16 /*host*/, float4* d_out, int64_t nq, cudaStream_t s);
cudaError_t launch_gather_reading(const float4* d_in, const uint32_t* d_perm, float4* d_out, int64_t nq,
                                  cudaStream_t s);
cudaError_t icp_device_setup();
// loop.cu: iterations 1.. of a k = 1 registration in one persistent cooperative kernel
// win3 = {gain, floor, max}: half-width of the quantile window = max(gain * |limit - previous limit|, floor * limit), tried when <= max * limit
size_t icp_loop_workspace_bytes();
// margin3 = {gain, min [m], max [cell edges]}: extra search radius = clamp(gain * the query's last motion, min, max * h)
cudaError_t launch_icp_loop(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                            int variant, const float* win3, const float* margin3, int64_t nq, cudaStream_t s);
// the same with RobustOutlierFilter compiled in (loop_robust.cu); launch_icp_loop forwards the chains that hold one
cudaError_t launch_icp_loop_robust(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                            int variant, const float* win3, const float* margin3, int64_t nq, cudaStream_t s);
// the part of the loop workspace that must be zero when the kernel starts (offset, bytes): cudaMemsetAsync before every launch
void icp_loop_workspace_zero_range(size_t* offset, size_t* bytes);
// ev_mid (optional): recorded between the select and the accumulate kernel (profiling).
cudaError_t launch_iteration_tail(const IcpParams& p, const GridIndex& g, IcpBuffers& b, int it,
                                  cudaStream_t s, int* launches, cudaEvent_t ev_mid, VarTrimScratch* var_scratch);
// rotate the `dim` rows starting at row `offset` of a point-major block with `stride` floats per point (observationDirections)
cudaError_t launch_rotate_rows(float* d_block, int stride, int offset, int dim, int64_t n, const float* T16, cudaStream_t s);
cudaError_t launch_transform(float* d_feat, int rows, int dim, float* d_normals, int64_t n,
                             const float* T16 /*host, 4x4*/, cudaStream_t s);

}  // namespace b200
