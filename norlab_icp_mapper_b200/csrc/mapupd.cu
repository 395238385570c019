// mapupd.cu -- the device-resident local map and the map-update steps that run on it.
//
// Reference behaviour replaced (all under /root/reference/norlab_icp_mapper):
//   Map::localPointCloud + CellManager (Map.h:38-46, RAMCellManager.cpp)   -> MapStore: every map point
//        stays in HBM; "loaded" marks membership of the local cloud, unloaded points are what the
//        reference parks in 20 m cells (Map.cpp:140-230) -- no per-point string keys, no host copies.
//   Map::loadCells / unloadCells (Map.cpp:71-128, 140-230)                   -> window_kernel flips flags
//   PointDistanceMapperModule::inPlaceUpdateMap (PointDistanceMapperModule.cpp:28-50)
//        -> 1-NN against the LIVE index (no second kd-tree), keep rule, ordered compaction, append
//   SurfaceNormalDataPointsFilter{knn} (examples/config.yaml:26-27 via Map.cpp:524)
//        -> self k-NN on the cell-sorted index + per-point covariance + Jacobi eigenvector
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"

namespace b200 {
namespace {

__global__ void __launch_bounds__(256) to_store_kernel(const float* __restrict__ in, int rows, int dim, long long n,
                                                       float4* __restrict__ out, uint8_t* __restrict__ loaded) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float4(in[i * rows + 0], in[i * rows + 1], dim == 3 ? in[i * rows + 2] : 0.f, 1.f);
    loaded[i] = 1;
}

// keep[i] = 1 iff dist2(input_i, its nearest map point) >= minDistNewPoint^2, distance evaluated on
// the map-frame coordinates exactly as the reference's libnabo call does (no centring involved).
__global__ void __launch_bounds__(256) pd_keep_kernel(const float* __restrict__ in, int rows, int dim, long long n_in,
                                                      const float4* __restrict__ store, const int32_t* __restrict__ nn_id,
                                                      float thr2, uint32_t* __restrict__ keep) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const int id = nn_id[i];
    uint32_t k = 1u;  // no neighbour at all (empty map): dist = +inf >= thr
    if (id >= 0) {
        const float4 q = store[id];
        const float dx = __fsub_rn(in[i * rows + 0], q.x), dy = __fsub_rn(in[i * rows + 1], q.y);
        const float dz = dim == 3 ? __fsub_rn(in[i * rows + 2], q.z) : 0.f;
        const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        k = d2 >= thr2 ? 1u : 0u;
    }
    keep[i] = k;
}

__global__ void __launch_bounds__(256) pd_append_kernel(const float* __restrict__ in, int rows, int dim, const float* __restrict__ in_nrm,
                                                        long long n_in, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ offs,
                                                        long long base, float4* __restrict__ store, float* __restrict__ store_nrm,
                                                        const float* __restrict__ in_prob, float* __restrict__ store_prob,
                                                        uint8_t* __restrict__ loaded, uint8_t* __restrict__ keep_out,
                                                        const float* __restrict__ in_extra, float* __restrict__ store_extra, int extra_rows) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    if (keep_out) keep_out[i] = (uint8_t)keep[i];
    if (!keep[i]) return;
    const long long dst = base + offs[i];
    store[dst] = make_float4(in[i * rows + 0], in[i * rows + 1], dim == 3 ? in[i * rows + 2] : 0.f, 1.f);
    loaded[dst] = 1;
    if (store_nrm && in_nrm)
        for (int c = 0; c < dim; ++c) store_nrm[dst * dim + c] = in_nrm[i * dim + c];
    if (store_prob && in_prob) store_prob[dst] = in_prob[i];
    if (store_extra && in_extra)
        for (int c = 0; c < extra_rows; ++c) store_extra[dst * extra_rows + c] = in_extra[i * extra_rows + c];
}

// unload: loaded points inside the slab's metric AABB leave the local cloud (Map.cpp:161-174);
// load: parked points whose 20 m grid coordinate lies in the slab come back (Map.cpp:79-99, cell ids
// are floor(x / CELL_SIZE), Map.cpp:232-235).
__global__ void __launch_bounds__(256) window_kernel(const float4* __restrict__ store, long long n, uint8_t* __restrict__ loaded,
                                                     uint8_t* __restrict__ touched, int load, float cell, int r0, int r1, int c0, int c1,
                                                     int a0, int a1, unsigned long long* __restrict__ changed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = store[i];
    if (load) {
        if (loaded[i]) return;
        const int gx = (int)floorf(p.x / cell), gy = (int)floorf(p.y / cell), gz = (int)floorf(p.z / cell);
        if (gx >= r0 && gx <= r1 && gy >= c0 && gy <= c1 && gz >= a0 && gz <= a1) {
            loaded[i] = 1;
            touched[i] = 1;
            atomicAdd(changed, 1ull);
        }
    } else {
        if (!loaded[i]) return;
        const float sx = (float)r0 * cell, ex = ((float)r1 + 1.f) * cell;
        const float sy = (float)c0 * cell, ey = ((float)c1 + 1.f) * cell;
        const float sz = (float)a0 * cell, ez = ((float)a1 + 1.f) * cell;
        if (p.x >= sx && p.x < ex && p.y >= sy && p.y < ey && p.z >= sz && p.z < ez) {
            loaded[i] = 0;
            touched[i] = 1;
            atomicAdd(changed, 1ull);
        }
    }
}

__global__ void __launch_bounds__(256) flags_to_u32_kernel(const uint8_t* __restrict__ f, long long n, uint32_t* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = f[i] ? 1u : 0u;
}

__global__ void __launch_bounds__(256) scatter_active_kernel(const uint8_t* __restrict__ f, const uint32_t* __restrict__ offs, long long n,
                                                             uint32_t* __restrict__ active) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && f[i]) active[offs[i]] = (uint32_t)i;
}

// One Jacobi rotation on the symmetric 3x3 (a00 a01 a02 a11 a12 a22) with eigenvector columns v.
#define B200_JROT(app, aqq, apq, apr, aqr, vp0, vp1, vp2, vq0, vq1, vq2)                         \
    if (fabs(apq) >= 1e-300) {                                                                   \
        const double theta = (aqq - app) / (2.0 * apq);                                          \
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));  \
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;                                     \
        const double napp = app - t * apq, naqq = aqq + t * apq;                                 \
        const double napr = c * apr - s * aqr, naqr = s * apr + c * aqr;                         \
        app = napp; aqq = naqq; apq = 0.0; apr = napr; aqr = naqr;                               \
        double tv;                                                                               \
        tv = c * vp0 - s * vq0; vq0 = s * vp0 + c * vq0; vp0 = tv;                               \
        tv = c * vp1 - s * vq1; vq1 = s * vp1 + c * vq1; vp1 = tv;                               \
        tv = c * vp2 - s * vq2; vq2 = s * vp2 + c * vq2; vp2 = tv;                               \
    }

// LPM SurfaceNormalDataPointsFilter: over the finite neighbours (the point itself included): mean,
// centred covariance, eigenvector of the smallest eigenvalue (unit, sign arbitrary).
// list (optional): the cell-sorted positions to (re)compute, n_list of them; row i of nn_pos / nn_d2 then belongs to
// position list[i].  kth (optional, per STORE index): squared distance to the k-th neighbour (+inf when fewer than k
// exist) -- what the incremental update compares new points against.
__global__ void __launch_bounds__(128) normals_kernel(GridView g, int dim, int knn, const int32_t* __restrict__ nn_pos,
                                                      const float* __restrict__ nn_d2, const uint32_t* __restrict__ list,
                                                      const unsigned int* __restrict__ n_list, float4* __restrict__ nrm_sorted,
                                                      float* __restrict__ store_nrm, float* __restrict__ kth) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (list ? (long long)*n_list : (long long)g.n)) return;
    const long long j = list ? (long long)list[i] : i;
    float mx = 0.f, my = 0.f, mz = 0.f;
    int cnt = 0;
    nn_pos += (i - j) * knn;  // rows are indexed by i, the code below by j
    for (int c = 0; c < knn; ++c) {
        const int p = nn_pos[j * knn + c];
        if (p < 0) continue;
        const float4 q = __ldg(g.pts + p);
        mx += q.x;
        my += q.y;
        mz += q.z;
        ++cnt;
    }
    mx /= (float)cnt;
    my /= (float)cnt;
    mz /= (float)cnt;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int c = 0; c < knn; ++c) {
        const int p = nn_pos[j * knn + c];
        if (p < 0) continue;
        const float4 q = __ldg(g.pts + p);
        const float dx = q.x - mx, dy = q.y - my, dz = q.z - mz;
        a00 += (double)(dx * dx);
        a01 += (double)(dx * dy);
        a02 += (double)(dx * dz);
        a11 += (double)(dy * dy);
        a12 += (double)(dy * dz);
        a22 += (double)(dz * dz);
    }
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;  // vXY: component Y of eigenvector X
    if (dim == 2) {
        a02 = a12 = 0.0;
        a22 = 1e300;  // never the smallest
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        if (a01 * a01 + a02 * a02 + a12 * a12 < 1e-300) break;
        B200_JROT(a00, a11, a01, a02, a12, v00, v01, v02, v10, v11, v12)
        B200_JROT(a00, a22, a02, a01, a12, v00, v01, v02, v20, v21, v22)
        B200_JROT(a11, a22, a12, a01, a02, v10, v11, v12, v20, v21, v22)
    }
    double nx = v00, ny = v01, nz = v02, best = a00;
    if (a11 < best) {
        best = a11;
        nx = v10;
        ny = v11;
        nz = v12;
    }
    if (a22 < best) {
        nx = v20;
        ny = v21;
        nz = v22;
    }
    nrm_sorted[j] = make_float4((float)nx, (float)ny, (float)nz, 0.f);
    const long long orig = __float_as_int(__ldg(g.pts + j).w);
    store_nrm[orig * dim + 0] = (float)nx;
    store_nrm[orig * dim + 1] = (float)ny;
    if (dim == 3) store_nrm[orig * dim + 2] = (float)nz;
    if (kth) kth[orig] = nn_d2[i * knn + (knn - 1)];  // ascending; +inf when the k-th neighbour does not exist
}

// Incremental SurfaceNormal, step 1: which OLD points (store index < n_old, loaded) have a CHANGED point -- appended, or
// moved into / out of the window -- within their k-th neighbour distance?  Their k-NN set, hence their normal, changes;
// everybody else's does not.  `nw` is a grid over the changed points only (map frame, not centred); one thread per old
// point scans the cells its ball touches, first hit wins.
__global__ void __launch_bounds__(256) normals_dirty_kernel(GridView nw, const float4* __restrict__ feat, const uint8_t* __restrict__ loaded,
                                                            const uint8_t* __restrict__ touched, const float* __restrict__ kth, long long n_old,
                                                            long long n_all, uint8_t* __restrict__ dirty) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    if (i >= n_old || touched[i]) {  // a new point, or one that came back into the window: always computed (if loaded)
        dirty[i] = loaded[i] ? 1 : 0;
        return;
    }
    uint8_t d = 0;
    if (loaded[i]) {
        const float4 q = feat[i];
        const float r2 = kth[i];
        if (!(r2 < 3.0e38f)) {
            d = 1;  // fewer than k neighbours so far (or unknown): any new point changes the set
        } else {
            const float r = sqrtf(r2);
            const float ux = (q.x - nw.ox) * nw.inv_h, uy = (q.y - nw.oy) * nw.inv_h, uz = (q.z - nw.oz) * nw.inv_h;
            const float slack = nw.slack + 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
            const float rt = r * nw.inv_h + slack;
            // reject early when the ball misses the new points' bounding box altogether
            if (ux + rt >= 0.f && uy + rt >= 0.f && uz + rt >= 0.f && ux - rt <= (float)nw.nx && uy - rt <= (float)nw.ny && uz - rt <= (float)nw.nz) {
                const int xa = max(0, (int)floorf(ux - rt)), xb = min(nw.nx - 1, (int)floorf(ux + rt));
                const int ylo = max(0, (int)floorf(uy - rt)), yhi = min(nw.ny - 1, (int)floorf(uy + rt));
                const int zlo = max(0, (int)floorf(uz - rt)), zhi = min(nw.nz - 1, (int)floorf(uz + rt));
                for (int z = zlo; z <= zhi && !d; ++z)
                    for (int y = ylo; y <= yhi && !d; ++y) {
                        if (xa > xb) continue;
                        const uint32_t* row = nw.cell_start + ((size_t)z * nw.ny + y) * (size_t)nw.nx;
                        const uint32_t s = __ldg(row + xa), e = __ldg(row + xb + 1);
                        for (uint32_t j = s; j < e; ++j) {
                            const float4 p = __ldg(nw.pts + j);
                            const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
                            if (dx * dx + dy * dy + dz * dz <= r2 * 1.000001f + 1e-12f) {  // ties and rounding count as changes
                                d = 1;
                                break;
                            }
                        }
                    }
            }
        }
    }
    dirty[i] = d;
}

// step 2: dirty flags per store index -> flags per cell-sorted position of the live index
__global__ void __launch_bounds__(256) normals_flag_positions_kernel(GridView g, const uint8_t* __restrict__ dirty, uint8_t* __restrict__ flag) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    flag[j] = dirty[__float_as_int(__ldg(g.pts + j).w)];
}

// kth (optional, per store index, valid below n_old): the point's squared k-th neighbour distance of the last pass -- appends can only
// shrink it, so it bounds the new search (written to .w, one ulp up so that a tie at the bound stays accepted); -1 = no bound known
__global__ void __launch_bounds__(256) normals_gather_queries_kernel(GridView g, const uint32_t* __restrict__ list, const unsigned int* __restrict__ n_list,
                                                                     float4* __restrict__ q, const float* __restrict__ kth, long long n_old) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)*n_list) return;
    float4 p = __ldg(g.pts + list[i]);
    const long long orig = __float_as_int(p.w);
    float b = -1.f;
    if (kth && orig < n_old) {
        const float v = kth[orig];
        if (v > 0.f && v < 3.0e38f) b = __uint_as_float(__float_as_uint(v) + 1u);
    }
    p.w = b;
    q[i] = p;
}

template <typename T>
cudaError_t regrow(T*& p, int64_t old_count, int64_t new_cap, cudaStream_t s) {
    T* np = nullptr;
    cudaError_t e = B200_CUDA_MALLOC((void**)&np, (size_t)std::max<int64_t>(new_cap, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && old_count > 0) {
        e = cudaMemcpyAsync(np, p, (size_t)old_count * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
    }
    if (p) B200_CUDA_FREE(p);
    p = np;
    return cudaSuccess;
}

unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void store_free(MapStore& m) {
    B200_CUDA_FREE(m.feat);
    B200_CUDA_FREE(m.nrm);
    B200_CUDA_FREE(m.prob);
    B200_CUDA_FREE(m.extra);
    B200_CUDA_FREE(m.extra2);
    B200_CUDA_FREE(m.loaded);
    B200_CUDA_FREE(m.touched);
    B200_CUDA_FREE(m.feat2);
    B200_CUDA_FREE(m.nrm2);
    B200_CUDA_FREE(m.prob2);
    B200_CUDA_FREE(m.loaded2);
    B200_CUDA_FREE(m.keys64_a);
    B200_CUDA_FREE(m.keys64_b);
    B200_CUDA_FREE(m.active);
    B200_CUDA_FREE(m.tmp_u32a);
    B200_CUDA_FREE(m.tmp_u32b);
    B200_CUDA_FREE(m.d_counter);
    m = MapStore{};
}

cudaError_t store_reserve(MapStore& m, int dim, int64_t n, cudaStream_t s) {
    cudaError_t e;
    if (n > m.cap) {
        const int64_t cap = grow_capacity(n);
        if ((e = regrow(m.feat, m.n, cap, s)) != cudaSuccess) return e;
        if ((e = regrow(m.nrm, m.n * dim, cap * dim, s)) != cudaSuccess) return e;
        if ((e = regrow(m.prob, m.n, cap, s)) != cudaSuccess) return e;
        if ((e = regrow(m.loaded, m.n, cap, s)) != cudaSuccess) return e;
        {   // "loaded flag flipped since the last SurfaceNormal pass" (incremental normals); new tail zeroed
            const int64_t old_cap = m.cap;
            if ((e = regrow(m.touched, m.touched ? old_cap : 0, cap, s)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(m.touched + old_cap, 0, (size_t)(cap - old_cap), s)) != cudaSuccess) return e;
        }
        B200_CUDA_FREE(m.feat2);
        B200_CUDA_FREE(m.nrm2);
        B200_CUDA_FREE(m.prob2);
        B200_CUDA_FREE(m.loaded2);
        m.feat2 = nullptr;
        m.nrm2 = m.prob2 = nullptr;
        m.loaded2 = nullptr;
        if ((e = B200_CUDA_MALLOC((void**)&m.feat2, (size_t)cap * sizeof(float4))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&m.nrm2, (size_t)cap * dim * sizeof(float))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&m.prob2, (size_t)cap * sizeof(float))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&m.loaded2, (size_t)cap)) != cudaSuccess) return e;
        B200_CUDA_FREE(m.active);
        m.active = nullptr;
        if ((e = B200_CUDA_MALLOC((void**)&m.active, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
        m.cap = cap;
    }
    if (m.extra_rows > 0 && m.cap * m.extra_rows > m.cap_extra) {
        const int64_t want = m.cap * m.extra_rows;
        if ((e = regrow(m.extra, m.n * m.extra_rows, want, s)) != cudaSuccess) return e;
        B200_CUDA_FREE(m.extra2);
        m.extra2 = nullptr;
        if ((e = B200_CUDA_MALLOC((void**)&m.extra2, (size_t)want * sizeof(float))) != cudaSuccess) return e;
        m.cap_extra = want;
    }
    if (!m.d_counter)
        if ((e = B200_CUDA_MALLOC((void**)&m.d_counter, 64)) != cudaSuccess) return e;
    return cudaSuccess;
}

static cudaError_t ensure_tmp(MapStore& m, int64_t n) {
    if (n <= m.cap_tmp) return cudaSuccess;
    B200_CUDA_FREE(m.tmp_u32a);
    B200_CUDA_FREE(m.tmp_u32b);
    m.tmp_u32a = m.tmp_u32b = nullptr;
    m.cap_tmp = 0;
    const int64_t cap = grow_capacity(n);
    cudaError_t e;
    if ((e = B200_CUDA_MALLOC((void**)&m.tmp_u32a, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = B200_CUDA_MALLOC((void**)&m.tmp_u32b, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
    m.cap_tmp = cap;
    return cudaSuccess;
}

// b200icp_map_reserve: the compaction scratch of the update steps, sized once (its regrowth cudaFree stalls the update path)
cudaError_t store_reserve_scratch(MapStore& m, int64_t n) { return ensure_tmp(m, n + 1); }

static cudaError_t exclusive_sum(GridIndex& scratch, const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n);
    if (need > scratch.cub_tmp_bytes) {
        B200_CUDA_FREE(scratch.cub_tmp);
        scratch.cub_tmp = nullptr;
        scratch.cub_tmp_bytes = 0;
        cudaError_t e = B200_CUDA_MALLOC(&scratch.cub_tmp, need + 256);
        if (e != cudaSuccess) return e;
        scratch.cub_tmp_bytes = need + 256;
    }
    size_t bytes = scratch.cub_tmp_bytes;
    return cub::DeviceScan::ExclusiveSum(scratch.cub_tmp, bytes, in, out, (int)n, s);
}

cudaError_t store_set_extra_rows(MapStore& m, int rows, cudaStream_t s) {
    if (rows == m.extra_rows) return cudaSuccess;
    m.extra_rows = rows;  // (content dropped: the caller refills it)
    if (rows > 0 && m.cap * rows > m.cap_extra) {
        B200_CUDA_FREE(m.extra);
        B200_CUDA_FREE(m.extra2);
        m.extra = m.extra2 = nullptr;
        m.cap_extra = 0;
        const int64_t want = std::max<int64_t>(m.cap, 1) * rows;
        cudaError_t e;
        if ((e = B200_CUDA_MALLOC((void**)&m.extra, (size_t)want * sizeof(float))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&m.extra2, (size_t)want * sizeof(float))) != cudaSuccess) return e;
        m.cap_extra = want;
    }
    (void)s;
    return cudaSuccess;
}

namespace {
struct RowList {
    int r[B200ICP_MAX_EXTRA_ROWS];
};
__global__ void __launch_bounds__(256) select_extra_kernel(const float* __restrict__ in, int in_rows, long long n, float* __restrict__ out, int out_rows,
                                                           RowList rr) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < out_rows; ++c) out[i * out_rows + c] = in[i * in_rows + rr.r[c]];
}
__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ out, float value, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = value;
}
}  // namespace

cudaError_t launch_select_rows(const float* d_in, int in_rows, int64_t n, float* d_out, int out_rows, const int* rows, cudaStream_t s) {
    if (out_rows < 0 || out_rows > B200ICP_MAX_EXTRA_ROWS) return cudaErrorInvalidValue;
    if (n == 0 || out_rows == 0) return cudaSuccess;
    RowList rr{};
    for (int i = 0; i < out_rows; ++i) rr.r[i] = rows[i];
    select_extra_kernel<<<blocks_for(n), 256, 0, s>>>(d_in, in_rows, (long long)n, d_out, out_rows, rr);
    return cudaGetLastError();
}

cudaError_t launch_fill(float* d_out, float value, int64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    fill_kernel<<<blocks_for(n), 256, 0, s>>>(d_out, value, (long long)n);
    return cudaGetLastError();
}

cudaError_t store_select_extra(MapStore& m, const int* rows, int n_rows, cudaStream_t s) {
    if (n_rows < 0 || n_rows > B200ICP_MAX_EXTRA_ROWS) return cudaErrorInvalidValue;
    for (int i = 0; i < n_rows; ++i)
        if (rows[i] < 0 || rows[i] >= m.extra_rows) return cudaErrorInvalidValue;
    if (n_rows > 0 && m.n > 0) {
        cudaError_t e = launch_select_rows(m.extra, m.extra_rows, m.n, m.extra2, n_rows, rows, s);
        if (e != cudaSuccess) return e;
        std::swap(m.extra, m.extra2);
    }
    m.extra_rows = n_rows;
    return cudaSuccess;
}

cudaError_t store_set(MapStore& m, const DevCloud& in, int dim, cudaStream_t s) {
    cudaError_t e;
    m.n = 0;  // nothing to preserve
    const int64_t n = in.n;
    if ((e = store_set_extra_rows(m, in.extra ? in.extra_rows : 0, s)) != cudaSuccess) return e;
    if ((e = store_reserve(m, dim, n, s)) != cudaSuccess) return e;
    to_store_kernel<<<blocks_for(n), 256, 0, s>>>(in.feat, in.rows, dim, (long long)n, m.feat, m.loaded);
    m.has_normals = in.nrm != nullptr;
    m.nrm_epoch_ok = false;
    if (in.nrm)
        if ((e = cudaMemcpyAsync(m.nrm, in.nrm, (size_t)n * dim * sizeof(float), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    m.has_prob = in.prob != nullptr;  // (the reference assigns the whole cloud: every descriptor comes along, Map.cpp:575-588)
    if (in.prob)
        if ((e = cudaMemcpyAsync(m.prob, in.prob, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if (m.extra_rows > 0)
        if ((e = cudaMemcpyAsync(m.extra, in.extra, (size_t)n * m.extra_rows * sizeof(float), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    m.n = n;
    m.n_active = n;
    m.all_loaded = true;
    return cudaGetLastError();
}

cudaError_t store_compact_active(MapStore& m, GridIndex& scratch, cudaStream_t s) {
    cudaError_t e;
    if (m.n == 0) {
        m.n_active = 0;
        return cudaSuccess;
    }
    if ((e = ensure_tmp(m, m.n + 1)) != cudaSuccess) return e;
    flags_to_u32_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.loaded, (long long)m.n, m.tmp_u32a);
    if ((e = cudaMemsetAsync(m.tmp_u32a + m.n, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, m.tmp_u32a, m.tmp_u32b, m.n + 1, s)) != cudaSuccess) return e;
    scatter_active_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.loaded, m.tmp_u32b, (long long)m.n, m.active);
    uint32_t total = 0;
    if ((e = cudaMemcpyAsync(&total, m.tmp_u32b + m.n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    m.n_active = total;
    m.all_loaded = (int64_t)total == m.n;
    return cudaGetLastError();
}

cudaError_t store_window(MapStore& m, int load, const int32_t* slab6, int64_t* changed, cudaStream_t s) {
    *changed = 0;
    if (m.n == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemsetAsync(m.d_counter, 0, sizeof(unsigned long long), s)) != cudaSuccess) return e;
    window_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.feat, (long long)m.n, m.loaded, m.touched, load, 20.0f, slab6[0], slab6[1], slab6[2],
                                                  slab6[3], slab6[4], slab6[5], m.d_counter);
    unsigned long long c = 0;
    if ((e = cudaMemcpyAsync(&c, m.d_counter, sizeof(c), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    *changed = (int64_t)c;
    if (c) m.nrm_touched = true;  // neighbourhoods at the window's edge changed: those points are marked in `touched`
    return cudaGetLastError();
}

// DataPoints::concatenate keeps only descriptors present in both clouds; a first cloud (empty store) brings its own.
// Returns which descriptors the store keeps and prepares the `extra` block.
static cudaError_t concat_layout(MapStore& m, const DevCloud& in, bool first, bool& keep_n, bool& keep_p, bool& keep_x, cudaStream_t s) {
    keep_n = first ? (in.nrm != nullptr) : (m.has_normals && in.nrm != nullptr);
    keep_p = first ? (in.prob != nullptr) : (m.has_prob && in.prob != nullptr);
    const int in_rows = in.extra ? in.extra_rows : 0;
    if (first) {
        cudaError_t e = store_set_extra_rows(m, in_rows, s);
        if (e != cudaSuccess) return e;
    } else if (in_rows != m.extra_rows) {
        m.extra_rows = 0;  // layouts differ and the caller did not reconcile them (store_select_extra): no common descriptor
    }
    keep_x = m.extra_rows > 0;
    return cudaSuccess;
}

cudaError_t store_insert_point_distance(MapStore& m, GridIndex& scratch, const DevCloud& in, int dim, const int32_t* d_nn_id, float min_dist,
                                        int64_t* n_kept, uint8_t* d_keep_out, cudaStream_t s) {
    cudaError_t e;
    *n_kept = 0;
    const int64_t n_in = in.n;
    if (n_in == 0) return cudaSuccess;
    if ((e = ensure_tmp(m, n_in + 1)) != cudaSuccess) return e;
    pd_keep_kernel<<<blocks_for(n_in), 256, 0, s>>>(in.feat, in.rows, dim, (long long)n_in, m.feat, d_nn_id, min_dist * min_dist, m.tmp_u32a);
    if ((e = cudaMemsetAsync(m.tmp_u32a + n_in, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, m.tmp_u32a, m.tmp_u32b, n_in + 1, s)) != cudaSuccess) return e;
    uint32_t total = 0;
    if ((e = cudaMemcpyAsync(&total, m.tmp_u32b + n_in, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    bool keep_n, keep_p, keep_x;
    if ((e = concat_layout(m, in, m.n == 0, keep_n, keep_p, keep_x, s)) != cudaSuccess) return e;
    if ((e = store_reserve(m, dim, m.n + total, s)) != cudaSuccess) return e;
    pd_append_kernel<<<blocks_for(n_in), 256, 0, s>>>(in.feat, in.rows, dim, in.nrm, (long long)n_in, m.tmp_u32a, m.tmp_u32b, (long long)m.n, m.feat,
                                                      keep_n ? m.nrm : nullptr, in.prob, keep_p ? m.prob : nullptr, m.loaded, d_keep_out, in.extra,
                                                      keep_x ? m.extra : nullptr, m.extra_rows);
    m.has_normals = keep_n;
    m.has_prob = keep_p;
    m.n += total;
    m.n_active += total;
    *n_kept = total;
    return cudaGetLastError();
}

namespace {

__global__ void __launch_bounds__(256) ones_kernel(uint32_t* __restrict__ a, uint32_t* __restrict__ offs, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        a[i] = 1u;
        offs[i] = (uint32_t)i;
    }
}

// Descend the LPM octree (utils/octree: child = (p > centre) per axis, child centre = centre +- r/2,
// r halves) `depth` levels with the same fp32 operations; the path is the voxel key.
__global__ void __launch_bounds__(256) octree_key_kernel(const float4* __restrict__ feat, const uint32_t* __restrict__ active, long long n_active,
                                                         int dim, float cx, float cy, float cz, float radius, int depth,
                                                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n_active) return;
    const uint32_t i = active ? active[j] : (uint32_t)j;
    const float4 p = feat[i];
    float c0 = cx, c1 = cy, c2 = cz, r = radius;
    unsigned long long key = 0;
    for (int l = 0; l < depth; ++l) {
        const float h = r * 0.5f;
        const unsigned b0 = p.x > c0, b1 = p.y > c1, b2 = (dim == 3) ? (p.z > c2) : 0u;
        key = (key << 3) | (unsigned long long)(b0 | (b1 << 1) | (b2 << 2));
        c0 += b0 ? h : -h;
        c1 += b1 ? h : -h;
        if (dim == 3) c2 += b2 ? h : -h;
        r = h;
    }
    keys[j] = key;
    vals[j] = i;
}

// counter-based generator of the random sampler: which member of the leaf `key` survives (deterministic per seed)
__device__ __forceinline__ unsigned long long octree_mix64(unsigned long long x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// sorted by key (stable: a leaf's members are in insertion order).  One survivor per leaf (LPM OctreeGrid samplers):
//   0 first point; 2 centroid (written over the first point); 1 random member (uniform; upstream draws from its own
//   generator, here hash(leaf key, seed) -- same distribution, reproducible); 3 medoid = the member closest to the
//   leaf's centroid (first one on ties).  In modes 1 and 3 the head thread of a run flags the whole run.
__global__ void __launch_bounds__(256) octree_mark_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                          long long n_active, int dim, int mode, unsigned long long seed, float4* __restrict__ feat,
                                                          float* __restrict__ nrm, float* __restrict__ prob, uint32_t* __restrict__ remove,
                                                          float* __restrict__ extra, int extra_rows) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n_active) return;
    const bool head = (j == 0) || (keys[j] != keys[j - 1]);
    const uint32_t i = vals[j];
    const bool centroid = mode == 2;
    if (mode == 0 || mode == 2) remove[i] = head ? 0u : 1u;
    if (head && (mode == 1 || mode == 3)) {
        long long end = j;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        while (end < n_active && keys[end] == keys[j]) {
            if (mode == 3) {
                const float4 f = feat[vals[end]];
                sx += f.x;
                sy += f.y;
                sz += f.z;
            }
            ++end;
        }
        const long long cnt = end - j;
        long long pick = j;
        if (mode == 1) {
            pick = j + (long long)(octree_mix64(keys[j] ^ octree_mix64(seed)) % (unsigned long long)cnt);
        } else {
            const float inv = 1.f / (float)cnt;
            const float mx = sx * inv, my = sy * inv, mz = sz * inv;
            float best = CUDART_INF_F;
            for (long long t = j; t < end; ++t) {
                const float4 f = feat[vals[t]];
                const float dx = f.x - mx, dy = f.y - my, dz = (dim == 3) ? f.z - mz : 0.f;
                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                if (d < best) {
                    best = d;
                    pick = t;
                }
            }
        }
        for (long long t = j; t < end; ++t) remove[vals[t]] = (t == pick) ? 0u : 1u;
    }
    if (head && centroid) {
        float sx = 0.f, sy = 0.f, sz = 0.f, sp = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
        int cnt = 0;
        for (long long t = j; t < n_active && keys[t] == keys[j]; ++t) {
            const uint32_t q = vals[t];
            const float4 f = feat[q];
            sx += f.x;
            sy += f.y;
            sz += f.z;
            if (prob) sp += prob[q];
            if (nrm) {
                n0 += nrm[(long long)q * dim + 0];
                n1 += nrm[(long long)q * dim + 1];
                if (dim == 3) n2 += nrm[(long long)q * dim + 2];
            }
            ++cnt;
        }
        const float inv = 1.f / (float)cnt;
        feat[i] = make_float4(sx * inv, sy * inv, sz * inv, 1.f);
        if (prob) prob[i] = sp * inv;
        if (nrm) {
            nrm[(long long)i * dim + 0] = n0 * inv;
            nrm[(long long)i * dim + 1] = n1 * inv;
            if (dim == 3) nrm[(long long)i * dim + 2] = n2 * inv;
        }
        for (int c = 0; c < extra_rows; ++c) {  // the centroid sampler averages every descriptor (LPM OctreeGrid CentroidSampler)
            float sx2 = 0.f;
            for (long long t = j; t < n_active && keys[t] == keys[j]; ++t) sx2 += extra[(long long)vals[t] * extra_rows + c];
            extra[(long long)i * extra_rows + c] = sx2 * inv;
        }
    }
}

__global__ void __launch_bounds__(256) cut_prob_kernel(const float* __restrict__ prob, const uint8_t* __restrict__ loaded, long long n, float thr,
                                                       int larger, uint32_t* __restrict__ remove) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = prob[i];
    remove[i] = (loaded[i] && (larger ? (v > thr) : (v < thr))) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) invert_flags_kernel(uint32_t* __restrict__ f, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) f[i] = f[i] ? 0u : 1u;
}

__global__ void __launch_bounds__(256) compact_store_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ offs, long long n, int dim,
                                                            const float4* __restrict__ feat, const float* __restrict__ nrm, const float* __restrict__ prob,
                                                            const uint8_t* __restrict__ loaded, float4* __restrict__ feat2, float* __restrict__ nrm2,
                                                            float* __restrict__ prob2, uint8_t* __restrict__ loaded2,
                                                            const float* __restrict__ extra, float* __restrict__ extra2, int extra_rows) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const long long d = offs[i];
    feat2[d] = feat[i];
    loaded2[d] = loaded[i];
    if (nrm)
        for (int c = 0; c < dim; ++c) nrm2[d * dim + c] = nrm[i * dim + c];
    if (prob) prob2[d] = prob[i];
    if (extra)
        for (int c = 0; c < extra_rows; ++c) extra2[d * extra_rows + c] = extra[i * extra_rows + c];
}

// remove the store points whose flag in tmp_u32a is 1 (order of the survivors preserved); flagged_are_loaded: they count in n_active
cudaError_t store_remove_flagged(MapStore& m, GridIndex& scratch, int dim, int64_t* n_removed, cudaStream_t s, bool flagged_are_loaded = true) {
    cudaError_t e;
    const int64_t n = m.n;
    invert_flags_kernel<<<blocks_for(n), 256, 0, s>>>(m.tmp_u32a, (long long)n);  // now: keep flags
    if ((e = cudaMemsetAsync(m.tmp_u32a + n, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, m.tmp_u32a, m.tmp_u32b, n + 1, s)) != cudaSuccess) return e;
    uint32_t kept = 0;
    if ((e = cudaMemcpyAsync(&kept, m.tmp_u32b + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    compact_store_kernel<<<blocks_for(n), 256, 0, s>>>(m.tmp_u32a, m.tmp_u32b, (long long)n, dim, m.feat, m.has_normals ? m.nrm : nullptr,
                                                       m.has_prob ? m.prob : nullptr, m.loaded, m.feat2, m.nrm2, m.prob2, m.loaded2,
                                                       m.extra_rows > 0 ? m.extra : nullptr, m.extra2, m.extra_rows);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    std::swap(m.feat, m.feat2);
    std::swap(m.nrm, m.nrm2);
    std::swap(m.prob, m.prob2);
    std::swap(m.loaded, m.loaded2);
    if (m.extra_rows > 0) std::swap(m.extra, m.extra2);
    *n_removed = n - (int64_t)kept;
    m.nrm_epoch_ok = false;  // points moved / vanished: the incremental normals bookkeeping starts over
    m.n = kept;
    if (flagged_are_loaded) m.n_active -= *n_removed;
    return cudaGetLastError();
}

}  // namespace

namespace {
struct FilterChain {
    b200icp_filter f[8];
    int n;
};
__global__ void __launch_bounds__(256) filter_flags_kernel(const float* __restrict__ feat, int rows, int dim, long long n, FilterChain ch,
                                                           uint32_t* __restrict__ keep) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = feat[i * rows + 0], y = feat[i * rows + 1], z = dim == 3 ? feat[i * rows + 2] : 0.f;
    bool k = true;
    for (int t = 0; t < ch.n && k; ++t) {
        const b200icp_filter& f = ch.f[t];
        if (f.kind == B200ICP_FILTER_BOUNDING_BOX) {
            bool inside = x >= f.lo[0] && x <= f.hi[0] && y >= f.lo[1] && y <= f.hi[1];
            if (dim == 3) inside = inside && z >= f.lo[2] && z <= f.hi[2];
            k = f.remove_inside ? !inside : inside;
        } else if (f.kind == B200ICP_FILTER_RANDOM_SAMPLING) {
            // keep with probability `dist`: uniform in [0, 1) from a counter-based generator keyed by (seed = dim, chain slot, point index)
            unsigned long long x = ((unsigned long long)(uint32_t)f.dim << 40) ^ ((unsigned long long)t << 36) ^ (unsigned long long)i;
            x += 0x9e3779b97f4a7c15ull;
            x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
            x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
            x ^= x >> 31;
            k = (float)(x >> 40) * (1.0f / 16777216.0f) < f.dist;
        } else {
            float v, lim;
            if (f.dim == -1) {
                v = dim == 3 ? sqrtf(x * x + y * y + z * z) : sqrtf(x * x + y * y);
                lim = fabsf(f.dist);
            } else {
                v = f.dim == 0 ? x : (f.dim == 1 ? y : z);
                lim = f.dist;
            }
            k = f.remove_inside ? (v > lim) : (v < lim);
        }
    }
    keep[i] = k ? 1u : 0u;
}
__global__ void __launch_bounds__(256) filter_compact_kernel(const float* __restrict__ in, int rows, long long n, const uint32_t* __restrict__ keep,
                                                             const uint32_t* __restrict__ offs, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    for (int c = 0; c < rows; ++c) out[(long long)offs[i] * rows + c] = in[i * rows + c];
}
}  // namespace

cudaError_t filter_cloud_device(MapStore& tmp, GridIndex& scratch, const DevCloud& in, int dim, const b200icp_filter* chain, int n_filters,
                                float* out_feat, float* out_nrm, float* out_prob, float* out_extra, int64_t* n_out, cudaStream_t s) {
    cudaError_t e;
    *n_out = 0;
    const int64_t n = in.n;
    if (n == 0) return cudaSuccess;
    if (n_filters > 8) return cudaErrorInvalidValue;
    if ((e = ensure_tmp(tmp, n + 1)) != cudaSuccess) return e;
    FilterChain ch;
    ch.n = n_filters;
    for (int i = 0; i < n_filters; ++i) ch.f[i] = chain[i];
    filter_flags_kernel<<<blocks_for(n), 256, 0, s>>>(in.feat, in.rows, dim, (long long)n, ch, tmp.tmp_u32a);
    if ((e = cudaMemsetAsync(tmp.tmp_u32a + n, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, tmp.tmp_u32a, tmp.tmp_u32b, n + 1, s)) != cudaSuccess) return e;
    filter_compact_kernel<<<blocks_for(n), 256, 0, s>>>(in.feat, in.rows, (long long)n, tmp.tmp_u32a, tmp.tmp_u32b, out_feat);
    if (in.nrm && out_nrm) filter_compact_kernel<<<blocks_for(n), 256, 0, s>>>(in.nrm, dim, (long long)n, tmp.tmp_u32a, tmp.tmp_u32b, out_nrm);
    if (in.prob && out_prob) filter_compact_kernel<<<blocks_for(n), 256, 0, s>>>(in.prob, 1, (long long)n, tmp.tmp_u32a, tmp.tmp_u32b, out_prob);
    if (in.extra && out_extra && in.extra_rows > 0)
        filter_compact_kernel<<<blocks_for(n), 256, 0, s>>>(in.extra, in.extra_rows, (long long)n, tmp.tmp_u32a, tmp.tmp_u32b, out_extra);
    uint32_t total = 0;
    if ((e = cudaMemcpyAsync(&total, tmp.tmp_u32b + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    *n_out = total;
    return cudaGetLastError();
}

cudaError_t store_append_all(MapStore& m, const DevCloud& in, int dim, cudaStream_t s) {
    cudaError_t e;
    const int64_t n_in = in.n;
    if (n_in == 0) return cudaSuccess;
    bool keep_n, keep_p, keep_x;
    if ((e = concat_layout(m, in, m.n == 0, keep_n, keep_p, keep_x, s)) != cudaSuccess) return e;
    if ((e = store_reserve(m, dim, m.n + n_in, s)) != cudaSuccess) return e;
    if ((e = ensure_tmp(m, n_in + 1)) != cudaSuccess) return e;
    ones_kernel<<<blocks_for(n_in), 256, 0, s>>>(m.tmp_u32a, m.tmp_u32b, (long long)n_in);
    pd_append_kernel<<<blocks_for(n_in), 256, 0, s>>>(in.feat, in.rows, dim, in.nrm, (long long)n_in, m.tmp_u32a, m.tmp_u32b, (long long)m.n, m.feat,
                                                      keep_n ? m.nrm : nullptr, in.prob, keep_p ? m.prob : nullptr, m.loaded, nullptr, in.extra,
                                                      keep_x ? m.extra : nullptr, m.extra_rows);
    m.has_normals = keep_n;
    m.has_prob = keep_p;
    m.n += n_in;
    m.n_active += n_in;
    return cudaGetLastError();
}

cudaError_t store_octree_filter(MapStore& m, GridIndex& scratch, int dim, float max_size_by_node, int sampling_method, uint64_t seed,
                                int64_t* n_removed, cudaStream_t s) {
    cudaError_t e;
    *n_removed = 0;
    if (m.n_active == 0) return cudaSuccess;
    if (!m.all_loaded || m.n_active != m.n)
        if ((e = store_compact_active(m, scratch, s)) != cudaSuccess) return e;
    const uint32_t* active = m.all_loaded ? nullptr : m.active;
    const int64_t na = m.n_active;
    // bounding box of the filtered cloud (LPM Octree::build): centre = min + (max - min) / 2, radius = max extent / 2
    if ((e = ensure_scratch(scratch, na)) != cudaSuccess) return e;
    float lo[3], hi[3];
    if ((e = cloud_bounds(scratch, reinterpret_cast<const float*>(m.feat), 4, dim, na, active, lo, hi, s)) != cudaSuccess) return e;
    float radii[3] = {0, 0, 0}, centre[3] = {0, 0, 0}, radius = 0.f;
    for (int d = 0; d < dim; ++d) {
        radii[d] = hi[d] - lo[d];
        centre[d] = lo[d] + radii[d] * 0.5f;
        radius = d == 0 ? radii[0] : (radius < radii[d] ? radii[d] : radius);
    }
    radius *= 0.5f;
    int depth = 0;
    for (float r = radius; (double)r * 2.0 > (double)max_size_by_node && depth < 22; r *= 0.5f) ++depth;
    if (depth > 21) return cudaErrorInvalidValue;  // key would not fit 63 bits
    if (na > m.cap_keys64) {
        B200_CUDA_FREE(m.keys64_a);
        B200_CUDA_FREE(m.keys64_b);
        m.keys64_a = m.keys64_b = nullptr;
        m.cap_keys64 = 0;
        const int64_t cap = grow_capacity(na);
        if ((e = B200_CUDA_MALLOC((void**)&m.keys64_a, (size_t)cap * 8)) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&m.keys64_b, (size_t)cap * 8)) != cudaSuccess) return e;
        m.cap_keys64 = cap;
    }
    if ((e = ensure_tmp(m, std::max<int64_t>(m.n, na) + 1)) != cudaSuccess) return e;
    octree_key_kernel<<<blocks_for(na), 256, 0, s>>>(m.feat, active, (long long)na, dim, centre[0], centre[1], centre[2], radius, depth,
                                                     m.keys64_a, scratch.vals_in);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, m.keys64_a, m.keys64_b, scratch.vals_in, scratch.vals_out, (int)na, 0, std::max(1, 3 * depth));
    if (need > scratch.cub_tmp_bytes) {
        B200_CUDA_FREE(scratch.cub_tmp);
        scratch.cub_tmp = nullptr;
        scratch.cub_tmp_bytes = 0;
        if ((e = B200_CUDA_MALLOC(&scratch.cub_tmp, need + 256)) != cudaSuccess) return e;
        scratch.cub_tmp_bytes = need + 256;
    }
    size_t bytes = scratch.cub_tmp_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(scratch.cub_tmp, bytes, m.keys64_a, m.keys64_b, scratch.vals_in, scratch.vals_out, (int)na, 0,
                                             std::max(1, 3 * depth), s)) != cudaSuccess)
        return e;
    if ((e = cudaMemsetAsync(m.tmp_u32a, 0, (size_t)(m.n + 1) * sizeof(uint32_t), s)) != cudaSuccess) return e;
    octree_mark_kernel<<<blocks_for(na), 256, 0, s>>>(m.keys64_b, scratch.vals_out, (long long)na, dim, sampling_method, (unsigned long long)seed, m.feat,
                                                      m.has_normals ? m.nrm : nullptr, m.has_prob ? m.prob : nullptr, m.tmp_u32a,
                                                      m.extra_rows > 0 ? m.extra : nullptr, m.extra_rows > 0 ? m.extra_rows : 0);
    return store_remove_flagged(m, scratch, dim, n_removed, s);
}

namespace {
__global__ void __launch_bounds__(256) loaded_flags_kernel(const uint8_t* __restrict__ loaded, long long n, uint32_t* __restrict__ remove) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) remove[i] = loaded[i] ? 1u : 0u;
}
}  // namespace

// localPointCloud = cloud (a host-signature MapperModule returned it): the loaded points go, the cloud's points come, parked points
// stay.  With nothing parked the cloud's descriptor set becomes the map's; else concatenate's intersection rule applies.
cudaError_t store_replace_loaded(MapStore& m, GridIndex& scratch, const DevCloud& in, int dim, cudaStream_t s) {
    cudaError_t e;
    if (m.n > 0) {
        if (m.n_active == m.n) {
            m.n = 0;
            m.n_active = 0;
        } else if (m.n_active > 0) {
            if ((e = ensure_tmp(m, m.n + 1)) != cudaSuccess) return e;
            loaded_flags_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.loaded, (long long)m.n, m.tmp_u32a);
            int64_t removed = 0;
            if ((e = store_remove_flagged(m, scratch, dim, &removed, s)) != cudaSuccess) return e;
        }
    }
    m.nrm_epoch_ok = false;
    if (in.n == 0) return cudaSuccess;
    return store_append_all(m, in, dim, s);
}

namespace {
__global__ void __launch_bounds__(256) parked_flags_kernel(const uint8_t* __restrict__ loaded, long long n, uint32_t* __restrict__ flag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = loaded[i] ? 0u : 1u;
}
}  // namespace

// Spill tier under the device grid (CellManager seam): the parked points (loaded = 0), compacted in insertion order into the
// secondary buffers feat2 / nrm2 / prob2 / extra2 for the caller to copy out ...
cudaError_t store_extract_parked(MapStore& m, GridIndex& scratch, int dim, int64_t* n_parked, cudaStream_t s) {
    cudaError_t e;
    *n_parked = 0;
    if (m.n == 0 || m.n_active == m.n) return cudaSuccess;
    const int64_t n = m.n;
    if ((e = ensure_tmp(m, n + 1)) != cudaSuccess) return e;
    parked_flags_kernel<<<blocks_for(n), 256, 0, s>>>(m.loaded, (long long)n, m.tmp_u32a);
    if ((e = cudaMemsetAsync(m.tmp_u32a + n, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, m.tmp_u32a, m.tmp_u32b, n + 1, s)) != cudaSuccess) return e;
    uint32_t total = 0;
    if ((e = cudaMemcpyAsync(&total, m.tmp_u32b + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    compact_store_kernel<<<blocks_for(n), 256, 0, s>>>(m.tmp_u32a, m.tmp_u32b, (long long)n, dim, m.feat, m.has_normals ? m.nrm : nullptr,
                                                       m.has_prob ? m.prob : nullptr, m.loaded, m.feat2, m.nrm2, m.prob2, m.loaded2,
                                                       m.extra_rows > 0 ? m.extra : nullptr, m.extra2, m.extra_rows);
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    *n_parked = total;
    return cudaGetLastError();
}

// ... and their removal from the store (tmp_u32a still holds the flags of store_extract_parked)
cudaError_t store_remove_parked(MapStore& m, GridIndex& scratch, int dim, cudaStream_t s) {
    if (m.n == 0 || m.n_active == m.n) return cudaSuccess;
    int64_t removed = 0;
    cudaError_t e = store_remove_flagged(m, scratch, dim, &removed, s, /*flagged_are_loaded=*/false);
    if (e == cudaSuccess) m.all_loaded = true;
    return e;
}

cudaError_t store_cut_prob(MapStore& m, GridIndex& scratch, int dim, float threshold, int use_larger_than, int64_t* n_removed, cudaStream_t s) {
    cudaError_t e;
    *n_removed = 0;
    if (m.n == 0 || !m.has_prob) return cudaSuccess;
    if ((e = ensure_tmp(m, m.n + 1)) != cudaSuccess) return e;
    cut_prob_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.prob, m.loaded, (long long)m.n, threshold, use_larger_than, m.tmp_u32a);
    return store_remove_flagged(m, scratch, dim, n_removed, s);
}

namespace {

struct M16 {
    float m[16];
};

__device__ __forceinline__ float3 xform(const M16& T, float x, float y, float z) {
    float3 o;
    o.x = __fadd_rn(__fmaf_rn(T.m[8], z, __fmaf_rn(T.m[4], y, __fmul_rn(T.m[0], x))), T.m[12]);
    o.y = __fadd_rn(__fmaf_rn(T.m[9], z, __fmaf_rn(T.m[5], y, __fmul_rn(T.m[1], x))), T.m[13]);
    o.z = __fadd_rn(__fmaf_rn(T.m[10], z, __fmaf_rn(T.m[6], y, __fmul_rn(T.m[2], x))), T.m[14]);
    return o;
}

// convertToSphericalCoordinates (DynamicPointsMapperModule.cpp:156-172)
__device__ __forceinline__ void spherical(float3 p, int dim, float& r, float& el, float& az) {
    r = (dim == 3) ? sqrtf(p.x * p.x + p.y * p.y + p.z * p.z) : sqrtf(p.x * p.x + p.y * p.y);
    el = (dim == 3) ? asinf(p.z / r) : 0.f;
    az = atan2f(p.y, p.x);
}

__global__ void __launch_bounds__(256) dyn_input_kernel(const float* __restrict__ in, int rows, int dim, M16 Tinv, long long n_in,
                                                        float4* __restrict__ in_sensor, float* __restrict__ angles) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const float3 p = xform(Tinv, in[i * rows + 0], in[i * rows + 1], dim == 3 ? in[i * rows + 2] : 0.f);
    float r, el, az;
    spherical(p, dim, r, el, az);
    in_sensor[i] = make_float4(p.x, p.y, p.z, r);
    angles[i * 2 + 0] = el;
    angles[i * 2 + 1] = az;
}

__global__ void __launch_bounds__(256) dyn_query_kernel(const float4* __restrict__ feat, const uint32_t* __restrict__ active, long long na, int dim,
                                                        M16 Tinv, float range, float4* __restrict__ q4) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= na) return;
    const float4 f = feat[active ? active[j] : (uint32_t)j];
    const float3 p = xform(Tinv, f.x, f.y, f.z);
    float r, el, az;
    spherical(p, dim, r, el, az);
    const float nanv = __int_as_float(0x7fc00000);
    q4[j] = (r < range) ? make_float4(el, az, 0.f, 1.f) : make_float4(nanv, nanv, nanv, 1.f);  // NaN -> no neighbours
}

// DynamicPointsMapperModule.cpp:86-150, same promotion to double where the reference writes `1.`
__global__ void __launch_bounds__(256) dyn_update_kernel(const float4* __restrict__ feat, const uint32_t* __restrict__ active, long long na, int dim,
                                                         M16 Tinv, DynParams prm, const float* __restrict__ nrm, float* __restrict__ prob,
                                                         const float4* __restrict__ in_sensor, const int32_t* __restrict__ ids,
                                                         const float* __restrict__ d2) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= na) return;
    const float dist = d2[j];
    const int id = ids[j];
    if (id < 0 || dist == CUDART_INF_F) return;
    const uint32_t i = active ? active[j] : (uint32_t)j;
    const float eps = 0.0001f;
    const float4 f = feat[i];
    const float3 lp = xform(Tinv, f.x, f.y, f.z);
    const float4 ip = in_sensor[id];
    const float lpNorm = (dim == 3) ? sqrtf(lp.x * lp.x + lp.y * lp.y + lp.z * lp.z) : sqrtf(lp.x * lp.x + lp.y * lp.y);
    const float inNorm = ip.w;
    const float ddx = ip.x - lp.x, ddy = ip.y - lp.y, ddz = ip.z - lp.z;
    const float delta = (dim == 3) ? sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) : sqrtf(ddx * ddx + ddy * ddy);
    const float d_max = prm.epsilonA * inNorm;
    // normal in the sensor frame (rotation part of Tinv)
    const float n0 = nrm[(long long)i * dim + 0], n1 = nrm[(long long)i * dim + 1], n2 = dim == 3 ? nrm[(long long)i * dim + 2] : 0.f;
    const float nsx = __fmaf_rn(Tinv.m[8], n2, __fmaf_rn(Tinv.m[4], n1, __fmul_rn(Tinv.m[0], n0)));
    const float nsy = __fmaf_rn(Tinv.m[9], n2, __fmaf_rn(Tinv.m[5], n1, __fmul_rn(Tinv.m[1], n0)));
    const float nsz = __fmaf_rn(Tinv.m[10], n2, __fmaf_rn(Tinv.m[6], n1, __fmul_rn(Tinv.m[2], n0)));
    const float dotn = nsx * (lp.x / lpNorm) + nsy * (lp.y / lpNorm) + nsz * (lp.z / lpNorm);
    const float w_v = (float)((double)eps + (1. - (double)eps) * fabs((double)dotn));
    const float w_d1 = (float)((double)eps + (1. - (double)eps) * (1. - (double)sqrtf(dist) / (double)(2 * prm.beamHalfAngle)));
    const float offset = delta - prm.epsilonD;
    float w_d2 = 1.f;
    if (delta < prm.epsilonD || lpNorm > inNorm) {
        w_d2 = eps;
    } else if (offset < d_max) {
        w_d2 = eps + (1 - eps) * offset / d_max;
    }
    float w_p2 = eps;
    if (delta < prm.epsilonD) {
        w_p2 = 1;
    } else if (offset < d_max) {
        w_p2 = (float)((double)eps + (1. - (double)eps) * (1. - (double)(offset / d_max)));
    }
    if ((inNorm + prm.epsilonD + d_max) >= lpNorm) {
        const float lastDyn = prob[i];
        const float c1 = (1 - (w_v * w_d1));
        const float c2 = w_v * w_d1;
        float probDynamic, probStatic;
        if (lastDyn < prm.thresholdDynamic) {
            probDynamic = c1 * lastDyn + c2 * w_d2 * ((1 - prm.alpha) * (1 - lastDyn) + prm.beta * lastDyn);
            probStatic = c1 * (1 - lastDyn) + c2 * w_p2 * (prm.alpha * (1 - lastDyn) + (1 - prm.beta) * lastDyn);
        } else {
            probDynamic = 1 - eps;
            probStatic = eps;
        }
        prob[i] = probDynamic / (probDynamic + probStatic);
    }
}

M16 to_m16(const float* T) {
    M16 m;
    memcpy(m.m, T, sizeof(m.m));
    return m;
}

}  // namespace

cudaError_t launch_dyn_input(const float* d_in, int rows, int dim, const float* Tinv16, int64_t n_in, float4* d_in_sensor, float* d_in_angles,
                             cudaStream_t s) {
    if (n_in <= 0) return cudaSuccess;
    dyn_input_kernel<<<blocks_for(n_in), 256, 0, s>>>(d_in, rows, dim, to_m16(Tinv16), (long long)n_in, d_in_sensor, d_in_angles);
    return cudaGetLastError();
}

cudaError_t launch_dyn_queries(const MapStore& m, int dim, const float* Tinv16, float sensor_max_range, float4* d_q4, cudaStream_t s) {
    if (m.n_active <= 0) return cudaSuccess;
    dyn_query_kernel<<<blocks_for(m.n_active), 256, 0, s>>>(m.feat, m.all_loaded ? nullptr : m.active, (long long)m.n_active, dim, to_m16(Tinv16),
                                                            sensor_max_range, d_q4);
    return cudaGetLastError();
}

cudaError_t launch_dyn_update(MapStore& m, int dim, const float* Tinv16, const DynParams& prm, const float4* d_in_sensor, const int32_t* d_ids,
                              const float* d_d2, cudaStream_t s) {
    if (m.n_active <= 0) return cudaSuccess;
    dyn_update_kernel<<<blocks_for(m.n_active), 256, 0, s>>>(m.feat, m.all_loaded ? nullptr : m.active, (long long)m.n_active, dim, to_m16(Tinv16),
                                                             prm, m.nrm, m.prob, d_in_sensor, d_ids, d_d2);
    return cudaGetLastError();
}

cudaError_t launch_normals(const GridView& g, int dim, int knn, const int32_t* d_nn_pos, const float* d_nn_d2, const uint32_t* d_list,
                           const unsigned int* d_n_list, long long list_capacity, float4* d_nrm_sorted, float* d_store_nrm, float* d_kth,
                           cudaStream_t s) {
    const long long n = d_list ? list_capacity : (long long)g.n;
    if (n <= 0) return cudaSuccess;
    normals_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, dim, knn, d_nn_pos, d_nn_d2, d_list, d_n_list, d_nrm_sorted, d_store_nrm, d_kth);
    return cudaGetLastError();
}

cudaError_t launch_normals_dirty(const GridView& new_points, const MapStore& m, const float* d_kth, int64_t n_old, uint8_t* d_dirty, cudaStream_t s) {
    if (m.n <= 0) return cudaSuccess;
    normals_dirty_kernel<<<blocks_for(m.n), 256, 0, s>>>(new_points, m.feat, m.loaded, m.touched, d_kth, (long long)n_old, (long long)m.n, d_dirty);
    return cudaGetLastError();
}

// flag[i] = 1 for the points that changed since the last pass: appended (i >= n_old) or flipped by the window
__global__ void __launch_bounds__(256) normals_changed_kernel(const uint8_t* __restrict__ touched, long long n_old, long long n_all, uint8_t* __restrict__ flag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_all) flag[i] = (i >= n_old || touched[i]) ? 1 : 0;
}
cudaError_t launch_normals_changed(const MapStore& m, int64_t n_old, uint8_t* d_flag, cudaStream_t s) {
    if (m.n <= 0) return cudaSuccess;
    normals_changed_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.touched, (long long)n_old, (long long)m.n, d_flag);
    return cudaGetLastError();
}
cudaError_t store_clear_touched(MapStore& m, cudaStream_t s) {
    m.nrm_touched = false;
    if (!m.touched || m.n <= 0) return cudaSuccess;
    return cudaMemsetAsync(m.touched, 0, (size_t)m.n, s);
}

cudaError_t launch_normals_positions(const GridView& g, const uint8_t* d_dirty, uint8_t* d_flag, cudaStream_t s) {
    if (g.n <= 0) return cudaSuccess;
    normals_flag_positions_kernel<<<blocks_for(g.n), 256, 0, s>>>(g, d_dirty, d_flag);
    return cudaGetLastError();
}

cudaError_t launch_normals_gather(const GridView& g, const uint32_t* d_list, const unsigned int* d_n_list, long long capacity, float4* d_q, cudaStream_t s,
                                  const float* d_kth, long long n_old) {
    if (capacity <= 0) return cudaSuccess;
    normals_gather_queries_kernel<<<blocks_for(capacity), 256, 0, s>>>(g, d_list, d_n_list, d_q, d_kth, n_old);
    return cudaGetLastError();
}

}  // namespace b200
