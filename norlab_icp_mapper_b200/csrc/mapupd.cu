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
                                                        uint8_t* __restrict__ loaded, uint8_t* __restrict__ keep_out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    if (keep_out) keep_out[i] = (uint8_t)keep[i];
    if (!keep[i]) return;
    const long long dst = base + offs[i];
    store[dst] = make_float4(in[i * rows + 0], in[i * rows + 1], dim == 3 ? in[i * rows + 2] : 0.f, 1.f);
    loaded[dst] = 1;
    if (store_nrm && in_nrm)
        for (int c = 0; c < dim; ++c) store_nrm[dst * dim + c] = in_nrm[i * dim + c];
}

// unload: loaded points inside the slab's metric AABB leave the local cloud (Map.cpp:161-174);
// load: parked points whose 20 m grid coordinate lies in the slab come back (Map.cpp:79-99, cell ids
// are floor(x / CELL_SIZE), Map.cpp:232-235).
__global__ void __launch_bounds__(256) window_kernel(const float4* __restrict__ store, long long n, uint8_t* __restrict__ loaded,
                                                     int load, float cell, int r0, int r1, int c0, int c1, int a0, int a1,
                                                     unsigned long long* __restrict__ changed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = store[i];
    if (load) {
        if (loaded[i]) return;
        const int gx = (int)floorf(p.x / cell), gy = (int)floorf(p.y / cell), gz = (int)floorf(p.z / cell);
        if (gx >= r0 && gx <= r1 && gy >= c0 && gy <= c1 && gz >= a0 && gz <= a1) {
            loaded[i] = 1;
            atomicAdd(changed, 1ull);
        }
    } else {
        if (!loaded[i]) return;
        const float sx = (float)r0 * cell, ex = ((float)r1 + 1.f) * cell;
        const float sy = (float)c0 * cell, ey = ((float)c1 + 1.f) * cell;
        const float sz = (float)a0 * cell, ez = ((float)a1 + 1.f) * cell;
        if (p.x >= sx && p.x < ex && p.y >= sy && p.y < ey && p.z >= sz && p.z < ez) {
            loaded[i] = 0;
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
__global__ void __launch_bounds__(128) normals_kernel(GridView g, int dim, int knn, const int32_t* __restrict__ nn_pos,
                                                      float4* __restrict__ nrm_sorted, float* __restrict__ store_nrm) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= g.n) return;
    float mx = 0.f, my = 0.f, mz = 0.f;
    int cnt = 0;
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
}

template <typename T>
cudaError_t regrow(T*& p, int64_t old_count, int64_t new_cap, cudaStream_t s) {
    T* np = nullptr;
    cudaError_t e = cudaMalloc((void**)&np, (size_t)std::max<int64_t>(new_cap, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && old_count > 0) {
        e = cudaMemcpyAsync(np, p, (size_t)old_count * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
    }
    if (p) cudaFree(p);
    p = np;
    return cudaSuccess;
}

unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void store_free(MapStore& m) {
    cudaFree(m.feat);
    cudaFree(m.nrm);
    cudaFree(m.loaded);
    cudaFree(m.active);
    cudaFree(m.tmp_u32a);
    cudaFree(m.tmp_u32b);
    cudaFree(m.d_counter);
    m = MapStore{};
}

cudaError_t store_reserve(MapStore& m, int dim, int64_t n, cudaStream_t s) {
    cudaError_t e;
    if (n > m.cap) {
        const int64_t cap = n + n / 2 + 4096;
        if ((e = regrow(m.feat, m.n, cap, s)) != cudaSuccess) return e;
        if ((e = regrow(m.nrm, m.n * dim, cap * dim, s)) != cudaSuccess) return e;
        if ((e = regrow(m.loaded, m.n, cap, s)) != cudaSuccess) return e;
        cudaFree(m.active);
        m.active = nullptr;
        if ((e = cudaMalloc((void**)&m.active, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
        m.cap = cap;
    }
    if (!m.d_counter)
        if ((e = cudaMalloc((void**)&m.d_counter, 64)) != cudaSuccess) return e;
    return cudaSuccess;
}

static cudaError_t ensure_tmp(MapStore& m, int64_t n) {
    if (n <= m.cap_tmp) return cudaSuccess;
    cudaFree(m.tmp_u32a);
    cudaFree(m.tmp_u32b);
    m.tmp_u32a = m.tmp_u32b = nullptr;
    m.cap_tmp = 0;
    const int64_t cap = n + n / 2 + 4096;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&m.tmp_u32a, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&m.tmp_u32b, (size_t)cap * sizeof(uint32_t))) != cudaSuccess) return e;
    m.cap_tmp = cap;
    return cudaSuccess;
}

static cudaError_t exclusive_sum(GridIndex& scratch, const uint32_t* in, uint32_t* out, int64_t n, cudaStream_t s) {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n);
    if (need > scratch.cub_tmp_bytes) {
        cudaFree(scratch.cub_tmp);
        scratch.cub_tmp = nullptr;
        scratch.cub_tmp_bytes = 0;
        cudaError_t e = cudaMalloc(&scratch.cub_tmp, need + 256);
        if (e != cudaSuccess) return e;
        scratch.cub_tmp_bytes = need + 256;
    }
    size_t bytes = scratch.cub_tmp_bytes;
    return cub::DeviceScan::ExclusiveSum(scratch.cub_tmp, bytes, in, out, (int)n, s);
}

cudaError_t store_set(MapStore& m, const float* d_in, int rows, int dim, const float* d_normals, int64_t n, cudaStream_t s) {
    cudaError_t e;
    m.n = 0;  // nothing to preserve
    if ((e = store_reserve(m, dim, n, s)) != cudaSuccess) return e;
    to_store_kernel<<<blocks_for(n), 256, 0, s>>>(d_in, rows, dim, (long long)n, m.feat, m.loaded);
    m.has_normals = d_normals != nullptr;
    if (d_normals)
        if ((e = cudaMemcpyAsync(m.nrm, d_normals, (size_t)n * dim * sizeof(float), cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
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
    window_kernel<<<blocks_for(m.n), 256, 0, s>>>(m.feat, (long long)m.n, m.loaded, load, 20.0f, slab6[0], slab6[1], slab6[2], slab6[3],
                                                  slab6[4], slab6[5], m.d_counter);
    unsigned long long c = 0;
    if ((e = cudaMemcpyAsync(&c, m.d_counter, sizeof(c), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    *changed = (int64_t)c;
    return cudaGetLastError();
}

cudaError_t store_insert_point_distance(MapStore& m, GridIndex& scratch, const float* d_in, int rows, int dim, const float* d_in_nrm,
                                        int64_t n_in, const int32_t* d_nn_id, float min_dist, int64_t* n_kept, uint8_t* d_keep_out,
                                        cudaStream_t s) {
    cudaError_t e;
    *n_kept = 0;
    if (n_in == 0) return cudaSuccess;
    if ((e = ensure_tmp(m, n_in + 1)) != cudaSuccess) return e;
    pd_keep_kernel<<<blocks_for(n_in), 256, 0, s>>>(d_in, rows, dim, (long long)n_in, m.feat, d_nn_id, min_dist * min_dist, m.tmp_u32a);
    if ((e = cudaMemsetAsync(m.tmp_u32a + n_in, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
    if ((e = exclusive_sum(scratch, m.tmp_u32a, m.tmp_u32b, n_in + 1, s)) != cudaSuccess) return e;
    uint32_t total = 0;
    if ((e = cudaMemcpyAsync(&total, m.tmp_u32b + n_in, sizeof(uint32_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if ((e = store_reserve(m, dim, m.n + total, s)) != cudaSuccess) return e;
    // DataPoints::concatenate keeps only descriptors present in both clouds
    const bool keep_normals = m.has_normals && d_in_nrm != nullptr;
    pd_append_kernel<<<blocks_for(n_in), 256, 0, s>>>(d_in, rows, dim, d_in_nrm, (long long)n_in, m.tmp_u32a, m.tmp_u32b, (long long)m.n,
                                                      m.feat, keep_normals ? m.nrm : nullptr, m.loaded, d_keep_out);
    m.has_normals = keep_normals;
    m.n += total;
    m.n_active += total;
    *n_kept = total;
    return cudaGetLastError();
}

cudaError_t launch_normals(const GridView& g, int dim, int knn, const int32_t* d_nn_pos, float4* d_nrm_sorted, float* d_store_nrm,
                           cudaStream_t s) {
    if (g.n <= 0) return cudaSuccess;
    normals_kernel<<<(unsigned)((g.n + 127) / 128), 128, 0, s>>>(g, dim, knn, d_nn_pos, d_nrm_sorted, d_store_nrm);
    return cudaGetLastError();
}

}  // namespace b200
