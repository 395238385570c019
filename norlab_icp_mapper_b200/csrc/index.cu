// index.cu -- GPU-resident spatial index that replaces libnabo's kd-tree.
//
// Reference behaviour replaced: KDTreeMatcher::init under PM::ICPSequence::setMap
// (/root/reference/norlab_icp_mapper/Map.cpp:111,178,528,581) and the throw-away
// Nabo::NNS::create in PointDistanceMapperModule.cpp:33-34.  Instead of an O(N log N)
// single-threaded tree build the cloud is binned into a dense uniform grid and radix-sorted by
// linear cell id (x fastest): one streaming pass for bounds + mean, one for keys, one stable sort,
// one gather.  A (y, z) row of cells is then a contiguous run of float4 points, which is what the
// search kernel (knn.cu) reads with coalesced 128-byte requests.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace b200 {

namespace {

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float ordered_to_float(unsigned int u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// out[0..2]: exact fixed-point coordinate sums (2^-16 m units, two's complement);
// out[3..5] / out[6..8]: min / max per axis in order-preserving uint encoding.
__global__ void __launch_bounds__(256) bbox_sum_kernel(const float* __restrict__ feat_all, int rows, int dim, long long n,
                                                       const uint32_t* __restrict__ subset,
                                                       unsigned long long* __restrict__ out) {
    long long sum[3] = {0, 0, 0};
    unsigned int mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* feat = feat_all + (subset ? (long long)subset[i] : i) * rows;
        for (int d = 0; d < dim; ++d) {
            const float v = feat[d];
            sum[d] += __double2ll_rn((double)v * 65536.0);
            const unsigned int o = float_to_ordered(v);
            mn[d] = min(mn[d], o);
            mx[d] = max(mx[d], o);
        }
    }
    for (int d = 0; d < 3; ++d) {
        for (int off = 16; off > 0; off >>= 1) {
            sum[d] += __shfl_down_sync(0xffffffffu, sum[d], off);
            mn[d] = min(mn[d], __shfl_down_sync(0xffffffffu, mn[d], off));
            mx[d] = max(mx[d], __shfl_down_sync(0xffffffffu, mx[d], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int d = 0; d < dim; ++d) {
            atomicAdd(&out[d], (unsigned long long)sum[d]);
            atomicMin((unsigned int*)&out[3 + d], mn[d]);
            atomicMax((unsigned int*)&out[6 + d], mx[d]);
        }
    }
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv_h, int n) {
    const float u = (x - o) * inv_h;
    int c = (int)floorf(fminf(fmaxf(u, 0.f), 16777216.f));
    return min(c, n - 1);
}

__global__ void __launch_bounds__(256) cell_key_kernel(const float* __restrict__ feat_all, int rows, int dim, long long n,
                                                       const uint32_t* __restrict__ subset, float mx, float my, float mz, GridView g,
                                                       float4* __restrict__ tmp_pts, uint32_t* __restrict__ keys,
                                                       uint32_t* __restrict__ vals, uint32_t* __restrict__ cell_count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long src = subset ? (long long)subset[i] : i;
    const float* feat = feat_all + src * rows;
    const float x = feat[0] - mx;
    const float y = feat[1] - my;
    const float z = (dim == 3) ? feat[2] - mz : 0.f;
    tmp_pts[i] = make_float4(x, y, z, __int_as_float((int)src));  // .w = index into the caller's cloud
    const int cx = cell_coord(x, g.ox, g.inv_h, g.nx);
    const int cy = cell_coord(y, g.oy, g.inv_h, g.ny);
    const int cz = cell_coord(z, g.oz, g.inv_h, g.nz);
    const uint32_t key = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
    keys[i] = key;
    vals[i] = (uint32_t)i;
    atomicAdd(&cell_count[key], 1u);
}

__global__ void __launch_bounds__(256) gather_sorted_kernel(const float4* __restrict__ tmp_pts, const float* __restrict__ normals,
                                                            int dim, const uint32_t* __restrict__ perm, long long n,
                                                            float4* __restrict__ pts, float4* __restrict__ nrm) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 p = tmp_pts[perm[j]];
    pts[j] = p;
    if (nrm) {
        const long long src = __float_as_int(p.w);
        const float nx = normals[src * dim + 0];
        const float ny = normals[src * dim + 1];
        const float nz = (dim == 3) ? normals[src * dim + 2] : 0.f;
        nrm[j] = make_float4(nx, ny, nz, 0.f);
    }
}

template <typename T>
cudaError_t ensure(T*& p, int64_t count) {
    if (p) B200_CUDA_FREE(p);
    p = nullptr;
    return B200_CUDA_MALLOC((void**)&p, (size_t)std::max<int64_t>(count, 1) * sizeof(T));
}

}  // namespace

void grid_free(GridIndex& g) {
    B200_CUDA_FREE(g.pts);
    B200_CUDA_FREE(g.normals);
    B200_CUDA_FREE(g.cell_start);
    B200_CUDA_FREE(g.tmp_pts);
    B200_CUDA_FREE(g.keys_in);
    B200_CUDA_FREE(g.keys_out);
    B200_CUDA_FREE(g.vals_in);
    B200_CUDA_FREE(g.vals_out);
    B200_CUDA_FREE(g.cub_tmp);
    B200_CUDA_FREE(g.d_reduce);
    g = GridIndex{};
}

cudaError_t ensure_scratch(GridIndex& g, int64_t n) {
    cudaError_t e;
    if (n > g.cap_scratch) {
        const int64_t cap = grow_capacity(n);
        if ((e = ensure(g.tmp_pts, cap)) != cudaSuccess) return e;
        if ((e = ensure(g.keys_in, cap)) != cudaSuccess) return e;
        if ((e = ensure(g.keys_out, cap)) != cudaSuccess) return e;
        if ((e = ensure(g.vals_in, cap)) != cudaSuccess) return e;
        if ((e = ensure(g.vals_out, cap)) != cudaSuccess) return e;
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (uint32_t*)nullptr, (int)std::min<int64_t>(cap, INT32_MAX), 0, 32);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, kMaxCells + 1);
        bytes = std::max(bytes, scan_bytes) + 256;
        if (bytes > g.cub_tmp_bytes) {
            if (g.cub_tmp) B200_CUDA_FREE(g.cub_tmp);
            g.cub_tmp = nullptr;
            if ((e = B200_CUDA_MALLOC(&g.cub_tmp, bytes)) != cudaSuccess) return e;
            g.cub_tmp_bytes = bytes;
        }
        g.cap_scratch = cap;
    }
    if (!g.d_reduce)
        if ((e = B200_CUDA_MALLOC((void**)&g.d_reduce, 16 * sizeof(unsigned long long))) != cudaSuccess) return e;
    return cudaSuccess;
}

// min / max per axis of a (subset of a) cloud; one streaming pass + a 72-byte read-back
cudaError_t cloud_bounds(GridIndex& g, const float* d_feat, int rows, int dim, int64_t n, const uint32_t* d_subset, float* lo3, float* hi3,
                         cudaStream_t s) {
    cudaError_t e = ensure_scratch(g, 1);
    if (e != cudaSuccess) return e;
    unsigned long long h_red[9] = {0, 0, 0, 0xffffffffull, 0xffffffffull, 0xffffffffull, 0, 0, 0};
    if ((e = cudaMemcpyAsync(g.d_reduce, h_red, sizeof(h_red), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4 * kSMs);
    bbox_sum_kernel<<<blocks, 256, 0, s>>>(d_feat, rows, dim, (long long)n, d_subset, g.d_reduce);
    if ((e = cudaMemcpyAsync(h_red, g.d_reduce, sizeof(h_red), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    for (int d = 0; d < 3; ++d) {
        lo3[d] = d < dim ? ordered_to_float((unsigned int)h_red[3 + d]) : 0.f;
        hi3[d] = d < dim ? ordered_to_float((unsigned int)h_red[6 + d]) : 0.f;
    }
    return cudaGetLastError();
}

cudaError_t sort_pairs(GridIndex& g, uint32_t* keys_in, uint32_t* keys_out, uint32_t* vals_in, uint32_t* vals_out,
                       int64_t n, int end_bit, cudaStream_t s) {
    cudaError_t e = ensure_scratch(g, n);
    if (e != cudaSuccess) return e;
    size_t bytes = g.cub_tmp_bytes;
    return cub::DeviceRadixSort::SortPairs(g.cub_tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                           std::max(1, std::min(32, end_bit)), s);
}

static int bits_for(uint64_t v) {
    int b = 1;
    while (b < 32 && (1ull << b) < v) ++b;
    return b;
}

cudaError_t grid_build(GridIndex& g, const float* d_feat, int rows, int dim, const float* d_normals, int64_t n,
                       bool centre, float cell_hint, cudaStream_t s, const uint32_t* d_subset) {
    cudaError_t e;
    if ((e = ensure_scratch(g, n)) != cudaSuccess) return e;

    // ---- pass 1: bounds + exact fixed-point sums -------------------------------------------
    unsigned long long h_red[9] = {0, 0, 0, 0xffffffffull, 0xffffffffull, 0xffffffffull, 0, 0, 0};
    if ((e = cudaMemcpyAsync(g.d_reduce, h_red, sizeof(h_red), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    {
        const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4 * kSMs);
        bbox_sum_kernel<<<blocks, 256, 0, s>>>(d_feat, rows, dim, (long long)n, d_subset, g.d_reduce);
    }
    if ((e = cudaMemcpyAsync(h_red, g.d_reduce, sizeof(h_red), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;

    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) {
        g.mean[d] = 0.f;
        if (d < dim) {
            if (centre) g.mean[d] = (float)((((double)(long long)h_red[d]) / 65536.0) / (double)n);
            lo[d] = ordered_to_float((unsigned int)h_red[3 + d]) - g.mean[d];
            hi[d] = ordered_to_float((unsigned int)h_red[6 + d]) - g.mean[d];
        }
    }

    // ---- choose the cell edge -------------------------------------------------------------
    // Dense table of at most kMaxCells cells; by default ~4 cells per point (surface-like clouds
    // occupy a few percent of them, giving O(10) points per occupied cell).
    double cells_per_point = 4.0;
    if (const char* env = getenv("B200ICP_CELLS_PER_POINT")) cells_per_point = std::max(0.01, atof(env));
    const double target = std::min<double>((double)kMaxCells, std::max<double>(64.0, cells_per_point * (double)n));
    double ext[3];
    for (int d = 0; d < 3; ++d) ext[d] = std::max(0.0, (double)hi[d] - (double)lo[d]);
    const double max_ext = std::max(ext[0], std::max(ext[1], ext[2]));
    auto count_cells = [&](double hh) {
        double c = 1.0;
        for (int d = 0; d < dim; ++d) c *= std::floor(ext[d] / hh) + 1.0;
        return c;
    };
    double hh;
    if (cell_hint > 0.f) {
        hh = cell_hint;
    } else if (max_ext <= 0.0) {
        hh = 1.0;
    } else {
        double a = max_ext * 1e-7, b = max_ext * 2.0;  // count_cells(b) == 1
        for (int it = 0; it < 80; ++it) {
            const double m = std::sqrt(a * b);
            if (count_cells(m) > target) a = m; else b = m;
        }
        hh = b;
    }
    while (count_cells(hh) > (double)kMaxCells) hh *= 1.1;
    GridView& v = g.view;
    v.h = (float)hh;
    v.inv_h = 1.0f / v.h;
    v.ox = lo[0];
    v.oy = lo[1];
    v.oz = lo[2];
    int nn[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        const float u = (hi[d] - lo[d]) * v.inv_h;  // same fp32 expression as cell_coord()
        nn[d] = (int)std::floor(u) + 1;
    }
    v.nx = nn[0];
    v.ny = nn[1];
    v.nz = nn[2];
    v.n = (int)n;
    const int max_n = std::max(v.nx, std::max(v.ny, v.nz));
    v.slack = 1e-6f * (float)(max_n + 8) + 1e-5f;
    const int64_t n_cells = (int64_t)v.nx * v.ny * v.nz;
    if (n_cells + 1 > g.cap_cells) {
        const int64_t cap_c = std::min<int64_t>(grow_capacity(n_cells + 1), (int64_t)kMaxCells + 2);
        if ((e = ensure(g.cell_start, cap_c)) != cudaSuccess) return e;
        g.cap_cells = cap_c;
    }
    if (n > g.cap_pts) {
        const int64_t cap = grow_capacity(n);
        if ((e = ensure(g.pts, cap)) != cudaSuccess) return e;
        g.cap_pts = cap;
    }
    if (d_normals && n > g.cap_normals) {
        const int64_t cap = grow_capacity(n);
        if ((e = ensure(g.normals, cap)) != cudaSuccess) return e;
        g.cap_normals = cap;
    }
    g.has_normals = d_normals != nullptr;

    // ---- pass 2: keys + per-cell counts; scan; stable sort; gather ---------------------------
    if ((e = cudaMemsetAsync(g.cell_start, 0, (size_t)(n_cells + 1) * sizeof(uint32_t), s)) != cudaSuccess) return e;
    const int blocks = (int)((n + 255) / 256);
    cell_key_kernel<<<blocks, 256, 0, s>>>(d_feat, rows, dim, (long long)n, d_subset, g.mean[0], g.mean[1], g.mean[2], v, g.tmp_pts,
                                           g.keys_in, g.vals_in, g.cell_start);
    size_t bytes = g.cub_tmp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(g.cub_tmp, bytes, g.cell_start, g.cell_start, (int)(n_cells + 1), s)) != cudaSuccess)
        return e;
    if ((e = sort_pairs(g, g.keys_in, g.keys_out, g.vals_in, g.vals_out, n, bits_for((uint64_t)n_cells), s)) != cudaSuccess)
        return e;
    gather_sorted_kernel<<<blocks, 256, 0, s>>>(g.tmp_pts, d_normals, dim, g.vals_out, (long long)n, g.pts,
                                                d_normals ? g.normals : nullptr);
    v.pts = g.pts;
    v.cell_start = g.cell_start;
    return cudaGetLastError();
}

}  // namespace b200
