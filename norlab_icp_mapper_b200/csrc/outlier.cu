// outlier.cu -- VarTrimmedDistOutlierFilter::optimizeInlierRatio on the device (libpointmatcher
// OutlierFiltersImpl.cpp, reached through the YAML `icp.outlierFilters` of
// /root/reference/norlab_icp_mapper/Mapper.cpp:72 -> PM::ICPSequence::loadFromYamlNode).
//
// Per iteration: the finite, positive squared distances sorted ascending (radix sort), their running sum
// (fp64 scan), then the index in [floor(minRatio n), floor(maxRatio n)) that minimises
//   FRMS_i = (1 / (id_i / n)^lambda)^2 * cumsum_i / id_i,   id_i = i + 1,  n = knn * Nq,
// and ratio = argmin / n.  The ratio lands in IcpState::dyn_quantile; the quantile select and the
// accumulate kernel then treat the filter like TrimmedDist with that ratio.  Everything stays on the
// device (no host round trip): entries that do not qualify are sorted as +inf and the search range is
// clipped to the number that do.  Upstream accumulates the running sum sequentially in fp32; the fp64 scan
// can move the argmin by a few indices where FRMS is flat (tests bound the effect on the ratio and the pose).
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace b200 {
namespace {

__global__ void __launch_bounds__(256) var_keys_kernel(const IcpState* __restrict__ st, const float* __restrict__ d2, int knn, long long cap,
                                                       float* __restrict__ keys, unsigned int* __restrict__ count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long m = (long long)st->nq * knn;
    float v = CUDART_INF_F;
    if (i < m && i < cap) {
        const float d = d2[i];
        if (d != CUDART_INF_F && d > 0.f) v = d;
    }
    if (i < cap) keys[i] = v;
    const unsigned bal = __ballot_sync(0xffffffffu, v != CUDART_INF_F);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, (unsigned)__popc(bal));
}

struct FiniteToDouble {
    __host__ __device__ __forceinline__ double operator()(const float& v) const { return (v > 3.0e38f) ? 0.0 : (double)v; }
};

__global__ void __launch_bounds__(1024) var_argmin_kernel(IcpState* __restrict__ st, const double* __restrict__ cums, int knn,
                                                          unsigned int* __restrict__ count, float min_ratio, float max_ratio, float lambda) {
    __shared__ double s_val[1024];
    __shared__ long long s_idx[1024];
    const long long m = (long long)st->nq * knn;
    const long long cnt = (long long)*count;
    long long min_el = (long long)floorf(min_ratio * (float)m), max_el = (long long)floorf(max_ratio * (float)m);
    if (max_el > cnt) max_el = cnt;
    if (min_el >= max_el) min_el = max_el > 0 ? max_el - 1 : 0;
    double best = CUDART_INF;
    long long best_i = min_el;
    for (long long i = min_el + threadIdx.x; i < max_el; i += 1024) {
        const double id = (double)(i + 1);
        const double inv = 1.0 / pow(id / (double)m, (double)lambda);
        const double frms = inv * inv * cums[i] / id;
        if (frms < best) {  // ascending i per thread: the first minimum is kept
            best = frms;
            best_i = i;
        }
    }
    s_val[threadIdx.x] = best;
    s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            const double ov = s_val[threadIdx.x + off];
            const long long oi = s_idx[threadIdx.x + off];
            if (ov < s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) {
                s_val[threadIdx.x] = ov;
                s_idx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (cnt == 0) {  // LPM: ConvergenceError("Inlier ratio optimization failed: no finite distances")
            st->status = B200ICP_ERR_CONVERGENCE;
            st->done = 1;
        }
        st->dyn_quantile = (float)s_idx[0] / (float)m;
        *count = 0u;  // ready for the next iteration
    }
}

// ---- RobustOutlierFilter: the scale estimate (LPM OutlierFiltersImpl.cpp RobustOutlierFilter::robustFiltering) ----
// mad:  scale = sqrt(median |d - median d|) over the finite squared distances d (Matches::getMedianAbsDeviation: both medians
//       are the element at index n / 2 of the sorted values);
// berg: first iteration 1.9 sqrt(median d) (Matches::getDistsQuantile(0.5)), afterwards scale <- 0.85 (scale - target) + target;
// std:  scale = sqrt(std d) over ALL entries, n - 1 in the denominator (Matches::getStandardDeviation; an unmatched entry's
//       +inf makes it NaN, as upstream).
__global__ void __launch_bounds__(256) robust_keys_kernel(const IcpState* __restrict__ st, const float* __restrict__ d2, int knn, long long cap,
                                                          float* __restrict__ keys, unsigned int* __restrict__ count) {
    if (st->done) return;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long m = (long long)st->nq * knn;
    float v = CUDART_INF_F;
    if (i < m && i < cap) v = d2[i];
    if (i < cap) keys[i] = v;
    const unsigned bal = __ballot_sync(0xffffffffu, v != CUDART_INF_F);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(count, (unsigned)__popc(bal));
}

// stage 0: median of the sorted finite distances -> scratch[0]; berg's first iteration ends here.
// stage 1: median of the sorted absolute deviations -> the mad scale.
__global__ void robust_median_kernel(IcpState* __restrict__ st, const float* __restrict__ sorted, unsigned int* __restrict__ count,
                                     float* __restrict__ scratch, int stage, int estimator) {
    if (st->done) return;
    const unsigned n = *count;
    if (n == 0) {  // LPM: ConvergenceError("no outlier to filter")
        st->status = B200ICP_ERR_CONVERGENCE;
        st->done = 1;
        return;
    }
    const float med = sorted[n / 2];
    if (stage == 0) {
        scratch[0] = med;
        if (estimator == B200ICP_SCALE_BERG) {
            st->robust_scale = 1.9f * sqrtf(med);
            *count = 0u;
        }
    } else {
        st->robust_scale = sqrtf(med);
        *count = 0u;  // ready for the next iteration
    }
}

__global__ void __launch_bounds__(256) robust_absdev_kernel(const IcpState* __restrict__ st, const float* __restrict__ sorted,
                                                            const unsigned int* __restrict__ count, const float* __restrict__ scratch,
                                                            float* __restrict__ keys, long long cap) {
    if (st->done) return;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= cap) return;
    keys[i] = (i < (long long)*count) ? fabsf(sorted[i] - scratch[0]) : CUDART_INF_F;
}

__global__ void robust_berg_step_kernel(IcpState* __restrict__ st, float target) {
    if (st->done) return;
    st->robust_scale = 0.85f * (st->robust_scale - target) + target;
}

__global__ void __launch_bounds__(1024) robust_std_kernel(IcpState* __restrict__ st, const float* __restrict__ d2, int knn) {
    if (st->done) return;
    __shared__ double s_red[32];
    __shared__ double s_mean;
    const long long m = (long long)st->nq * knn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int pass = 0; pass < 2; ++pass) {
        double acc = 0.0;
        const double mean = pass ? s_mean : 0.0;
        for (long long i = threadIdx.x; i < m; i += 1024) {
            const double v = (double)d2[i];
            acc += pass ? (v - mean) * (v - mean) : v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < 32; ++w) tot += s_red[w];
            if (pass == 0) s_mean = tot / (double)m;
            else st->robust_scale = sqrtf((float)sqrt(tot / (double)(m - 1)));
        }
        __syncthreads();
    }
}

}  // namespace

void var_trimmed_free(VarTrimScratch& v) {
    B200_CUDA_FREE(v.keys_in);
    B200_CUDA_FREE(v.keys_out);
    B200_CUDA_FREE(v.cums);
    B200_CUDA_FREE(v.cub_tmp);
    B200_CUDA_FREE(v.d_count);
    v = VarTrimScratch{};
}

static cudaError_t ensure_sort_scratch(VarTrimScratch& v, long long cap, cudaStream_t s);

cudaError_t launch_robust_scale(VarTrimScratch& v, const IcpParams& p, int f, IcpBuffers& b, int it, cudaStream_t s, int* launches) {
    const int mode = p.outlier_mode[f];
    const int estimator = (mode >> 8) & 15, nb = (mode >> 16) & 0x7fff;
    const int iteration = it + 1;  // LPM counts from 1; here it restarts with every registration
    if (estimator == B200ICP_SCALE_NONE || !(iteration <= nb || nb == 0)) return cudaSuccess;
    if (estimator == B200ICP_SCALE_STD) {
        robust_std_kernel<<<1, 1024, 0, s>>>(b.state, b.match_d2, p.knn);
        *launches += 1;
        return cudaGetLastError();
    }
    if (estimator == B200ICP_SCALE_BERG && iteration > 1) {
        robust_berg_step_kernel<<<1, 1, 0, s>>>(b.state, p.outlier_param[f]);
        *launches += 1;
        return cudaGetLastError();
    }
    const long long cap = (long long)b.cap_nq * p.knn;
    cudaError_t e = ensure_sort_scratch(v, cap, s);
    if (e != cudaSuccess) return e;
    float* scratch = reinterpret_cast<float*>(v.d_count) + 4;
    robust_keys_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(b.state, b.match_d2, p.knn, cap, v.keys_in, v.d_count);
    size_t bytes = v.cub_bytes;
    if ((e = cub::DeviceRadixSort::SortKeys(v.cub_tmp, bytes, v.keys_in, v.keys_out, (int)cap, 0, 32, s)) != cudaSuccess) return e;
    robust_median_kernel<<<1, 1, 0, s>>>(b.state, v.keys_out, v.d_count, scratch, 0, estimator);
    *launches += 4;
    if (estimator == B200ICP_SCALE_MAD) {
        robust_absdev_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(b.state, v.keys_out, v.d_count, scratch, v.keys_in, cap);
        bytes = v.cub_bytes;
        if ((e = cub::DeviceRadixSort::SortKeys(v.cub_tmp, bytes, v.keys_in, v.keys_out, (int)cap, 0, 32, s)) != cudaSuccess) return e;
        robust_median_kernel<<<1, 1, 0, s>>>(b.state, v.keys_out, v.d_count, scratch, 1, estimator);
        *launches += 4;
    }
    return cudaGetLastError();
}

static cudaError_t ensure_sort_scratch(VarTrimScratch& v, long long cap, cudaStream_t s) {
    cudaError_t e;
    if (cap > v.cap) {
        var_trimmed_free(v);
        if ((e = B200_CUDA_MALLOC((void**)&v.keys_in, (size_t)cap * sizeof(float))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&v.keys_out, (size_t)cap * sizeof(float))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&v.cums, (size_t)cap * sizeof(double))) != cudaSuccess) return e;
        if ((e = B200_CUDA_MALLOC((void**)&v.d_count, 64)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(v.d_count, 0, 64, s)) != cudaSuccess) return e;
        size_t need_sort = 0, need_scan = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, need_sort, v.keys_in, v.keys_out, (int)cap, 0, 32);
        auto it = thrust::make_transform_iterator((const float*)v.keys_out, FiniteToDouble());
        cub::DeviceScan::InclusiveSum(nullptr, need_scan, it, v.cums, (int)cap);
        v.cub_bytes = (need_sort > need_scan ? need_sort : need_scan) + 256;
        if ((e = B200_CUDA_MALLOC(&v.cub_tmp, v.cub_bytes)) != cudaSuccess) return e;
        v.cap = cap;
    }
    return cudaSuccess;
}

cudaError_t launch_var_trimmed_ratio(VarTrimScratch& v, const IcpParams& p, int f, IcpBuffers& b, cudaStream_t s, int* launches) {
    const long long cap = (long long)b.cap_nq * p.knn;
    cudaError_t e = ensure_sort_scratch(v, cap, s);
    if (e != cudaSuccess) return e;
    var_keys_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(b.state, b.match_d2, p.knn, cap, v.keys_in, v.d_count);
    size_t bytes = v.cub_bytes;
    if ((e = cub::DeviceRadixSort::SortKeys(v.cub_tmp, bytes, v.keys_in, v.keys_out, (int)cap, 0, 32, s)) != cudaSuccess) return e;
    auto it = thrust::make_transform_iterator((const float*)v.keys_out, FiniteToDouble());
    bytes = v.cub_bytes;
    if ((e = cub::DeviceScan::InclusiveSum(v.cub_tmp, bytes, it, v.cums, (int)cap, s)) != cudaSuccess) return e;
    var_argmin_kernel<<<1, 1024, 0, s>>>(b.state, v.cums, p.knn, v.d_count, p.outlier_param[f], p.outlier_param2[f], p.outlier_param3[f]);
    *launches += 6;
    return cudaGetLastError();
}

}  // namespace b200
