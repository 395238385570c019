// icp.cu -- everything in one ICP iteration after the correspondence search, plus the rigid
// transform kernels.
//
// Replaces, inside PM::ICPSequence::operator() (/root/reference/norlab_icp_mapper/Mapper.cpp:213,
// restated in SURVEY.md 3.2 / Appendix A):
//   outlierFilters.compute          -> exact radix-select of the distance quantile (3 histogram
//                                      passes over the float bit patterns) + threshold tests
//   ErrorElements + errorMinimizer  -> one fused gather/accumulate kernel: 27 point-to-plane sums
//                                      (21 upper-triangular entries of A, 6 of b) or 16
//                                      point-to-point sums, fp32 products accumulated in fp64 with
//                                      warp shuffles, fixed-order two-stage reduction (run-to-run
//                                      deterministic), then the last block solves the 6x6 / SVD,
//                                      composes T_iter and runs the transformation checkers.
//   transformations.apply           -> fused into the consumers (T_iter is applied on the fly).
// No host synchronisation happens inside the loop: the state lives in IcpState on the device.
#include <algorithm>

#include "icp_device.cuh"

namespace b200 {
namespace {

// ---- reading preparation: (dim+1) x N upload -> float4 in the refMean frame ---------------------
__global__ void __launch_bounds__(256) prep_reading_kernel(const float* __restrict__ in, int rows, int dim, Mat4 Tpre,
                                                           float4* __restrict__ out, GridView g, uint32_t* __restrict__ keys,
                                                           uint32_t* __restrict__ vals, long long nq, int coarse_shift,
                                                           unsigned int* __restrict__ pmax2_bits, const float* __restrict__ max_dist_desc,
                                                           int strict) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float n2 = 0.f;
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (i < nq) {
        float4 r;
        r.x = in[i * rows + 0];
        r.y = in[i * rows + 1];
        r.z = (dim == 3) ? in[i * rows + 2] : 0.f;
        r.w = 1.f;
        p = apply_T(Tpre.m, r);
        float w = 1.f;
        if (max_dist_desc) {  // the reading's `maxSearchDist` descriptor: this point's squared search radius rides in .w
            const float r = max_dist_desc[i];
            w = r * r;
            if (strict && w < CUDART_INF_F) w = __uint_as_float(__float_as_uint(w) - (w > 0.f ? 1u : 0u));  // '<' == '<=' the float below
        }
        out[i] = make_float4(p.x, p.y, p.z, w);
        n2 = p.x * p.x + p.y * p.y + p.z * p.z;
        if (!(n2 < 3.0e38f)) n2 = 0.f;  // NaN / inf points never pair: they do not count
    }
    if (pmax2_bits) {  // (whole warps reach this point: the early exit comes after it)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n2 = fmaxf(n2, __shfl_xor_sync(0xffffffffu, n2, o));
        if ((threadIdx.x & 31) == 0 && n2 > 0.f) atomicMax(pmax2_bits, __float_as_uint(n2));  // non-negative floats order like their bits
    }
    if (i >= nq) return;
    if (keys) {
        const float lim = 16777216.f;
        const int cx = min((int)floorf(fminf(fmaxf((p.x - g.ox) * g.inv_h, 0.f), lim)), g.nx - 1);
        const int cy = min((int)floorf(fminf(fmaxf((p.y - g.oy) * g.inv_h, 0.f), lim)), g.ny - 1);
        const int cz = min((int)floorf(fminf(fmaxf((p.z - g.oz) * g.inv_h, 0.f), lim)), g.nz - 1);
        if (coarse_shift > 0) {
            // locality is all the sort is for: blocks of 2^s x 2^s x 2^(s-1) cells need half the radix passes of the full cell id
            const int sz = max(coarse_shift - 1, 0);
            const uint32_t nxb = (uint32_t)((g.nx - 1) >> coarse_shift) + 1u, nyb = (uint32_t)((g.ny - 1) >> coarse_shift) + 1u;
            keys[i] = ((uint32_t)(cz >> sz) * nyb + (uint32_t)(cy >> coarse_shift)) * nxb + (uint32_t)(cx >> coarse_shift);
        } else {
            keys[i] = ((uint32_t)cz * (uint32_t)g.ny + (uint32_t)cy) * (uint32_t)g.nx + (uint32_t)cx;
        }
        vals[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(256) prep_normals_kernel(const float* __restrict__ in, int dim, Mat4 Tpre, float4* __restrict__ out, long long nq) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float x = in[i * dim + 0], y = in[i * dim + 1], z = (dim == 3) ? in[i * dim + 2] : 0.f;
    const float* T = Tpre.m;
    out[i] = make_float4(__fmaf_rn(T[8], z, __fmaf_rn(T[4], y, __fmul_rn(T[0], x))), __fmaf_rn(T[9], z, __fmaf_rn(T[5], y, __fmul_rn(T[1], x))),
                         __fmaf_rn(T[10], z, __fmaf_rn(T[6], y, __fmul_rn(T[2], x))), 0.f);
}

__global__ void __launch_bounds__(256) gather_reading_kernel(const float4* __restrict__ in, const uint32_t* __restrict__ perm,
                                                             float4* __restrict__ out, long long nq) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < nq) out[i] = in[perm[i]];
}

// ---- RigidTransformation::compute (Mapper.cpp:197,221; Map.cpp:523,525) --------------------------
__global__ void __launch_bounds__(256) transform_kernel(float* __restrict__ feat, int rows, int dim, float* __restrict__ normals,
                                                        long long n, Mat4 T) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r;
    r.x = feat[i * rows + 0];
    r.y = feat[i * rows + 1];
    r.z = (dim == 3) ? feat[i * rows + 2] : 0.f;
    r.w = 1.f;
    const float3 p = apply_T(T.m, r);
    feat[i * rows + 0] = p.x;
    feat[i * rows + 1] = p.y;
    if (dim == 3) feat[i * rows + 2] = p.z;
    if (normals) {
        float4 v;
        v.x = normals[i * dim + 0];
        v.y = normals[i * dim + 1];
        v.z = (dim == 3) ? normals[i * dim + 2] : 0.f;
        const float* M = T.m;
        const float ox = __fmaf_rn(M[8], v.z, __fmaf_rn(M[4], v.y, __fmul_rn(M[0], v.x)));
        const float oy = __fmaf_rn(M[9], v.z, __fmaf_rn(M[5], v.y, __fmul_rn(M[1], v.x)));
        const float oz = __fmaf_rn(M[10], v.z, __fmaf_rn(M[6], v.y, __fmul_rn(M[2], v.x)));
        normals[i * dim + 0] = ox;
        normals[i * dim + 1] = oy;
        if (dim == 3) normals[i * dim + 2] = oz;
    }
}

// descriptors that rotate with the cloud (`observationDirections`): same arithmetic as the normals above
__global__ void __launch_bounds__(256) rotate_rows_kernel(float* __restrict__ block, int stride, int offset, int dim, long long n, Mat4 T) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* v = block + i * stride + offset;
    const float vx = v[0], vy = v[1], vz = (dim == 3) ? v[2] : 0.f;
    const float* M = T.m;
    v[0] = __fmaf_rn(M[8], vz, __fmaf_rn(M[4], vy, __fmul_rn(M[0], vx)));
    v[1] = __fmaf_rn(M[9], vz, __fmaf_rn(M[5], vy, __fmul_rn(M[1], vx)));
    if (dim == 3) v[2] = __fmaf_rn(M[10], vz, __fmaf_rn(M[6], vy, __fmul_rn(M[2], vx)));
}

// ---- exact quantile of the finite squared distances (LPM Matches::getDistsQuantile) -------------
// One thread-block cluster of kSelCtas CTAs does the whole 3-pass radix select in ONE launch: each
// CTA keeps its slice of the distances in shared memory (read from L2 once), builds a shared-memory
// histogram per pass, merges it into a tiny global histogram, and the hardware cluster barrier
// separates the passes.  Every CTA then picks the bucket redundantly (2048 words from L2).
constexpr int kSelCtas = 8;
constexpr int kSelThreads = 512;
constexpr int kSelCache = 47 * 1024;  // floats cached per CTA (188 KB of dynamic shared memory)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Block-wide (kSelThreads threads): smallest bin with cumulative count > rank.
__device__ uint32_t select_pick(const uint32_t* hist, int nbins, uint32_t rank, bool rank_is_fraction, float q,
                                uint32_t* s_bin, uint32_t* s_res, uint32_t* s_warp /* kSelThreads / 32 + 1 */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = nbins / kSelThreads;  // 4 or 2
    uint32_t loc[4];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        loc[j] = (j < per) ? __ldcg(hist + tid * per + j) : 0u;
        sum += loc[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) {
        const uint32_t v = s_warp[w];
        if (w < warp) before += v;
        total += v;
    }
    incl += before;
    if (rank_is_fraction) {
        // idx = size_t(values.size() * quantile) evaluated in fp32; quantile == 1 -> max element
        rank = (q == 1.0f) ? (total ? total - 1u : 0u) : (uint32_t)((float)total * q);
        if (total && rank >= total) rank = total - 1u;
    }
    const uint32_t excl = incl - sum;
    if (total && rank >= excl && rank < incl) {
        uint32_t run = excl;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < per && rank >= run && rank < run + loc[j]) {
                *s_bin = (uint32_t)(tid * per + j);
                *s_res = rank - run;
            }
            run += loc[j];
        }
    }
    __syncthreads();
    return total;
}

__global__ void __cluster_dims__(kSelCtas, 1, 1) __launch_bounds__(kSelThreads)
    select_kernel(IcpState* __restrict__ st, const float* __restrict__ d2, int knn, uint32_t* __restrict__ hist, float quantile) {
    if (st->done) return;  // uniform over the cluster
    const bool stamper = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamper) B200_STAMP(st, 0);
    extern __shared__ uint4 s_cache4[];
    __shared__ uint32_t sh[kHistBins];
    __shared__ uint32_t s_scan[kSelThreads / 32 + 1];
    __shared__ uint32_t s_bin, s_res;
    // this CTA's slice, in units of 4 distances (16-byte loads; the buffer is padded past m)
    const int m = st->nq * knn;
    const int m4 = (m + 3) >> 2;
    const int chunk4 = (m4 + kSelCtas - 1) / kSelCtas;
    const int lo4 = min(m4, chunk4 * (int)blockIdx.x), hi4 = min(m4, lo4 + chunk4);
    const int n4 = hi4 - lo4;
    const bool cached = n4 * 4 <= kSelCache;
    const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(d2) + lo4;
    uint32_t prefix = 0, rank = 0, total = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int nbins = (pass == 2) ? 1024 : kHistBins;
        for (int i = threadIdx.x; i < nbins; i += kSelThreads) sh[i] = 0u;
        if (threadIdx.x == 0) {
            s_bin = 0;
            s_res = 0;
        }
        __syncthreads();
#pragma unroll 2
        for (int i = threadIdx.x; i < n4; i += kSelThreads) {
            uint4 v;
            if (pass == 0 || !cached) {
                v = __ldg(src4 + i);
                // entries past m (padding of the last group of 4) can never be selected
                const int e = (lo4 + i) * 4;
                if (e + 1 >= m) v.y = 0xffffffffu;
                if (e + 2 >= m) v.z = 0xffffffffu;
                if (e + 3 >= m) v.w = 0xffffffffu;
                if (cached) s_cache4[i] = v;
            } else {
                v = s_cache4[i];
            }
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t bits = w[c];
                if (pass == 0) {
                    if (bits < 0x7f800000u) atomicAdd(&sh[bits >> 21], 1u);  // finite only
                } else if (pass == 1) {
                    if ((bits >> 21) == prefix) atomicAdd(&sh[(bits >> 10) & 2047u], 1u);
                } else {
                    if ((bits >> 10) == prefix) atomicAdd(&sh[bits & 1023u], 1u);
                }
            }
        }
        __syncthreads();
        if (stamper) B200_STAMP(st, 1 + pass * 3);
        uint32_t* gh = hist + pass * kHistBins;
        for (int i = threadIdx.x; i < nbins; i += kSelThreads)
            if (sh[i]) atomicAdd(&gh[i], sh[i]);
        __threadfence();
        cluster_sync_all();
        if (stamper) B200_STAMP(st, 2 + pass * 3);
        const uint32_t tot = select_pick(gh, nbins, rank, pass == 0, quantile < 0.f ? st->dyn_quantile : quantile, &s_bin, &s_res, s_scan);
        if (pass == 0) total = tot;
        rank = s_res;
        prefix = (pass == 0) ? s_bin : ((prefix << (pass == 1 ? 11 : 10)) | s_bin);
        __syncthreads();
        if (stamper) B200_STAMP(st, 3 + pass * 3);
    }
    // everyone has read the histograms: clear them for the next iteration
    cluster_sync_all();
    for (int i = blockIdx.x * kSelThreads + threadIdx.x; i < 3 * kHistBins; i += kSelCtas * kSelThreads) hist[i] = 0u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (total == 0) {  // LPM: ConvergenceError("no outlier to filter")
            st->status = B200ICP_ERR_CONVERGENCE;
            st->done = 1;
        }
        st->limit = __uint_as_float(prefix);
        B200_STAMP(st, 10);
    }
}

template <int MIN>
__global__ void __launch_bounds__(256) accumulate_kernel(IcpParams prm, GridView g, const float4* __restrict__ nrm,
                                                         const float4* __restrict__ reading, const int32_t* __restrict__ mpos,
                                                         const float* __restrict__ md2, IcpState* __restrict__ st,
                                                         double* __restrict__ partials, float* __restrict__ trace) {
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) B200_STAMP(st, 11);
    constexpr int NS = SumLayout<MIN>::N;
    const int nq = st->nq;
    const int K = prm.knn;
    const long long m = (long long)nq * K;
    float T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = st->T[i];
    const float qlimit = st->limit;
    const float rscale = st->robust_scale;

    // per-thread sums stay in fp32 (a thread sees only a handful of pairs); everything above the
    // thread level is accumulated in fp64 in a fixed order
    float acc[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = 0.f;

    for (long long e = blockIdx.x * 256ll + threadIdx.x; e < m; e += (long long)gridDim.x * 256ll) {
        accumulate_entry<MIN>(acc, prm, T, g, nrm, reading, e, K, mpos[e], md2[e], qlimit, rscale);
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) B200_STAMP(st, 12);
    // warp shuffle reduction, then one partial per block
    __shared__ double s_part[8][NS];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        float v32[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v32[i] = (i < NS) ? acc[i] : 0.f;
        const float tot = warp_reduce_32slots(v32, lane);  // lane l now holds the warp total of slot l
        if (lane < NS) s_part[warp][lane] = (double)tot;
    }
    __syncthreads();
    if (threadIdx.x < kAccSlots) {
        double v = 0.0;
        if (threadIdx.x < NS) {
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) v += s_part[wv][threadIdx.x];
        }
        partials[(size_t)blockIdx.x * kAccSlots + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) B200_STAMP(st, 13);
    // fixed-order final reduction -> deterministic sums
    __shared__ double s_red[8][kAccSlots];
    __shared__ double s_sum[kAccSlots];
    {
        const int slot = threadIdx.x & 31, part = threadIdx.x >> 5;
        constexpr int PER = (kAccBlocks + 7) / 8;
        double tmp[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {  // all loads in flight together, then a fixed-order sum
            const unsigned b = part + 8u * j;
            tmp[j] = (b < gridDim.x) ? __ldcg(partials + (size_t)b * kAccSlots + slot) : 0.0;
        }
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < PER; ++j) v += tmp[j];
        s_red[part][slot] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
#pragma unroll
        for (int part = 0; part < 8; ++part) v += s_red[part][threadIdx.x];
        s_sum[threadIdx.x] = v;
    }
    __syncthreads();
    __shared__ float s_scratch[16];
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            st->ticket = 0u;
            B200_STAMP(st, 14);
        }
        finish_warp(prm, st, s_sum, NS, trace, s_scratch);
        if (threadIdx.x == 0) B200_STAMP(st, 15);
    }
}

}  // namespace

// ---- host launchers -------------------------------------------------------------------------------
cudaError_t launch_prep_reading(const float* d_in, int rows, int dim, const float* Tpre16, float4* d_out,
                                const GridView* g_for_keys, uint32_t* d_keys, uint32_t* d_vals, int64_t nq, cudaStream_t s, int coarse_shift,
                                unsigned int* d_pmax2_bits, const float* d_max_dist_desc, int strict) {
    if (nq <= 0) return cudaSuccess;
    Mat4 T;
    memcpy(T.m, Tpre16, sizeof(T.m));
    GridView g{};
    if (g_for_keys) g = *g_for_keys;
    prep_reading_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(d_in, rows, dim, T, d_out, g, g_for_keys ? d_keys : nullptr,
                                                                      d_vals, (long long)nq, coarse_shift, d_pmax2_bits, d_max_dist_desc, strict);
    return cudaGetLastError();
}

cudaError_t launch_prep_normals(const float* d_in, int dim, const float* Tpre16, float4* d_out, int64_t nq, cudaStream_t s) {
    if (nq <= 0) return cudaSuccess;
    Mat4 T;
    memcpy(T.m, Tpre16, sizeof(T.m));
    prep_normals_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(d_in, dim, T, d_out, (long long)nq);
    return cudaGetLastError();
}

cudaError_t launch_gather_reading(const float4* d_in, const uint32_t* d_perm, float4* d_out, int64_t nq, cudaStream_t s) {
    if (nq <= 0) return cudaSuccess;
    gather_reading_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(d_in, d_perm, d_out, (long long)nq);
    return cudaGetLastError();
}

cudaError_t launch_transform(float* d_feat, int rows, int dim, float* d_normals, int64_t n, const float* T16, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    Mat4 T;
    memcpy(T.m, T16, sizeof(T.m));
    transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_feat, rows, dim, d_normals, (long long)n, T);
    return cudaGetLastError();
}

cudaError_t launch_rotate_rows(float* d_block, int stride, int offset, int dim, int64_t n, const float* T16, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    Mat4 T;
    memcpy(T.m, T16, sizeof(T.m));
    rotate_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_block, stride, offset, dim, (long long)n, T);
    return cudaGetLastError();
}

cudaError_t icp_device_setup() {  // once per device (context creation)
    return cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelCache * sizeof(uint32_t)));
}

cudaError_t launch_iteration_tail(const IcpParams& p, const GridIndex& g, IcpBuffers& b, int it, cudaStream_t s, int* launches,
                                  cudaEvent_t ev_mid, VarTrimScratch* var_scratch) {
    const long long m = (long long)b.cap_nq * p.knn;
    for (int f = 0; f < p.n_outlier; ++f)
        if (p.outlier_kind[f] == B200ICP_OUTLIER_ROBUST) {
            if (!var_scratch) return cudaErrorInvalidValue;
            cudaError_t e = launch_robust_scale(*var_scratch, p, f, b, it, s, launches);
            if (e != cudaSuccess) return e;
        }
    if (p.quantile_filter >= 0 && p.outlier_kind[p.quantile_filter] == B200ICP_OUTLIER_VAR_TRIMMED_DIST) {
        if (!var_scratch) return cudaErrorInvalidValue;
        cudaError_t e = launch_var_trimmed_ratio(*var_scratch, p, p.quantile_filter, b, s, launches);
        if (e != cudaSuccess) return e;
    }
    if (p.quantile_filter >= 0) {
        const size_t dyn = (size_t)kSelCache * sizeof(uint32_t);
        select_kernel<<<kSelCtas, kSelThreads, dyn, s>>>(b.state, b.match_d2, p.knn, b.hist, p.quantile);
        *launches += 1;
    }
    if (ev_mid) cudaEventRecord(ev_mid, s);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((m + 511) / 512, kAccBlocks));
    const float4* nrm = g.has_normals ? g.normals : nullptr;
    if (p.minimizer == B200ICP_MIN_POINT_TO_PLANE)
        accumulate_kernel<0><<<blocks, 256, 0, s>>>(p, g.view, nrm, b.reading, b.match_pos, b.match_d2, b.state, b.partials, b.trace);
    else if (p.minimizer == B200ICP_MIN_POINT_TO_POINT)
        accumulate_kernel<1><<<blocks, 256, 0, s>>>(p, g.view, nrm, b.reading, b.match_pos, b.match_d2, b.state, b.partials, b.trace);
    else
        accumulate_kernel<2><<<blocks, 256, 0, s>>>(p, g.view, nrm, b.reading, b.match_pos, b.match_d2, b.state, b.partials, b.trace);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace b200
