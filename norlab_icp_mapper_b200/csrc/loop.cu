// loop.cu -- the whole ICP loop of one registration as ONE persistent cooperative kernel (k = 1).
//
// Same steps as the kernel-per-step path (knn.cu / icp.cu) -- warm ball search, exact radix-select
// of the distance quantile, gather + error sums, fixed-order reduction, one-warp solve, checkers --
// but the grid stays resident (one 1024-thread CTA per SM) and the steps are separated by a
// device-wide barrier (one atomic + acquire spin, ~1 us) instead of a kernel boundary (~4 us of
// launch gap each on B200).  Each CTA owns a fixed slice of the reading for the whole registration;
// every CTA reduces the per-CTA partial sums in the same fixed order and runs the 6x6 solve and the
// checkers redundantly, so T_iter and the stop decision are bit-identical everywhere and nothing has
// to be broadcast.  (The slices differ from the kernel-per-step path's, so the error sums -- and
// therefore the pose -- agree with that path to rounding, not bit for bit; each path is run-to-run
// deterministic.)
//
// Replaces the body of PM::ICPSequence::operator() (/root/reference/norlab_icp_mapper/Mapper.cpp:213)
// from the second iteration on; iteration 0's cold search is the stand-alone knn_kernel.
#include <cooperative_groups.h>

#include "icp_device.cuh"
#include "knn_device.cuh"

namespace b200 {
namespace {

constexpr int kLoopThreads = 1024;
constexpr int kLoopWarps = kLoopThreads / 32;
// Exact quantile inside the loop kernel: level 0 = 12 bits [30:19] of the float pattern (4096 bins,
// histogrammed while searching, merged through global memory); the bucket that holds the quantile
// then contains ~1 % of the distances, which every CTA pulls into shared memory as a candidate list
// and finishes locally (10 + 9 bits).  Falls back to two more global passes if the bucket is larger
// than the list.
constexpr int kSel0Bins = 4096;
constexpr int kSel0Shift = 19;
constexpr int kSelListCap = 4096;
// layout of the `hist` buffer (uint32): [0, 4096) level 0 | [4096, 6144) fallback level 1 | [6144, 6400) fallback level 2 |
// [6400] list counter | [8192, 8192 + kSelListCap) candidate list
constexpr int kHistL1 = 4096, kHistL2 = 6144, kHistCount = 6400, kHistList = 8192;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Monotonic counter barrier: the k-th barrier completes when the counter reaches k * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = epoch * gridDim.x;
        while (ld_acquire_u32(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// Block-wide (kLoopThreads threads): smallest bin with cumulative count > rank (see icp.cu select_pick).
// kShared: the histogram lives in shared memory (plain loads) instead of global memory (L2 loads).
template <bool kShared>
__device__ uint32_t loop_pick(const uint32_t* hist, int nbins, uint32_t rank, bool rank_is_fraction, float q, uint32_t* s_bin,
                              uint32_t* s_res, uint32_t* s_cnt, uint32_t* s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nbins + kLoopThreads - 1) / kLoopThreads;  // 1..4
    uint32_t loc[4];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int bin = tid * per + j;
        loc[j] = (j < per && bin < nbins) ? (kShared ? hist[bin] : __ldcg(hist + bin)) : 0u;
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
    for (int w = 0; w < kLoopWarps; ++w) {
        const uint32_t v = s_warp[w];
        if (w < warp) before += v;
        total += v;
    }
    incl += before;
    if (rank_is_fraction) {
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
                *s_cnt = loc[j];
            }
            run += loc[j];
        }
    }
    __syncthreads();
    return total;
}

template <int MIN, int kLoopG /* lanes per query in the warm search */>
__global__ void __launch_bounds__(kLoopThreads, 1)
    icp_loop_kernel(IcpParams prm, GridView g, const float4* __restrict__ nrm, const float4* __restrict__ reading,
                    int32_t* __restrict__ mpos, float* __restrict__ md2, IcpState* __restrict__ gst, uint32_t* __restrict__ hist,
                    double* __restrict__ partials, unsigned* __restrict__ bar_counter, float* __restrict__ trace, int max_iters,
                    int variant_flags) {
    constexpr int NS = SumLayout<MIN>::N;
    __shared__ IcpState st;
    __shared__ uint32_t sh[kSel0Bins];   // radix level 0 histogram, then staging / the candidate list
    __shared__ uint32_t sh2[1024];       // local radix levels
    __shared__ uint32_t s_warp[kLoopWarps + 1];
    __shared__ uint32_t s_bin, s_res, s_cnt, s_stage, s_base;
    __shared__ double s_part[kLoopWarps][NS];
    __shared__ double s_red[kLoopWarps][kAccSlots];
    __shared__ double s_sum[kAccSlots];
    __shared__ float s_scratch[16];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(IcpState) / 4); i += kLoopThreads)
        reinterpret_cast<uint32_t*>(&st)[i] = reinterpret_cast<const uint32_t*>(gst)[i];
    __syncthreads();
    const int nq = st.nq;
    unsigned epoch = 0;
    const bool use_quantile = prm.quantile_filter >= 0;
    const int lig = lane & (kLoopG - 1);
    const unsigned gmask = group_mask<kLoopG>(lane);
    constexpr int kPerSweep = kLoopThreads / kLoopG;  // queries per CTA per sweep

    for (int it = 0; it < max_iters; ++it) {
        if (st.done) break;  // identical in every CTA
        const bool stamper = blockIdx.x == 0 && tid == 0;
        if (stamper) B200_STAMP(gst, 20);
        const bool searched = st.iter > 0;  // iteration 0's matches come from the cold kernel
        unsigned long long t_iter0 = 0;
        if (blockIdx.x == 0 && tid == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_iter0));
            if (it == 0) st.loop_total_ns = t_iter0;  // start mark, turned into a duration at the end
        }
        if (use_quantile) {
            for (int i = tid; i < kSel0Bins; i += kLoopThreads) sh[i] = 0u;
            __syncthreads();
        }
        // ---- correspondence search (iteration 0 was done by the cold kernel) ----------------------
        if (st.iter > 0) {
            for (long long base = (long long)blockIdx.x * kPerSweep; base < nq; base += (long long)gridDim.x * kPerSweep) {
                const long long qi = base + tid / kLoopG;
                if (qi < nq) {
                    const float4 q4 = __ldg(reading + qi);
                    const int prev = mpos[qi];
                    const float3 q = apply_T(st.T, q4);
                    float bd = CUDART_INF_F;
                    int bp = -1;
                    const bool finite_q = (fabsf(q.x) < 3.0e38f) && (fabsf(q.y) < 3.0e38f) && (fabsf(q.z) < 3.0e38f);
                    if (finite_q) {
                        float tau = prm.max_r2;
                        if (prev >= 0) {
                            const float dprev = dist2_exact(q.x, q.y, q.z, __ldg(g.pts + prev));
                            if (dprev <= prm.max_r2) {
                                tau = dprev;
                                bd = dprev;
                                bp = prev;
                            }
                        }
                        if (tau < CUDART_INF_F) search_ball<kLoopG>(g, q.x, q.y, q.z, tau, bd, bp, lig);
                    }
#pragma unroll
                    for (int o = kLoopG / 2; o > 0; o >>= 1) {
                        const float od = __shfl_xor_sync(gmask, bd, o);
                        const int op = __shfl_xor_sync(gmask, bp, o);
                        if (od < bd || (od == bd && (unsigned)op < (unsigned)bp)) {
                            bd = od;
                            bp = op;
                        }
                    }
                    if (!(bd <= prm.max_r2)) {
                        bd = CUDART_INF_F;
                        bp = -1;
                    }
                    if (lig == 0) {
                        mpos[qi] = bp;
                        md2[qi] = bd;
                        if (use_quantile && bd < CUDART_INF_F) atomicAdd(&sh[__float_as_uint(bd) >> kSel0Shift], 1u);  // radix level 0
                    }
                }
            }
        }
        __syncthreads();  // this CTA's matches are written before any of its threads reads them
        if (blockIdx.x == 0 && tid == 0 && searched) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            st.loop_search_ns += t1 - t_iter0;
            st.loop_iters_timed += 1;
        }
        if (stamper) B200_STAMP(gst, 21);
        // ---- exact quantile of the finite distances (LPM Matches::getDistsQuantile) -------------------
        float qlimit = 0.f;
        bool fallback_used = false;
        if (use_quantile) {
            // level 0: histogram of bits [30:19] over this CTA's slice (already done while searching)
            if (!searched) {
                for (long long sweep = tid / kPerSweep;; sweep += kLoopThreads / kPerSweep) {
                    const long long base = (sweep * gridDim.x + blockIdx.x) * kPerSweep;
                    if (base >= nq) break;
                    const long long qi = base + (tid % kPerSweep);
                    if (qi < nq) {
                        const uint32_t bits = __float_as_uint(md2[qi]);
                        if (bits < 0x7f800000u) atomicAdd(&sh[bits >> kSel0Shift], 1u);
                    }
                }
                __syncthreads();
            }
            for (int i = tid; i < kSel0Bins; i += kLoopThreads)
                if (sh[i]) atomicAdd(&hist[i], sh[i]);
            if (tid == 0) {
                s_bin = 0;
                s_res = 0;
                s_cnt = 0;
                s_stage = 0;
            }
            if (stamper) B200_STAMP(gst, 22);
            grid_barrier(bar_counter, epoch);
            if (stamper) B200_STAMP(gst, 23);
            const uint32_t total = loop_pick<false>(hist, kSel0Bins, 0u, true, prm.quantile, &s_bin, &s_res, &s_cnt, s_warp);
            const uint32_t b1 = s_bin, r1 = s_res, c1 = s_cnt;
            __syncthreads();
            if (stamper) B200_STAMP(gst, 24);
            uint32_t low = 0;  // the 19 low bits of the quantile
            if (total == 0) {
                // LPM: ConvergenceError("no outlier to filter"); leave the buffers clean and stop everywhere
                grid_barrier(bar_counter, epoch);
                if (blockIdx.x == 0)
                    for (int i = tid; i < kSel0Bins; i += kLoopThreads) hist[i] = 0u;
                if (tid == 0) {
                    st.status = B200ICP_ERR_CONVERGENCE;
                    st.done = 1;
                }
                __syncthreads();
                break;
            }
            if (c1 <= (uint32_t)kSelListCap && !(variant_flags & 8)) {
                // candidate list: this CTA's distances that fall in bucket b1 -> staged in shared memory,
                // one global atomic per CTA reserves their place in the list
                for (long long sweep = tid / kPerSweep;; sweep += kLoopThreads / kPerSweep) {
                    const long long base = (sweep * gridDim.x + blockIdx.x) * kPerSweep;
                    if (base >= nq) break;
                    const long long qi = base + (tid % kPerSweep);
                    if (qi < nq) {
                        const uint32_t bits = __float_as_uint(md2[qi]);
                        if (bits < 0x7f800000u && (bits >> kSel0Shift) == b1) sh[atomicAdd(&s_stage, 1u)] = bits;
                    }
                }
                __syncthreads();
                if (tid == 0) s_base = s_stage ? atomicAdd(&hist[kHistCount], s_stage) : 0u;
                __syncthreads();
                for (uint32_t i = tid; i < s_stage; i += kLoopThreads) hist[kHistList + s_base + i] = sh[i];
                if (stamper) B200_STAMP(gst, 25);
                grid_barrier(bar_counter, epoch);
                if (stamper) B200_STAMP(gst, 26);
                if (blockIdx.x == 0) {  // level 0 histogram and the counter are no longer read by anyone
                    for (int i = tid; i < kSel0Bins; i += kLoopThreads) hist[i] = 0u;
                    if (tid == 0) hist[kHistCount] = 0u;
                }
                for (uint32_t i = tid; i < c1; i += kLoopThreads) sh[i] = __ldcg(hist + kHistList + i);
                // local level 1: bits [18:9]
                sh2[tid] = 0u;
                if (tid == 0) {
                    s_bin = 0;
                    s_res = 0;
                }
                __syncthreads();
                for (uint32_t i = tid; i < c1; i += kLoopThreads) atomicAdd(&sh2[(sh[i] >> 9) & 1023u], 1u);
                __syncthreads();
                loop_pick<true>(sh2, 1024, r1, false, 0.f, &s_bin, &s_res, &s_cnt, s_warp);
                const uint32_t b2 = s_bin, r2 = s_res;
                __syncthreads();
                // local level 2: bits [8:0]
                if (tid < 512) sh2[tid] = 0u;
                if (tid == 0) {
                    s_bin = 0;
                    s_res = 0;
                }
                __syncthreads();
                for (uint32_t i = tid; i < c1; i += kLoopThreads)
                    if (((sh[i] >> 9) & 1023u) == b2) atomicAdd(&sh2[sh[i] & 511u], 1u);
                __syncthreads();
                loop_pick<true>(sh2, 512, r2, false, 0.f, &s_bin, &s_res, &s_cnt, s_warp);
                low = (b2 << 9) | s_bin;
                __syncthreads();
                if (stamper) B200_STAMP(gst, 27);
            } else {
                // fallback (bucket larger than the list): two more global passes, bits [18:8] and [7:0]
                fallback_used = true;
                uint32_t rank = r1, pre = b1;
                for (int pass = 1; pass <= 2; ++pass) {
                    const int nbins = (pass == 1) ? 2048 : 256;
                    uint32_t* gh = hist + (pass == 1 ? kHistL1 : kHistL2);
                    for (int i = tid; i < nbins; i += kLoopThreads) sh[i] = 0u;
                    if (tid == 0) {
                        s_bin = 0;
                        s_res = 0;
                    }
                    __syncthreads();
                    for (long long sweep = tid / kPerSweep;; sweep += kLoopThreads / kPerSweep) {
                        const long long base = (sweep * gridDim.x + blockIdx.x) * kPerSweep;
                        if (base >= nq) break;
                        const long long qi = base + (tid % kPerSweep);
                        if (qi < nq) {
                            const uint32_t bits = __float_as_uint(md2[qi]);
                            if (bits >= 0x7f800000u) continue;
                            if (pass == 1) {
                                if ((bits >> kSel0Shift) == pre) atomicAdd(&sh[(bits >> 8) & 2047u], 1u);
                            } else {
                                if ((bits >> 8) == pre) atomicAdd(&sh[bits & 255u], 1u);
                            }
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < nbins; i += kLoopThreads)
                        if (sh[i]) atomicAdd(&gh[i], sh[i]);
                    grid_barrier(bar_counter, epoch);
                    if (blockIdx.x == 0) {  // the previous level is no longer read by anyone
                        if (pass == 1)
                            for (int i = tid; i < kSel0Bins; i += kLoopThreads) hist[i] = 0u;
                        else
                            for (int i = tid; i < 2048; i += kLoopThreads) hist[kHistL1 + i] = 0u;
                    }
                    loop_pick<false>(gh, nbins, rank, false, 0.f, &s_bin, &s_res, &s_cnt, s_warp);
                    rank = s_res;
                    pre = (pass == 1) ? ((pre << 11) | s_bin) : ((pre << 8) | s_bin);
                    __syncthreads();
                }
                low = pre & ((1u << kSel0Shift) - 1u);
            }
            qlimit = __uint_as_float((b1 << kSel0Shift) | low);
        }
        // ---- ErrorElements + error sums over this CTA's slice ----------------------------------------
        float acc[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = 0.f;
        {
            float T[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) T[i] = st.T[i];
            for (long long sweep = tid / kPerSweep;; sweep += kLoopThreads / kPerSweep) {
                const long long base = (sweep * gridDim.x + blockIdx.x) * kPerSweep;
                if (base >= nq) break;
                const long long qi = base + (tid % kPerSweep);
                if (qi < nq) accumulate_entry<MIN>(acc, prm, T, g, nrm, reading, qi, 1, mpos[qi], md2[qi], qlimit);
            }
        }
        {
            float v32[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v32[i] = (i < NS) ? acc[i] : 0.f;
            const float tot = warp_reduce_32slots(v32, lane);  // lane l now holds the warp total of slot l
            if (lane < NS) s_part[warp][lane] = (double)tot;
        }
        __syncthreads();
        if (tid < kAccSlots) {
            double v = 0.0;
            if (tid < NS) {
#pragma unroll
                for (int wv = 0; wv < kLoopWarps; ++wv) v += s_part[wv][tid];
            }
            partials[(size_t)blockIdx.x * kAccSlots + tid] = v;
        }
        if (stamper) B200_STAMP(gst, 31);
        grid_barrier(bar_counter, epoch);
        if (stamper) B200_STAMP(gst, 12);
        if (fallback_used && blockIdx.x == 0)
            for (int i = tid; i < 256; i += kLoopThreads) hist[kHistL2 + i] = 0u;
        // ---- fixed-order reduction of the per-CTA partials, done identically by every CTA -----------
        {
            const int slot = tid & 31, part = tid >> 5;  // 32 parts
            double v = 0.0;
            for (unsigned b = part; b < gridDim.x; b += kLoopWarps) v += __ldcg(partials + (size_t)b * kAccSlots + slot);
            s_red[part][slot] = v;
        }
        __syncthreads();
        if (tid < kAccSlots) {
            double v = 0.0;
#pragma unroll
            for (int part = 0; part < kLoopWarps; ++part) v += s_red[part][tid];
            s_sum[tid] = v;
        }
        __syncthreads();
        if (stamper) B200_STAMP(gst, 13);
        if (tid < 32) {
            if (tid == 0) st.limit = qlimit;
            finish_warp(prm, &st, s_sum, NS, blockIdx.x == 0 ? trace : nullptr, s_scratch);
        }
        __syncthreads();
        if (stamper) B200_STAMP(gst, 14);
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            st.loop_total_ns = t1 - st.loop_total_ns;
        }
        __syncthreads();
        for (int i = tid; i < (int)(sizeof(IcpState) / 4); i += kLoopThreads)
            reinterpret_cast<uint32_t*>(gst)[i] = reinterpret_cast<const uint32_t*>(&st)[i];
    }
}

template <int MIN, int G>
cudaError_t launch_loop_t(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                          int variant_flags, cudaStream_t s) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_loop_kernel<MIN, G>, kLoopThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int blocks = std::min(n_sms, kLoopMaxBlocks);
    IcpParams prm = p;
    GridView view = g.view;
    const float4* nrm = g.has_normals ? g.normals : nullptr;
    const float4* reading = b.reading;
    int32_t* mpos = b.match_pos;
    float* md2 = b.match_d2;
    IcpState* st = b.state;
    uint32_t* hist = b.hist;
    double* partials = b.partials;
    float* trace = b.trace;
    void* args[] = {&prm, &view, &nrm, &reading, &mpos, &md2, &st, &hist, &partials, &bar_counter, &trace, &max_iters, &variant_flags};
    return cudaLaunchCooperativeKernel((void*)icp_loop_kernel<MIN, G>, dim3(blocks), dim3(kLoopThreads), args, 0, s);
}

template <int MIN>
cudaError_t launch_loop_g(int variant, const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters,
                          int n_sms, cudaStream_t s) {
    switch ((variant >> 8) & 0xf) {  // experimental: lanes per query in the warm search
        case 1: return launch_loop_t<MIN, 1>(p, g, b, bar_counter, max_iters, n_sms, variant, s);
        case 2: return launch_loop_t<MIN, 2>(p, g, b, bar_counter, max_iters, n_sms, variant, s);
        case 8: return launch_loop_t<MIN, 8>(p, g, b, bar_counter, max_iters, n_sms, variant, s);
        default: return launch_loop_t<MIN, 4>(p, g, b, bar_counter, max_iters, n_sms, variant, s);
    }
}

}  // namespace

cudaError_t launch_icp_loop(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                            int variant, cudaStream_t s) {
    if (p.knn != 1) return cudaErrorInvalidValue;
    if (p.minimizer == B200ICP_MIN_POINT_TO_PLANE) return launch_loop_g<0>(variant, p, g, b, bar_counter, max_iters, n_sms, s);
    if (p.minimizer == B200ICP_MIN_POINT_TO_POINT) return launch_loop_g<1>(variant, p, g, b, bar_counter, max_iters, n_sms, s);
    return launch_loop_g<2>(variant, p, g, b, bar_counter, max_iters, n_sms, s);
}

}  // namespace b200
