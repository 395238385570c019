// knn.cu -- exact k-nearest-neighbour search on the cell-sorted grid (the dominant kernel).
//
// Replaces KDTreeMatcher::findClosests -> Nabo::NNS::knn inside the ICP loop
// (/root/reference/norlab_icp_mapper/Mapper.cpp:213) and the direct Nabo::NNS::knn call sites
// (MapperModules/PointDistanceMapperModule.cpp:36, DynamicPointsMapperModule.cpp:78).
// Semantics kept: squared L2 over the first `dim` coordinates, self matches
// allowed, accepted iff dist2 <= maxDist^2, ascending per query, missing = id -1 / dist +inf,
// epsilon = 0 (exact).
//
// Mapping to the hardware: a query is served by a group of G lanes (G = 8 for k <= 8, 16, 32).
// The group walks Chebyshev shells of cells around the query's cell.  Within a shell every lane
// describes one (y, z) row: because points are sorted by linear cell id with x fastest, the cells
// [x0..x1] of a row are ONE contiguous run of float4, located with two loads from the cell table
// and then read by the group with coalesced 16-byte loads (G lanes = G*16 contiguous bytes).
// Rows and x-ranges that cannot hold anything closer than the current k-th best are pruned with
// exact geometric bounds (slack-protected against fp32 rounding), and the walk stops as soon as
// the k-th best is inside the radius the visited block guarantees.  k = 1 keeps a per-lane best
// and reduces with shuffles; k > 1 keeps ONE sorted list per group, lane j holding the j-th best,
// with ballot-driven insertion (warp-ballot top-k).
#include "knn_device.cuh"

namespace b200 {
namespace {

template <int G, typename Acc>
__global__ void __launch_bounds__(256) knn_kernel(GridView g, const float4* __restrict__ queries, const int* __restrict__ d_nq,
                                                  const IcpState* __restrict__ state, int k, float max_r2,
                                                  int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                                  int want_original_ids, int variant, int warm, KnnSpec spec) {
    if (state && state->done) return;
    const int nq = *d_nq;
    const int lane = threadIdx.x & 31;
    const int lig = lane & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const long long qi = (long long)blockIdx.x * (256 / G) + threadIdx.x / G;
    if (qi >= nq) return;  // whole group leaves together

    const float4 q4 = __ldg(queries + qi);
    float qx = q4.x, qy = q4.y, qz = q4.z;
    if (state) {  // stepReading = T_iter * reading, same operation order as the oracle
        const float* T = state->T;
        qx = __fadd_rn(__fmaf_rn(T[8], q4.z, __fmaf_rn(T[4], q4.y, __fmul_rn(T[0], q4.x))), T[12]);
        qy = __fadd_rn(__fmaf_rn(T[9], q4.z, __fmaf_rn(T[5], q4.y, __fmul_rn(T[1], q4.x))), T[13]);
        qz = __fadd_rn(__fmaf_rn(T[10], q4.z, __fmaf_rn(T[6], q4.y, __fmul_rn(T[2], q4.x))), T[14]);
    }
    if (max_r2 < 0.f) max_r2 = q4.w;  // per-point search radius (the reading's `maxSearchDist` descriptor, squared by the prep kernel)
    Acc acc;
    float bound = CUDART_INF_F;
    if (Acc::kPerLaneOutput && warm) {
        // warm start for k > 1 (ICP iterations >= 1): out_ids still holds the previous iteration's k
        // matches -- k distinct real map points -- so the largest of their distances to the moved query
        // bounds the new k-th distance.  The bound only prunes; the result stays exact.
        float dj = 0.f;
        if (lig < k) {
            const int pj = out_ids[qi * k + lig];
            dj = (pj >= 0) ? dist2_exact(qx, qy, qz, __ldg(g.pts + pj)) : CUDART_INF_F;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) dj = fmaxf(dj, __shfl_xor_sync(gmask, dj, o));
        if (dj < CUDART_INF_F) bound = __uint_as_float(__float_as_uint(dj) + 1u);  // next float up: ties with the bound stay accepted
    }
    // speculative bound (k > 1, cold): the k nearest are expected within sqrt(spec.bound2); only closer candidates enter the list
    // and the walk stops there.  k points found inside it ARE the k nearest; a query that finds fewer is listed for an unbounded rerun.
    if (Acc::kPerLaneOutput && spec.list) {
        float b2 = spec.bound2;
        if (spec.per_query && q4.w > 0.f) {
            b2 = q4.w;
        } else if (!(b2 > 0.f)) {
            // points of the query's own cell -> surface density -> radius expected to hold 2.5 k of them, at most one cell edge
            const int cx = min(max((int)floorf((qx - g.ox) * g.inv_h), 0), g.nx - 1), cy = min(max((int)floorf((qy - g.oy) * g.inv_h), 0), g.ny - 1);
            const int cz = min(max((int)floorf((qz - g.oz) * g.inv_h), 0), g.nz - 1);
            const uint32_t* cell = g.cell_start + ((size_t)cz * g.ny + cy) * (size_t)g.nx + cx;
            const float c0 = fmaxf((float)(__ldg(cell + 1) - __ldg(cell)), 1.f);
            b2 = g.h * g.h * fminf(1.0f, fmaxf(0.02f, (2.5f * (float)k) / (3.14159265f * c0)));
        }
        bound = fminf(bound, b2);
    }
    acc.init(k, bound);
    search_shells<G, Acc>(g, acc, qx, qy, qz, max_r2, variant, lig, gmask);
    float od;
    int op;
    acc.finish(gmask, lig, k, max_r2, od, op);
    if (Acc::kPerLaneOutput && spec.list) {
        const int op_k = __shfl_sync(gmask, op, k - 1, G);
        if (op_k < 0 && lig == 0) {
            const unsigned slot = atomicAdd(spec.count, 1u);
            if (slot < spec.capacity) spec.list[slot] = (uint32_t)qi;
        }
    }
    const int n_out = (Acc::kPerLaneOutput) ? k : 1;
    if (lig < n_out) {
        int id = op;
        if (want_original_ids && op >= 0) id = __float_as_int(__ldg(g.pts + op).w);
        const long long o = qi * k + (Acc::kPerLaneOutput ? lig : 0);
        out_ids[o] = id;
        out_d2[o] = od;
    }
}

// ICP iterations >= 1, k = 1: match_pos holds the previous iteration's matches on entry.
template <int G>
__global__ void __launch_bounds__(256) nn1_warm_kernel(GridView g, const float4* __restrict__ reading,
                                                       const IcpState* __restrict__ state, float max_r2,
                                                       int32_t* __restrict__ match_pos, float* __restrict__ match_d2, int) {
    if (state->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) B200_STAMP(const_cast<IcpState*>(state), 16);
    const int nq = state->nq;
    const int lane = threadIdx.x & 31;
    const int lig = lane & (G - 1);
    const unsigned gmask = group_mask<G>(lane);
    const long long qi = (long long)blockIdx.x * (256 / G) + threadIdx.x / G;
    if (qi >= nq) return;
    const float4 q4 = __ldg(reading + qi);
    if (max_r2 < 0.f) max_r2 = q4.w;  // per-point search radius
    const int prev = match_pos[qi];
    const float* T = state->T;
    const float qx = __fadd_rn(__fmaf_rn(T[8], q4.z, __fmaf_rn(T[4], q4.y, __fmul_rn(T[0], q4.x))), T[12]);
    const float qy = __fadd_rn(__fmaf_rn(T[9], q4.z, __fmaf_rn(T[5], q4.y, __fmul_rn(T[1], q4.x))), T[13]);
    const float qz = __fadd_rn(__fmaf_rn(T[10], q4.z, __fmaf_rn(T[6], q4.y, __fmul_rn(T[2], q4.x))), T[14]);
    float bd = CUDART_INF_F;
    int bp = -1;
    const bool finite_q = (fabsf(qx) < 3.0e38f) && (fabsf(qy) < 3.0e38f) && (fabsf(qz) < 3.0e38f);
    if (finite_q) {
        float tau = max_r2;
        if (prev >= 0) {
            const float dprev = dist2_exact(qx, qy, qz, __ldg(g.pts + prev));
            if (dprev <= max_r2) {
                tau = dprev;
                bd = dprev;
                bp = prev;
            }
        }
        // tau == inf only when maxDist is unbounded AND the cold search of iteration 0 found
        // nothing, i.e. the map cannot offer this query a neighbour at all: leave it unmatched.
        if (tau < CUDART_INF_F) search_ball<G>(g, qx, qy, qz, tau, bd, bp, lig);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(gmask, bd, o);
        const int op = __shfl_xor_sync(gmask, bp, o);
        if (od < bd || (od == bd && (unsigned)op < (unsigned)bp)) {
            bd = od;
            bp = op;
        }
    }
    if (!(bd <= max_r2)) {
        bd = CUDART_INF_F;
        bp = -1;
    }
    if (lig == 0) {
        match_pos[qi] = bp;
        match_d2[qi] = bd;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) B200_STAMP(const_cast<IcpState*>(state), 17);
}

// Cold k = 1 search with a finite maxDist (iteration 0 of a registration): the bounded ball search of the loop kernel with
// tau0 = maxDist^2 -- the 2 x 2 nearest cell rows first, balanced 4-lane scan, then whatever the tightened ball still touches.
// Same exact result and tie rule as knn_kernel<8, Acc1>; about half its time at maxDist ~ 2-3 cell edges.
__global__ void __launch_bounds__(256) nn1_cold_kernel(GridView g, const float4* __restrict__ queries, const int* __restrict__ d_nq,
                                                       const IcpState* __restrict__ state, float max_r2, int32_t* __restrict__ out_ids,
                                                       float* __restrict__ out_d2, int want_original_ids) {
    if (state && state->done) return;
    const int nq = *d_nq;
    const int lane = threadIdx.x & 31;
    const int lig = lane & 3;
    const unsigned gmask = group_mask<4>(lane);
    const long long qi = (long long)blockIdx.x * 64 + threadIdx.x / 4;
    if (qi >= nq) return;  // whole group leaves together
    const float4 q4 = __ldg(queries + qi);
    if (max_r2 < 0.f) max_r2 = q4.w;  // per-point search radius
    float qx = q4.x, qy = q4.y, qz = q4.z;
    if (state) {
        const float* T = state->T;
        qx = __fadd_rn(__fmaf_rn(T[8], q4.z, __fmaf_rn(T[4], q4.y, __fmul_rn(T[0], q4.x))), T[12]);
        qy = __fadd_rn(__fmaf_rn(T[9], q4.z, __fmaf_rn(T[5], q4.y, __fmul_rn(T[1], q4.x))), T[13]);
        qz = __fadd_rn(__fmaf_rn(T[10], q4.z, __fmaf_rn(T[6], q4.y, __fmul_rn(T[2], q4.x))), T[14]);
    }
    float bd = CUDART_INF_F, sd = CUDART_INF_F;
    int bp = -1;
    const bool finite_q = (fabsf(qx) < 3.0e38f) && (fabsf(qy) < 3.0e38f) && (fabsf(qz) < 3.0e38f);
    if (finite_q) search_ball4<4>(g, qx, qy, qz, max_r2, 0.f, bd, bp, sd, lig, gmask);
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(gmask, bd, o);
        const int op = __shfl_xor_sync(gmask, bp, o);
        if (od < bd || (od == bd && (unsigned)op < (unsigned)bp)) {
            bd = od;
            bp = op;
        }
    }
    if (!(bd <= max_r2)) {
        bd = CUDART_INF_F;
        bp = -1;
    }
    if (lig == 0) {
        out_ids[qi] = (want_original_ids && bp >= 0) ? __float_as_int(__ldg(g.pts + bp).w) : bp;
        out_d2[qi] = bd;
    }
}

template <int G>
cudaError_t launch_warm_one(const GridView& g, const float4* reading, int cap, const IcpState* st, float max_r2,
                            int32_t* pos, float* d2, int variant, cudaStream_t s) {
    const int per_block = 256 / G;
    const int blocks = (cap + per_block - 1) / per_block;
    if (blocks <= 0) return cudaSuccess;
    nn1_warm_kernel<G><<<blocks, 256, 0, s>>>(g, reading, st, max_r2, pos, d2, variant);
    return cudaGetLastError();
}

template <int G, typename Acc>
cudaError_t launch_one(const GridView& g, const float4* q, const int* d_nq, int cap, const IcpState* st, int k, float max_r2,
                       int32_t* ids, float* d2, int want_orig, int variant, cudaStream_t s, int warm = 0, KnnSpec spec = KnnSpec()) {
    const int per_block = 256 / G;
    const int blocks = (cap + per_block - 1) / per_block;
    if (blocks <= 0) return cudaSuccess;
    knn_kernel<G, Acc><<<blocks, 256, 0, s>>>(g, q, d_nq, st, k, max_r2, ids, d2, want_orig, variant, warm, spec);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_knn(const GridView& g, const float4* d_queries, const int* d_nq, int nq_capacity,
                       const IcpState* st, int k, float max_r2, int32_t* out_ids, float* out_d2, int want_original_ids,
                       int variant, cudaStream_t s, KnnSpec spec) {
    if (k < 1 || k > 32) return cudaErrorInvalidValue;
    if (k == 1 && max_r2 < 3.0e38f && !(variant & 256) && sqrtf(max_r2) * g.inv_h <= 4.0f) {  // (bit 8: the shell-walk kernel instead)
        const int blocks = (nq_capacity + 63) / 64;
        if (blocks <= 0) return cudaSuccess;
        nn1_cold_kernel<<<blocks, 256, 0, s>>>(g, d_queries, d_nq, st, max_r2, out_ids, out_d2, want_original_ids);
        return cudaGetLastError();
    }
    if (k == 1) return launch_one<8, Acc1<8>>(g, d_queries, d_nq, nq_capacity, st, 1, max_r2, out_ids, out_d2, want_original_ids, variant, s);
    const int warm = (variant & 0x10000) ? 1 : 0;  // set by the ICP loop from iteration 1 on (out_ids = previous matches, positions)
    // lanes per query: at least k (lane j holds the j-th best); few queries get more lanes each -- a query's latency is what
    // is left to cut when the whole batch does not even fill the SMs (10 k queries x 8 lanes = a quarter of the B200)
    int lanes = k <= 8 ? 8 : (k <= 16 ? 16 : 32);
    while (lanes < 32 && (long long)nq_capacity * lanes * 2 <= (long long)kSMs * 2048) lanes *= 2;
    if (lanes == 8) return launch_one<8, AccK<8>>(g, d_queries, d_nq, nq_capacity, st, k, max_r2, out_ids, out_d2, want_original_ids, variant, s, warm, spec);
    if (lanes == 16) return launch_one<16, AccK<16>>(g, d_queries, d_nq, nq_capacity, st, k, max_r2, out_ids, out_d2, want_original_ids, variant, s, warm, spec);
    return launch_one<32, AccK<32>>(g, d_queries, d_nq, nq_capacity, st, k, max_r2, out_ids, out_d2, want_original_ids, variant, s, warm, spec);
}

cudaError_t launch_nn1_warm(const GridView& g, const float4* d_reading, int nq_capacity, const IcpState* st, float max_r2,
                            int32_t* match_pos, float* match_d2, int variant, cudaStream_t s) {
    return launch_warm_one<4>(g, d_reading, nq_capacity, st, max_r2, match_pos, match_d2, variant, s);
}

}  // namespace b200
