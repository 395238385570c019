// knn_device.cuh -- device-side building blocks of the exact k-NN search (see knn.cu for the design
// notes): accumulators, the pruned shell walk (cold search) and the ball search (warm search).
// Included by knn.cu (stand-alone kernels) and loop.cu (persistent ICP loop kernel).
#pragma once
#include "common.cuh"

namespace b200 {
namespace {


template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
    if constexpr (G == 32) {
        return 0xffffffffu;
    } else {
        return ((1u << G) - 1u) << (lane & ~(G - 1));
    }
}

__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, const float4& p) {
    const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ int floor_to_int(float v) {  // v already clamped to a sane range
    return (int)floorf(v);
}

// ---- k = 1: per-lane best, shuffle reduction ---------------------------------------------------
template <int G>
struct Acc1 {
    static constexpr bool kPerLaneOutput = false;
    float d;
    int pos;
    __device__ __forceinline__ void init(int, float) {
        d = CUDART_INF_F;
        pos = -1;
    }
    __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                         int lig, unsigned, float) {
        for (uint32_t j = s + lig; j < e; j += G) {
            const float4 p = __ldg(pts + j);
            const float dd = dist2_exact(qx, qy, qz, p);
            if (dd < d || (dd == d && j < (uint32_t)pos)) {  // ties: lowest cell-sorted position wins
                d = dd;
                pos = (int)j;
            }
        }
    }
    // one candidate per lane (cd = inf: none)
    __device__ __forceinline__ void consume(float cd, uint32_t j, int, unsigned, float) {
        if (cd < d || (cd == d && j < (uint32_t)pos)) {
            d = cd;
            pos = (int)j;
        }
    }
    // squared radius beyond which nothing can improve the result
    __device__ __forceinline__ float tau(unsigned gmask, float max_r2) const {
        float v = d;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(gmask, v, o));
        return fminf(v, max_r2);
    }
    __device__ __forceinline__ void finish(unsigned gmask, int lig, int, float max_r2, float& out_d, int& out_pos) {
        float bd = d;
        int bp = pos;
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
        out_d = bd;
        out_pos = bp;
        (void)lig;
    }
};

// ---- k > 1: one sorted list per group, lane j holds the j-th best ------------------------------
// TRACK (loop.cu): the search also covers `margin` beyond the k-th distance and remembers, per lane, the smallest
// distance among the points it tested and did not keep -- together a lower bound for every point NOT in the result.
template <int G, bool TRACK = false>
struct AccK {
    static constexpr bool kPerLaneOutput = true;
    float d;      // my entry
    int pos;
    float kth;    // group-uniform: current k-th best (inf until k found), never above `bound`
    float bound;  // group-uniform: a proven upper bound of the final k-th distance (warm start) or inf
    int k;
    float margin;  // TRACK: extra radius covered around the ball of the k-th distance
    float sd;      // TRACK, per lane: smallest dist2 tested and refused, or pushed out of the list
    __device__ __forceinline__ void init(int k_, float bound_, float margin_ = 0.f) {
        d = CUDART_INF_F;
        pos = -1;
        bound = bound_;
        kth = bound_;
        k = k_;
        margin = margin_;
        sd = CUDART_INF_F;
    }
    // one candidate per lane (cd = inf: none), taken in lane order: ballot-driven insertion into the group's sorted list
    __device__ __forceinline__ void consume(float cd, uint32_t j, int lig, unsigned gmask, float max_r2) {
        bool pass = (cd < kth) && (cd <= max_r2);
        if (TRACK && !pass) sd = fminf(sd, cd);
        unsigned m = __ballot_sync(gmask, pass) & gmask;
        while (m) {
            const int src = __ffs(m) - 1;  // absolute lane
            const float nd = __shfl_sync(gmask, cd, src);
            const int np = (int)__shfl_sync(gmask, j, src);
            if (TRACK) sd = fminf(sd, __shfl_sync(gmask, d, k - 1, G));  // the entry this insertion pushes out (inf: list not full)
            // rank of the newcomer = entries <= nd (stable: goes after equal distances)
            const unsigned le = __ballot_sync(gmask, d <= nd) & gmask;
            const int r = __popc(le);
            const float pd = __shfl_up_sync(gmask, d, 1, G);
            const int pp = __shfl_up_sync(gmask, pos, 1, G);
            if (lig == r) {
                d = nd;
                pos = np;
            } else if (lig > r) {
                d = pd;
                pos = pp;
            }
            if (lig >= k) {
                d = CUDART_INF_F;
                pos = -1;
            }
            kth = fminf(bound, __shfl_sync(gmask, d, k - 1, G));
            m &= m - 1;
            const bool still = pass && lig != (src & (G - 1)) && (cd < kth);  // (the lane just served is done: its point is in the list)
            if (TRACK && pass && !still && lig != (src & (G - 1))) sd = fminf(sd, cd);  // overtaken before its turn came
            pass = still;
            m &= __ballot_sync(gmask, pass);
        }
    }
    __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t s, uint32_t e, float qx, float qy, float qz,
                                         int lig, unsigned gmask, float max_r2) {
        for (uint32_t j0 = s; j0 < e; j0 += G) {
            const uint32_t j = j0 + lig;
            float cd = CUDART_INF_F;
            if (j < e) cd = dist2_exact(qx, qy, qz, __ldg(pts + j));
            consume(cd, j, lig, gmask, max_r2);
        }
    }
    __device__ __forceinline__ float tau(unsigned, float max_r2) const {
        if (TRACK) {
            const float r = sqrtf(fminf(kth, max_r2)) + margin;
            return r * r;
        }
        return fminf(kth, max_r2);
    }
    __device__ __forceinline__ void finish(unsigned, int, int, float, float& out_d, int& out_pos) {
        out_d = d;
        out_pos = pos;
    }
};

// ---- k > 1, warm (loop.cu): COLLECT the points inside a fixed bound, select afterwards --------------------------
// The cached k matches bound the new k-th distance, and the ball of that radius holds little more than k points.  Inserting
// them one by one into a sorted list (AccK) is a chain of ~10 dependent cross-lane operations per point; here every lane
// just keeps what it saw below the bound in C register slots (no cross-lane traffic during the scan), and the k nearest
// are picked afterwards by rank counting (AccCollect::select): every collected point is broadcast once and each lane counts,
// for its own points, how many precede them on (distance, position).  A lane that runs out of slots flags the query for the
// AccK path.  `sd` as in AccK<.., TRACK>: the smallest distance tested and refused.
template <int G, int C>
struct AccCollect {
    static constexpr bool kPerLaneOutput = true;
    float cd[C];
    int cp[C];
    int cnt;
    bool ovf;
    float kth;  // the fixed gate (group-uniform)
    float margin, sd;
    int k;
    __device__ __forceinline__ void init(int k_, float bound_, float margin_ = 0.f) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            cd[c] = CUDART_INF_F;
            cp[c] = -1;
        }
        cnt = 0;
        ovf = false;
        kth = bound_;
        margin = margin_;
        sd = CUDART_INF_F;
        k = k_;
    }
    __device__ __forceinline__ void consume(float d, uint32_t j, int, unsigned, float max_r2) {
        if ((d < kth) && (d <= max_r2)) {
            if (cnt < C) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c == cnt) {
                        cd[c] = d;
                        cp[c] = (int)j;
                    }
                cnt += 1;
            } else {
                ovf = true;
            }
        } else {
            sd = fminf(sd, d);
        }
    }
    __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t s, uint32_t e, float qx, float qy, float qz, int lig, unsigned gmask,
                                         float max_r2) {
#pragma unroll 2
        for (uint32_t j = s + (uint32_t)lig; j < e; j += G) consume(dist2_exact(qx, qy, qz, __ldg(pts + j)), j, lig, gmask, max_r2);
    }
    __device__ __forceinline__ float tau(unsigned, float max_r2) const {
        const float r = sqrtf(fminf(kth, max_r2)) + margin;
        return r * r;
    }
    // rank of each of this lane's points among all collected points of the group ((distance, position) ascending); returns the
    // number collected.  rank[c] is meaningful for c < cnt.
    __device__ __forceinline__ int select(int (&rank)[C], int lig, unsigned gmask) const {
#pragma unroll
        for (int c = 0; c < C; ++c) rank[c] = 0;
        int total = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            unsigned m = __ballot_sync(gmask, c < cnt) & gmask;  // lanes that hold a point in slot c (group-uniform loop below)
            total += __popc(m);
            while (m) {
                const int src = __ffs(m) - 1;  // absolute lane
                m &= m - 1;
                const float nd = __shfl_sync(gmask, cd[c], src);
                const int np = __shfl_sync(gmask, cp[c], src);
#pragma unroll
                for (int s = 0; s < C; ++s) rank[s] += (nd < cd[s]) || (nd == cd[s] && (unsigned)np < (unsigned)cp[s]);
            }
        }
        (void)lig;
        return total;
    }
};

// One shell of cells with Chebyshev distance in (Rprev, R] around (cx, cy, cz), pruned by tau.
template <int G, typename Acc>
__device__ __forceinline__ void visit_shell(const GridView& g, Acc& acc, float qx, float qy, float qz, float ux, float uy,
                                            float uz, int cx, int cy, int cz, int R, int Rprev, float tau, float slack,
                                            float max_r2, int lig, unsigned gmask) {
    const float rt = fminf(sqrtf(tau) * g.inv_h + slack, 3.0e8f);
    const int ylo = max(max(cy - R, 0), floor_to_int(fmaxf(uy - rt, -1.f)));
    const int yhi = min(min(cy + R, g.ny - 1), floor_to_int(fminf(uy + rt, (float)g.ny)));
    const int zlo = max(max(cz - R, 0), floor_to_int(fmaxf(uz - rt, -1.f)));
    const int zhi = min(min(cz + R, g.nz - 1), floor_to_int(fminf(uz + rt, (float)g.nz)));
    const int wy = yhi - ylo + 1, wz = zhi - zlo + 1;
    if (wy <= 0 || wz <= 0) return;
    const int nrows = wy * wz;
    for (int base = 0; base < nrows; base += G) {
        const int r = base + lig;
        uint32_t s1 = 0, e1 = 0, s2 = 0, e2 = 0;
        if (r < nrows) {
            const int y = ylo + r % wy, z = zlo + r / wy;
            const int dy = y - cy, dz = z - cz;
            float gy = dy == 0 ? 0.f : (dy > 0 ? (float)y - uy : uy - (float)(y + 1));
            float gz = dz == 0 ? 0.f : (dz > 0 ? (float)z - uz : uz - (float)(z + 1));
            gy = fmaxf(gy - slack, 0.f) * g.h;
            gz = fmaxf(gz - slack, 0.f) * g.h;
            const float gyz2 = gy * gy + gz * gz;
            if (gyz2 <= tau) {
                const float rx = fminf(sqrtf(fmaxf(tau - gyz2, 0.f)) * g.inv_h + slack, 3.0e8f);
                const int xa = max(max(cx - R, 0), floor_to_int(fmaxf(ux - rx, -1.f)));
                const int xb = min(min(cx + R, g.nx - 1), floor_to_int(fminf(ux + rx, (float)g.nx)));
                const uint32_t* row = g.cell_start + ((size_t)z * g.ny + y) * (size_t)g.nx;
                const bool inner = max(abs(dy), abs(dz)) <= Rprev;
                if (!inner) {
                    if (xa <= xb) {
                        s1 = __ldg(row + xa);
                        e1 = __ldg(row + xb + 1);
                    }
                } else {
                    const int xb1 = min(xb, cx - Rprev - 1);
                    const int xa2 = max(xa, cx + Rprev + 1);
                    if (xa <= xb1) {
                        s1 = __ldg(row + xa);
                        e1 = __ldg(row + xb1 + 1);
                    }
                    if (xa2 <= xb) {
                        s2 = __ldg(row + xa2);
                        e2 = __ldg(row + xb + 1);
                    }
                }
            }
        }
        if (g.nz == 1) {
            // 2-D maps: a shell has few rows and their runs are long -- run by run, G points at a time (flattening them, as
            // below, costs more in bookkeeping than it saves: cfg 4 + 5 %)
            unsigned mrun = __ballot_sync(gmask, (e1 > s1) || (e2 > s2)) & gmask;
            while (mrun) {
                const int src = __ffs(mrun) - 1;
                mrun &= mrun - 1;
                const uint32_t a1 = __shfl_sync(gmask, s1, src), b1 = __shfl_sync(gmask, e1, src);
                const uint32_t a2 = __shfl_sync(gmask, s2, src), b2 = __shfl_sync(gmask, e2, src);
                if (b1 > a1) acc.scan(g.pts, a1, b1, qx, qy, qz, lig, gmask, max_r2);
                if (b2 > a2) acc.scan(g.pts, a2, b2, qx, qy, qz, lig, gmask, max_r2);
            }
            continue;
        }
        // The (up to 2 G) runs of this chunk of rows are consumed as ONE flat sequence, G points at a time in the same order
        // a run-by-run scan would take them (rows in lane order, x ascending): every batch keeps all G lanes busy whatever
        // the run lengths, and the next batch's points are already in flight while the current one is inserted -- the
        // search is a chain of dependent L2 round trips, one per batch instead of one or more per run.
        const uint32_t len1 = e1 > s1 ? e1 - s1 : 0u, len2 = e2 > s2 ? e2 - s2 : 0u;
        uint32_t incl = len1 + len2;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const uint32_t v = __shfl_up_sync(gmask, incl, o, G);
            if (lig >= o) incl += v;
        }
        const uint32_t excl = incl - (len1 + len2);
        const uint32_t total = __shfl_sync(gmask, incl, G - 1, G);
        if (total == 0u) continue;  // group-uniform
        auto locate = [&](uint32_t t) -> uint32_t {  // flat index -> point position (t clamped by the caller)
            int r = 0;  // largest lane with excl <= t
#pragma unroll
            for (int step = G / 2; step > 0; step >>= 1) {
                const uint32_t ex = __shfl_sync(gmask, excl, r + step, G);
                if (ex <= t) r += step;
            }
            const uint32_t off = t - __shfl_sync(gmask, excl, r, G);
            const uint32_t rs1 = __shfl_sync(gmask, s1, r, G), rl1 = __shfl_sync(gmask, len1, r, G), rs2 = __shfl_sync(gmask, s2, r, G);
            return off < rl1 ? rs1 + off : rs2 + (off - rl1);
        };
        uint32_t jn = locate(min((uint32_t)lig, total - 1u));
        float4 pn = __ldg(g.pts + jn);
        for (uint32_t t0 = 0; t0 < total; t0 += G) {
            const float4 pc = pn;
            const uint32_t jc = jn;
            const bool valid = t0 + (uint32_t)lig < total;
            if (t0 + G < total) {  // group-uniform
                jn = locate(min(t0 + G + (uint32_t)lig, total - 1u));
                pn = __ldg(g.pts + jn);
            }
            acc.consume(valid ? dist2_exact(qx, qy, qz, pc) : CUDART_INF_F, jc, lig, gmask, max_r2);
        }
    }
}

// Cold search: walk Chebyshev shells around the query's cell until the k-th best is inside the
// radius the visited block guarantees (or the whole grid has been seen).
template <int G, typename Acc>
__device__ __forceinline__ void search_shells(const GridView& g, Acc& acc, float qx, float qy, float qz, float max_r2,
                                              int variant, int lig, unsigned gmask) {
    // grid coordinates (same fp32 expression as the builder's cell_coord)
    const float lim = 1.0e8f;
    const float ux = fminf(fmaxf((qx - g.ox) * g.inv_h, -lim), lim);
    const float uy = fminf(fmaxf((qy - g.oy) * g.inv_h, -lim), lim);
    const float uz = fminf(fmaxf((qz - g.oz) * g.inv_h, -lim), lim);
    const int cx = floor_to_int(ux), cy = floor_to_int(uy), cz = floor_to_int(uz);
    const float slack = g.slack + 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
    // first shell that can touch the grid
    int R0 = 0;
    R0 = max(R0, cx < 0 ? -cx : (cx > g.nx - 1 ? cx - (g.nx - 1) : 0));
    R0 = max(R0, cy < 0 ? -cy : (cy > g.ny - 1 ? cy - (g.ny - 1) : 0));
    R0 = max(R0, cz < 0 ? -cz : (cz > g.nz - 1 ? cz - (g.nz - 1) : 0));
    int R = R0, Rprev = R0 - 1;
    if ((variant & 1) && R0 == 0) R = 1;  // variant bit 0: no own-cell pre-pass
    const bool finite_q = (fabsf(qx) < 3.0e38f) && (fabsf(qy) < 3.0e38f) && (fabsf(qz) < 3.0e38f);  // false for NaN too
    if (!finite_q) return;
    while (true) {
        const float tau = acc.tau(gmask, max_r2);
        visit_shell<G, Acc>(g, acc, qx, qy, qz, ux, uy, uz, cx, cy, cz, R, Rprev, tau, slack, max_r2, lig, gmask);
        // radius guaranteed by the visited block: distance to the nearest face that still has cells behind it
        float gu = CUDART_INF_F;
        if (cx - R > 0) gu = fminf(gu, ux - (float)(cx - R));
        if (cx + R < g.nx - 1) gu = fminf(gu, (float)(cx + R + 1) - ux);
        if (cy - R > 0) gu = fminf(gu, uy - (float)(cy - R));
        if (cy + R < g.ny - 1) gu = fminf(gu, (float)(cy + R + 1) - uy);
        if (cz - R > 0) gu = fminf(gu, uz - (float)(cz - R));
        if (cz + R < g.nz - 1) gu = fminf(gu, (float)(cz + R + 1) - uz);
        if (gu == CUDART_INF_F) break;  // whole grid visited
        const float gm = fmaxf(gu - slack, 0.f) * g.h;
        if (acc.tau(gmask, max_r2) <= gm * gm) break;
        Rprev = R;
        R = R + 1;
    }
}

// Warm search for k > 1 (loop.cu): k real map points are known to lie within acc.bound of the query, so ONE pass over the
// cells that intersect the ball of radius sqrt(min(bound, maxDist^2)) + margin sees everything that matters -- no shell walk,
// no exit test.  On return every map point within that radius has been offered to the accumulator.
template <int G, typename Acc>
__device__ __forceinline__ void search_ball_k(const GridView& g, Acc& acc, float qx, float qy, float qz, float max_r2, int lig,
                                              unsigned gmask) {
    const float lim = 1.0e8f;
    const float ux = fminf(fmaxf((qx - g.ox) * g.inv_h, -lim), lim);
    const float uy = fminf(fmaxf((qy - g.oy) * g.inv_h, -lim), lim);
    const float uz = fminf(fmaxf((qz - g.oz) * g.inv_h, -lim), lim);
    const int cx = floor_to_int(ux), cy = floor_to_int(uy), cz = floor_to_int(uz);
    const float slack = g.slack + 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
    const float tau = acc.tau(gmask, max_r2);
    const int R = (int)fminf(sqrtf(tau) * g.inv_h + slack + 2.f, 1.0e6f);  // Chebyshev cell radius that contains the ball
    visit_shell<G, Acc>(g, acc, qx, qy, qz, ux, uy, uz, cx, cy, cz, R, -1, tau, slack, max_r2, lig, gmask);
}

// Warm search (k = 1, ICP iterations >= 1): the previous iteration's match is a real map point, so
// its distance to the moved query bounds the new nearest distance.  Every cell that intersects the
// ball of that radius is scanned in ONE pass -- no shell walk, no pre-pass.  Each lane of the group
// owns whole (y, z) rows of the ball's cover, so the loads of different rows are in flight together
// and the inner loop has no cross-lane traffic; the result is exact for the same reason the bound
// is valid, and identical to the cold search thanks to the (dist2, position) tie rule.
template <int G>
__device__ __forceinline__ void search_ball(const GridView& g, float qx, float qy, float qz, float tau, float& bd, int& bp,
                                            int lig) {
    const float lim = 1.0e8f;
    const float ux = fminf(fmaxf((qx - g.ox) * g.inv_h, -lim), lim);
    const float uy = fminf(fmaxf((qy - g.oy) * g.inv_h, -lim), lim);
    const float uz = fminf(fmaxf((qz - g.oz) * g.inv_h, -lim), lim);
    const float slack = g.slack + 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
    const float rt = fminf(sqrtf(tau) * g.inv_h + slack, 3.0e8f);
    const int ylo = max(0, floor_to_int(fmaxf(uy - rt, -1.f)));
    const int yhi = min(g.ny - 1, floor_to_int(fminf(uy + rt, (float)g.ny)));
    const int zlo = max(0, floor_to_int(fmaxf(uz - rt, -1.f)));
    const int zhi = min(g.nz - 1, floor_to_int(fminf(uz + rt, (float)g.nz)));
    const int wy = yhi - ylo + 1, wz = zhi - zlo + 1;
    if (wy <= 0 || wz <= 0) return;
    const int nrows = wy * wz;
    for (int r = lig; r < nrows; r += G) {
        const int y = ylo + r % wy, z = zlo + r / wy;
        // distance from the query to the row's y/z slab (0 when inside it)
        float gy = fmaxf(fmaxf((float)y - uy, uy - (float)(y + 1)), 0.f);
        float gz = fmaxf(fmaxf((float)z - uz, uz - (float)(z + 1)), 0.f);
        gy = fmaxf(gy - slack, 0.f) * g.h;
        gz = fmaxf(gz - slack, 0.f) * g.h;
        const float gyz2 = gy * gy + gz * gz;
        if (gyz2 > tau) continue;
        const float rx = fminf(sqrtf(fmaxf(tau - gyz2, 0.f)) * g.inv_h + slack, 3.0e8f);
        const int xa = max(0, floor_to_int(fmaxf(ux - rx, -1.f)));
        const int xb = min(g.nx - 1, floor_to_int(fminf(ux + rx, (float)g.nx)));
        if (xa > xb) continue;
        const uint32_t* row = g.cell_start + ((size_t)z * g.ny + y) * (size_t)g.nx;
        const uint32_t s = __ldg(row + xa), e = __ldg(row + xb + 1);
#pragma unroll 4
        for (uint32_t j = s; j < e; ++j) {
            const float4 p = __ldg(g.pts + j);
            const float dd = dist2_exact(qx, qy, qz, p);
            if (dd < bd || (dd == bd && j < (uint32_t)bp)) {
                bd = dd;
                bp = (int)j;
            }
        }
        tau = fminf(tau, bd);
    }
}

// Ball search with a known bound tau0 (4 lanes per query): the loop kernel's S phase and the cold search of iteration 0
// when maxDist is finite.
//  phase 1: the 2 x 2 block of (y, z) cell rows nearest to the query, one row per lane -- two cell-table
//           loads give the row's contiguous run of points -- then a group-wide minimum: this usually
//           finds the nearest neighbour and tightens the bound a loose previous match gave;
//  phase 2: whatever other rows the tightened ball still touches (usually none), lane-strided.
// Unlike search_ball (knn_device.cuh) the covered radius is sqrt(min(best, tau0)) + m instead of
// sqrt(best): on return every map point within that radius of the query has been looked at, so besides
// the exact nearest neighbour (bd, bp per lane; ties: lowest position) the lanes also know sd = the
// smallest squared distance to any OTHER point seen.  bp / bd come in as the previous match (or -1 / inf)
// in every lane.
__device__ __forceinline__ void test_point(float qx, float qy, float qz, const float4& p, uint32_t j, float& bd, int& bp, float& sd) {
    const float dd = dist2_exact(qx, qy, qz, p);
    if (dd < bd || (dd == bd && j < (uint32_t)bp)) {
        sd = bd;  // sd >= bd always: the dethroned best becomes the second
        bd = dd;
        bp = (int)j;
    } else {
        sd = fminf(sd, dd);
    }
}

// run of points [s, e) of the cells [xa, xb] of row (y, z) that a ball of squared radius cov2 around the query can touch
__device__ __forceinline__ void row_run(const GridView& g, float ux, float slack, int y, int z, float gyz2, float cov2, uint32_t& s, uint32_t& e) {
    const float rx = fminf(sqrtf(fmaxf(cov2 - gyz2, 0.f)) * g.inv_h + slack, 3.0e8f);
    const int xa = max(0, floor_to_int(fmaxf(ux - rx, -1.f)));
    const int xb = min(g.nx - 1, floor_to_int(fminf(ux + rx, (float)g.nx)));
    if (xa > xb) return;
    const uint32_t* row = g.cell_start + ((size_t)z * g.ny + y) * (size_t)g.nx;
    s = __ldg(row + xa);
    e = __ldg(row + xb + 1);
}

__device__ __forceinline__ float row_gap2(const GridView& g, float uy, float uz, float slack, int y, int z) {
    float gy = fmaxf(fmaxf((float)y - uy, uy - (float)(y + 1)), 0.f);
    float gz = fmaxf(fmaxf((float)z - uz, uz - (float)(z + 1)), 0.f);
    gy = fmaxf(gy - slack, 0.f) * g.h;
    gz = fmaxf(gz - slack, 0.f) * g.h;
    return gy * gy + gz * gz;
}

// The runs [s, e) the four ROW lanes of a group hold (one cell row each, possibly empty), concatenated, are scanned by the
// SL lanes of the group together: balanced, and neighbouring lanes read contiguous bytes.  Returns the group-wide best distance.
// SL = 4: a lane quartet per query (the row lanes are the scanning lanes).  SL = 32: a whole warp per query, rows still described
// by its lanes 0..3 -- one round trip for a run of up to 32 points instead of one per 4, for the handful of queries a converged
// iteration still has to search (their latency is what every other CTA waits for at the barrier).
template <int SL>
__device__ __forceinline__ float scan_runs(const GridView& g, float qx, float qy, float qz, uint32_t s, uint32_t e, float& bd, int& bp, float& sd,
                                           int lane_s, unsigned gmask) {
    const uint32_t n = e - s;
    const uint32_t n0 = __shfl_sync(gmask, n, 0, SL), n1 = __shfl_sync(gmask, n, 1, SL), n2 = __shfl_sync(gmask, n, 2, SL), n3 = __shfl_sync(gmask, n, 3, SL);
    const uint32_t p1 = n0, p2 = p1 + n1, p3 = p2 + n2, N = p3 + n3;
    const uint32_t o0 = __shfl_sync(gmask, s, 0, SL), o1 = __shfl_sync(gmask, s, 1, SL) - p1, o2 = __shfl_sync(gmask, s, 2, SL) - p2,
                   o3 = __shfl_sync(gmask, s, 3, SL) - p3;
#pragma unroll 2
    for (uint32_t k = (uint32_t)lane_s; k < N; k += SL) {
        const uint32_t j = k + (k >= p2 ? (k >= p3 ? o3 : o2) : (k >= p1 ? o1 : o0));
        test_point(qx, qy, qz, __ldg(g.pts + j), j, bd, bp, sd);
    }
    float v = bd;
#pragma unroll
    for (int o = SL / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(gmask, v, o));
    return v;
}

// bd / bp / sd start at inf / -1 / inf in every lane: the previous match only bounds the ball (tau0) and is found
// again by the scan like any other point.
template <int SL>
__device__ __forceinline__ void search_ball4(const GridView& g, float qx, float qy, float qz, float tau0, float m, float& bd, int& bp,
                                             float& sd, int lane_s, unsigned gmask) {
    constexpr int G = 4;  // rows described per round
    const int lig = lane_s;
    const bool row_lane = lane_s < G;
    const float lim = 1.0e8f;
    const float ux = fminf(fmaxf((qx - g.ox) * g.inv_h, -lim), lim);
    const float uy = fminf(fmaxf((qy - g.oy) * g.inv_h, -lim), lim);
    const float uz = fminf(fmaxf((qz - g.oz) * g.inv_h, -lim), lim);
    const float slack = g.slack + 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
    // phase 1: own row, the nearer y neighbour, the nearer z neighbour, and the diagonal
    const int cy = floor_to_int(uy), cz = floor_to_int(uz);
    const int y1 = (uy - (float)cy >= 0.5f) ? cy + 1 : cy - 1;
    const int z1 = (uz - (float)cz >= 0.5f) ? cz + 1 : cz - 1;
    float gb = tau0;  // group-uniform bound on the final best distance
    {
        const int y = (lig & 1) ? y1 : cy, z = (lig & 2) ? z1 : cz;
        uint32_t s = 0, e = 0;
        if (row_lane && y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
            const float gyz2 = row_gap2(g, uy, uz, slack, y, z);
            const float cov = sqrtf(gb) + m;
            if (gyz2 <= cov * cov) row_run(g, ux, slack, y, z, gyz2, cov * cov, s, e);
        }
        gb = fminf(scan_runs<SL>(g, qx, qy, qz, s, e, bd, bp, sd, lane_s, gmask), tau0);
    }
    // phase 2: the rest of the (tightened) ball's cover -- usually nothing; four rows at a time, same scan
    const float rt = fminf((sqrtf(gb) + m) * g.inv_h + slack, 3.0e8f);
    const int ylo = max(0, floor_to_int(fmaxf(uy - rt, -1.f)));
    const int yhi = min(g.ny - 1, floor_to_int(fminf(uy + rt, (float)g.ny)));
    const int zlo = max(0, floor_to_int(fmaxf(uz - rt, -1.f)));
    const int zhi = min(g.nz - 1, floor_to_int(fminf(uz + rt, (float)g.nz)));
    const int wy = yhi - ylo + 1, wz = zhi - zlo + 1;
    if (wy <= 0 || wz <= 0) return;
    if (ylo >= min(cy, y1) && yhi <= max(cy, y1) && zlo >= min(cz, z1) && zhi <= max(cz, z1)) return;  // inside the 2 x 2 block
    const int nrows = wy * wz;
    for (int base = 0; base < nrows; base += G) {  // group-uniform trip count
        const int r = base + lig;
        uint32_t s = 0, e = 0;
        if (row_lane && r < nrows) {
            const int y = ylo + r % wy, z = zlo + r / wy;
            if (!((y == cy || y == y1) && (z == cz || z == z1))) {  // (those were phase 1)
                const float gyz2 = row_gap2(g, uy, uz, slack, y, z);
                const float cov = sqrtf(gb) + m;  // covered radius from here on (never below the final one)
                if (gyz2 <= cov * cov) row_run(g, ux, slack, y, z, gyz2, cov * cov, s, e);
            }
        }
        gb = fminf(scan_runs<SL>(g, qx, qy, qz, s, e, bd, bp, sd, lane_s, gmask), tau0);
    }
}

}  // namespace
}  // namespace b200
