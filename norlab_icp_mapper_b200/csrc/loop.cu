// loop.cu -- the whole ICP loop of one registration as ONE persistent cooperative kernel (k = 1).
//
// Replaces the body of PM::ICPSequence::operator() (/root/reference/norlab_icp_mapper/Mapper.cpp:213)
// from the second iteration on; iteration 0's cold search is the stand-alone knn_kernel.
//
// The grid stays resident (one 1024-thread CTA per SM); each CTA owns a fixed slice of the reading for
// the whole registration and keeps that slice's MATCH STATE IN SHARED MEMORY across iterations:
// reading point, matched map point (+ its position), its normal, the squared distance, and a proven
// lower bound L on the distance from the query to every OTHER map point.  One iteration is
//
//   V  verify   (thread per query, shared memory only): the query moved by delta since the bound was
//               established; if dist(q, match) + delta < L the old match is still the exact nearest
//               neighbour (triangle inequality) -- no search.  Otherwise the query goes on a work list.
//   S  search   (4 lanes per listed query): exact ball search on the cell-sorted grid, bounded by the old
//               match, widened by a small margin m so that it also yields the 2nd-nearest distance / the
//               covered radius = the next bound L.  After ICP's first iterations only a few percent of
//               the queries need this.
//   C  classify (thread per query): outlier weights, error-minimiser sums, quantile bookkeeping.
//   one device-wide barrier, then every CTA finishes the iteration redundantly and bit-identically:
//   exact quantile limit, fixed-order reduction, 6x6 solve, checkers ("one-barrier iteration" below).
//
// Results are the same as an exhaustive search every iteration: the verification only ever skips work
// whose outcome is proven (nn_variant bit 5 disables it; tests compare).  The general quantile path
// (3 barriers, level-0 histogram through global memory) remains for iterations whose quantile window
// cannot be predicted.  Every CTA reduces the per-CTA partial sums in the same fixed order and runs the
// solve and the checkers redundantly, so T_iter and the stop decision are bit-identical everywhere and
// nothing has to be broadcast; each path is run-to-run deterministic.
#include <cooperative_groups.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cmath>
#include <type_traits>

#include "icp_device.cuh"
#include "knn_device.cuh"

namespace b200 {
namespace {

// development (stamped build only): per-CTA %globaltimer stamps of the LAST iteration, 32 slots per CTA in the `partials`
// buffer (idle on the paths that are being looked at); read back with b200icp_debug_cta_stamps
#ifdef B200ICP_STAMPS
#define B200_CTA_STAMP(buf, k)                                                                   \
    do {                                                                                         \
        if (threadIdx.x == 0) {                                                                  \
            unsigned long long t_;                                                               \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                               \
            reinterpret_cast<unsigned long long*>(buf)[(size_t)blockIdx.x * kAccSlots + (k)] = t_; \
        }                                                                                        \
    } while (0)
#else
#define B200_CTA_STAMP(buf, k) do { } while (0)
#endif

// This file is compiled twice: as loop.cu (chains without a RobustOutlierFilter: its scale selects and weights are compiled out,
// which is worth 2 % of the headline kernel in registers and spills) and, through loop_robust.cu, with B200ICP_LOOP_ROBUST defined
// (launch_icp_loop_robust, to which launch_icp_loop hands the chains that hold one).
#ifdef B200ICP_LOOP_ROBUST
constexpr bool kLoopRobust = true;
#else
constexpr bool kLoopRobust = false;
#endif
constexpr int kLoopThreads = 1024;
constexpr int kLoopWarps = kLoopThreads / 32;
constexpr int kLoopG = 4;                       // lanes per query in the search phase
// reading points are dealt to the CTAs round-robin in chunks of 32 consecutive points, 8 for small readings (balance: with
// 32 a 10 k-point reading gives 17 of the 148 CTAs 96 points and the others 64, and everybody waits for those 17)
constexpr int kChunkShiftLarge = 5, kChunkShiftSmall = 3;
constexpr int kSmallReading = 64 * 1024, kTinyReading = 16 * 1024;
constexpr int kCacheCapMax = 2048;              // entries per CTA whose match state can live in shared memory (the rest spills to global)
// The cache is sized per launch (`cache_cap` = the entries a CTA of this registration owns, rounded up to 32, at most kCacheCapMax):
// the shared memory it does not take stays L1 -- cfg 2 needs 704 of the 2048 entries (55 KB instead of 128 KB dynamic), which
// more than doubles the L1 the search phase's point loads go through (hard variant: loop 1.40 -> 1.35 ms).
// dynamic shared memory: float4 r4[cap] | float4 pp[cap] | float4 nv[cap] | float d2[cap] | uint32 list[1024] | the fine histogram of a
// two-barrier iteration
constexpr size_t kLoopEntryBytes = 3 * sizeof(float4) + sizeof(float);
constexpr size_t kLoopListBytes = (size_t)kLoopThreads * sizeof(uint32_t);  // the search phase's work list (one block of entries)
constexpr size_t kLoopCacheBytes = (size_t)kCacheCapMax * kLoopEntryBytes + kLoopListBytes;
__host__ __device__ constexpr size_t loop_cache_bytes(int cache_cap) { return (size_t)cache_cap * kLoopEntryBytes + kLoopListBytes; }
constexpr size_t kLoopScratchBytes = 4160 * sizeof(uint32_t);
constexpr size_t kLoopDynSmem = kLoopCacheBytes + kLoopScratchBytes;
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
constexpr int kHistDebug = 12288;  // 8 words per iteration (first 256 iterations): development record written by CTA 0
constexpr int kHistStage1 = 16384;  // two more level-0 + fine histograms (kStage1Words each), used alternately by the two-barrier iteration
constexpr int kHistStat = 16000;   // [0] queries that went through the search phase (all CTAs, whole registration)

// ---- one-barrier iteration ("fast path") --------------------------------------------------------------
// Once ICP settles the Trimmed/quantile limit moves by a few percent per iteration, so the limit of the
// previous iteration predicts a narrow WINDOW [win_lo, win_hi] (float bit patterns of dist2) that will
// contain the new limit.  Each CTA then classifies its pairs while it still has them at hand:
//   dist2 <  win_lo : certainly kept    -> error sums accumulated right away (per-CTA partial)
//   dist2 in window : candidate         -> (p, n, dot | q) tuple + dist2 bits appended to the CTA's segment
//   dist2 >  win_hi : certainly dropped -> counted only
// and publishes {counts, sums, candidates} (see below).  After ONE device-wide barrier every CTA reads the three counts,
// checks that the quantile's rank really falls among the candidates, pulls the candidate list (~0.1 % of the pairs) and
// finds the exact limit there (same value the 3-level radix select returns), adds the candidates at or below the limit,
// and solves -- redundantly and bit-identically in every CTA.  If the prediction fails (rank outside the window, or the
// list overflows) all CTAs take the next stage / the general path below for that iteration.
// Buffers are double-buffered on the iteration's parity: a CTA can only be one barrier ahead of another.
constexpr int kCandCap = 2048;   // entries of the candidate list (2 per thread after the barrier)
constexpr int kSmallCand = 128;  // up to here the candidates are finished by four warps with rank counting (the usual case)
// Everything a CTA publishes before the barrier is ORDER-FREE: 64-bit integer accumulators in global memory, added to with
// atomics, and a candidate list appended to with one atomic per warp that holds candidates.
//  * error sums of the pairs whose fate is certain: fixed point (associative: the result does not depend on the arrival
//    order, so poses stay run-to-run bitwise deterministic), instead of 148 per-CTA records every CTA had to read back;
//  * the counts (pairs below / inside / above the quantile's window, list entries) in four more slots of the same array;
//  * candidates (pairs inside the window): tuple (p, dist2 bits | v) at a position handed out by the `inside` counter.  The
//    list order depends on arrival, so after the barrier the kept candidates are summed in fixed point as well (a coarser
//    grid: one 32-bit warp reduction per sum).
// One slot per 128-byte line (the atomics and the read-back of different slots go to different L2 slices), three buffers
// used round-robin (see the zeroing protocol at the barrier); the candidate list is double-buffered on the round's parity.
constexpr int kIsumStride = 16;  // unsigned long long per slot line
constexpr int kIsumBufs = 3;
constexpr int kSlotFlag = kAccSlots - 1;  // a sum left the fixed-point range somewhere
constexpr int kSlotBelow = kAccSlots, kSlotAbove = kAccSlots + 1, kSlotCand = kAccSlots + 2, kSlotRank = kAccSlots + 3;
constexpr int kIsumSlots = kAccSlots + 4;
constexpr size_t kFastCandOff = 0;                                                          // float4 [2][kCandCap][2]
constexpr size_t kFastCandEnd = kFastCandOff + 2 * (size_t)kCandCap * 2 * sizeof(float4);
// (Robust's windowed median select, loop_window_median: key lists [2][kRselCap] before, counters [3][4 slots][16] after the accumulators)
constexpr int kRselCap = 4096;
constexpr size_t kFastRselListOff = (kFastCandEnd + 127) / 128 * 128;                       // uint32 [2][kRselCap]
constexpr size_t kFastIsumOff = kFastRselListOff + 2 * (size_t)kRselCap * sizeof(uint32_t);  // unsigned long long [3][kIsumSlots][16]
constexpr size_t kFastRselCntOff = kFastIsumOff + (size_t)kIsumBufs * kIsumSlots * kIsumStride * sizeof(unsigned long long);  // ull [3][4][16]
constexpr size_t kFastBytes = kFastRselCntOff + (size_t)kIsumBufs * 4 * kIsumStride * sizeof(unsigned long long);
// Two-barrier iteration: next to the level-0 histogram (bits [30:19]) a FINE histogram (64 bins per level-0 bucket, bits
// [18:13]) over the 65 level-0 buckets around the previous limit (+- 2 octaves): when the quantile's bucket is among them
// the window is one fine bin -- a few dozen candidates instead of a thousand -- and the four-warp finish applies.
constexpr int kFineSpan = 32, kFineBuckets = 2 * kFineSpan + 1, kFineShift = 13, kFineSub = 64;
constexpr int kFineBins = kFineBuckets * kFineSub;  // 4160
constexpr int kStage1Words = kSel0Bins + kFineBins;
static_assert((size_t)kFineBins * sizeof(uint32_t) <= kLoopScratchBytes, "the scratch region of the dynamic shared memory holds the fine histogram");

// log2 of the chunk of consecutive reading points the CTAs are dealt (host: sizes the match cache; device: the dealing itself)
__host__ __device__ inline int loop_chunk_shift(long long nq, int variant_flags) {
    int cshift = (nq <= kSmallReading || (variant_flags & 512)) ? kChunkShiftSmall : kChunkShiftLarge;
    if (nq <= kTinyReading) cshift = 0;  // plain round-robin (cfg 4, 10 k points: another -3.6 %)
    if ((variant_flags >> 10) & 3) cshift = ((variant_flags >> 10) & 3) == 3 ? 4 : ((variant_flags >> 10) & 3) - 1;  // development: chunks of 1 / 2 / 16
    return cshift;
}

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Monotonic counter barrier: the k-th barrier completes when the counter reaches k * gridDim.x.  Arrival is one
// release-reduction (no round trip for a returned value; cumulativity orders the CTA's earlier writes and atomics, which
// thread 0 observed through the __syncthreads, before it), the wait polls with acquire loads.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        const unsigned target = epoch * gridDim.x;
        while (ld_acquire_u32(counter) < target) {
        }
    }
    __syncthreads();
}

// s_cnt32[w] = count of warp w (written before the preceding __syncthreads): returns the sum over the warps before `warp`
// and the block total.  One shared load + a 5-step shuffle scan per thread.
__device__ __forceinline__ void block_prefix32(const uint32_t* s_cnt32, int lane, int warp, uint32_t& before, uint32_t& total) {
    const uint32_t v = s_cnt32[lane];
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += u;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    before = __shfl_sync(0xffffffffu, incl - v, warp);
}

// Block-wide (kLoopThreads threads): smallest bin with cumulative count > rank (see icp.cu select_pick).
// kShared: the histogram lives in shared memory (plain loads) instead of global memory (L2 loads).
template <bool kShared>
__device__ uint32_t loop_pick(const uint32_t* hist, int nbins, uint32_t rank, bool rank_is_fraction, float q, uint32_t* s_bin,
                              uint32_t* s_res, uint32_t* s_cnt, uint32_t* s_warp, bool rank_is_half = false) {
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
    block_prefix32(s_warp, lane, warp, before, total);
    incl += before;
    if (rank_is_fraction) {
        rank = (q == 1.0f) ? (total ? total - 1u : 0u) : (uint32_t)((float)total * q);
        if (total && rank >= total) rank = total - 1u;
    }
    if (rank_is_half) rank = total >> 1;  // element n / 2 of the sorted values (Matches::getMedianAbsDeviation)
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

// ---- RobustOutlierFilter's scale inside the loop kernel -----------------------------------------------
// What the kernel-per-step path gets from two device-wide sorts (outlier.cu launch_robust_scale): the element of rank n / 2
// among the n finite match distances (mode 0), or among their absolute deviations from `centre` (mode 1).  Exact radix
// select over the keys' float bit patterns, every CTA in lockstep and with the same result: level 0 (bits [30:19]) of each
// CTA's slice -> `h0` in global memory -> barrier -> bucket of the rank; when that bucket holds <= kSelListCap keys they are
// gathered into one list (second barrier) and every CTA finishes locally, else two more global levels (bits [18:8], [7:0]).
// Zeroing protocol: a region is zeroed by CTA 0 after a barrier every CTA has passed having finished reading it, and written
// again only after a later barrier, which CTA 0 reaches after the zeroing; the level-2 region is handed over to the caller
// (`l2_dirty`: zeroed after the next barrier).  Returns false when there is no finite distance (buffers left clean).
struct LoopSelCtx {
    const float4* s_pp;   // match cache: (x, y, z, bit-cast position) ...
    const float* s_d2;    // ... and squared distance of the entries below cache_cap
    const float4* sp_pp;  // the same for the spilled entries (global memory, indexed by match row)
    const float* md2;
    uint32_t *sh, *sh2, *s_bin, *s_res, *s_cnt, *s_stage, *s_base, *s_warp;
    uint32_t* hist;
    unsigned* bar_counter;
    int n_ent, K, cshift, cache_cap;
};

__device__ __noinline__ bool loop_exact_median(const LoopSelCtx& c, int mode, float centre, uint32_t* h0, unsigned& epoch, bool& l2_dirty,
                                               uint32_t& out_bits) {
    const int tid = threadIdx.x;
    auto key_of = [&](int e) -> uint32_t {  // 0xffffffff: not a finite match distance
        int pos;
        float d;
        if (e < c.cache_cap) {
            pos = __float_as_int(c.s_pp[e].w);
            d = c.s_d2[e];
        } else {
            const int ql = c.K == 1 ? e : e / c.K;
            const long long qi = (((long long)(ql >> c.cshift) * gridDim.x + blockIdx.x) << c.cshift) + (ql & ((1 << c.cshift) - 1));
            const long long pi = c.K == 1 ? qi : qi * c.K + (e - ql * c.K);
            pos = __float_as_int(c.sp_pp[pi].w);
            d = c.md2[pi];
        }
        if (!(pos >= 0 && d < CUDART_INF_F)) return 0xffffffffu;
        return __float_as_uint(mode ? fabsf(d - centre) : d);
    };
    for (int i = tid; i < kSel0Bins; i += kLoopThreads) c.sh[i] = 0u;
    if (tid == 0) {
        *c.s_bin = 0;
        *c.s_res = 0;
        *c.s_cnt = 0;
        *c.s_stage = 0;
    }
    __syncthreads();
    for (int e = tid; e < c.n_ent; e += kLoopThreads) {
        const uint32_t k = key_of(e);
        if (k != 0xffffffffu) atomicAdd(&c.sh[k >> kSel0Shift], 1u);
    }
    __syncthreads();
    for (int i = tid; i < kSel0Bins; i += kLoopThreads)
        if (c.sh[i]) atomicAdd(&h0[i], c.sh[i]);
    grid_barrier(c.bar_counter, epoch);
    if (l2_dirty) {  // (an earlier select's last level: everybody has read it by now)
        if (blockIdx.x == 0)
            for (int i = tid; i < 256; i += kLoopThreads) c.hist[kHistL2 + i] = 0u;
        l2_dirty = false;
    }
    const uint32_t total = loop_pick<false>(h0, kSel0Bins, 0u, false, 0.f, c.s_bin, c.s_res, c.s_cnt, c.s_warp, /*rank_is_half=*/true);
    const uint32_t b1 = *c.s_bin, r1 = *c.s_res, c1 = *c.s_cnt;
    __syncthreads();
    if (total == 0) {
        grid_barrier(c.bar_counter, epoch);
        if (blockIdx.x == 0)
            for (int i = tid; i < kSel0Bins; i += kLoopThreads) h0[i] = 0u;
        return false;
    }
    uint32_t low = 0;  // the 19 low bits
    if (c1 <= (uint32_t)kSelListCap) {
        for (int e = tid; e < c.n_ent; e += kLoopThreads) {
            const uint32_t k = key_of(e);
            if (k != 0xffffffffu && (k >> kSel0Shift) == b1) c.sh[atomicAdd(c.s_stage, 1u)] = k;
        }
        __syncthreads();
        if (tid == 0) *c.s_base = *c.s_stage ? atomicAdd(&c.hist[kHistCount], *c.s_stage) : 0u;
        __syncthreads();
        for (uint32_t i = tid; i < *c.s_stage; i += kLoopThreads) c.hist[kHistList + *c.s_base + i] = c.sh[i];
        grid_barrier(c.bar_counter, epoch);
        if (blockIdx.x == 0) {
            for (int i = tid; i < kSel0Bins; i += kLoopThreads) h0[i] = 0u;
            if (tid == 0) c.hist[kHistCount] = 0u;
        }
        for (uint32_t i = tid; i < c1; i += kLoopThreads) c.sh[i] = __ldcg(c.hist + kHistList + i);
        c.sh2[tid] = 0u;  // local level 1: bits [18:9]
        if (tid == 0) {
            *c.s_bin = 0;
            *c.s_res = 0;
        }
        __syncthreads();
        for (uint32_t i = tid; i < c1; i += kLoopThreads) atomicAdd(&c.sh2[(c.sh[i] >> 9) & 1023u], 1u);
        __syncthreads();
        loop_pick<true>(c.sh2, 1024, r1, false, 0.f, c.s_bin, c.s_res, c.s_cnt, c.s_warp);
        const uint32_t b2 = *c.s_bin, r2 = *c.s_res;
        __syncthreads();
        if (tid < 512) c.sh2[tid] = 0u;  // local level 2: bits [8:0]
        if (tid == 0) {
            *c.s_bin = 0;
            *c.s_res = 0;
        }
        __syncthreads();
        for (uint32_t i = tid; i < c1; i += kLoopThreads)
            if (((c.sh[i] >> 9) & 1023u) == b2) atomicAdd(&c.sh2[c.sh[i] & 511u], 1u);
        __syncthreads();
        loop_pick<true>(c.sh2, 512, r2, false, 0.f, c.s_bin, c.s_res, c.s_cnt, c.s_warp);
        low = (b2 << 9) | *c.s_bin;
        __syncthreads();
    } else {
#ifdef B200ICP_STAMPS
        if (blockIdx.x == 0 && tid == 0) printf("loop_exact_median: three-level select (mode %d, bucket of %u keys)\n", mode, c1);
#endif
        uint32_t rank = r1, pre = b1;
        for (int pass = 1; pass <= 2; ++pass) {
            const int nbins = (pass == 1) ? 2048 : 256;
            uint32_t* gh = c.hist + (pass == 1 ? kHistL1 : kHistL2);
            for (int i = tid; i < nbins; i += kLoopThreads) c.sh[i] = 0u;
            if (tid == 0) {
                *c.s_bin = 0;
                *c.s_res = 0;
            }
            __syncthreads();
            for (int e = tid; e < c.n_ent; e += kLoopThreads) {
                const uint32_t k = key_of(e);
                if (k == 0xffffffffu) continue;
                if (pass == 1) {
                    if ((k >> kSel0Shift) == pre) atomicAdd(&c.sh[(k >> 8) & 2047u], 1u);
                } else {
                    if ((k >> 8) == pre) atomicAdd(&c.sh[k & 255u], 1u);
                }
            }
            __syncthreads();
            for (int i = tid; i < nbins; i += kLoopThreads)
                if (c.sh[i]) atomicAdd(&gh[i], c.sh[i]);
            grid_barrier(c.bar_counter, epoch);
            if (blockIdx.x == 0) {  // the previous level is no longer read by anyone
                if (pass == 1)
                    for (int i = tid; i < kSel0Bins; i += kLoopThreads) h0[i] = 0u;
                else
                    for (int i = tid; i < 2048; i += kLoopThreads) c.hist[kHistL1 + i] = 0u;
            }
            loop_pick<false>(gh, nbins, rank, false, 0.f, c.s_bin, c.s_res, c.s_cnt, c.s_warp);
            rank = *c.s_res;
            pre = (pass == 1) ? ((pre << 11) | *c.s_bin) : ((pre << 8) | *c.s_bin);
            __syncthreads();
        }
        l2_dirty = true;
        low = pre & ((1u << kSel0Shift) - 1u);
    }
    out_bits = (b1 << kSel0Shift) | low;
    return true;
}

// The same order statistic with ONE barrier when the previous iterations predict where it lies: keys below / above the window
// [wlo, whi] (float bit patterns) are only counted, the keys inside are appended to a global list; after the barrier every CTA
// reads the three counts, and -- if the rank n / 2 falls inside the window and the list did not overflow -- pulls the list and
// selects locally.  Returns false when the prediction failed (the caller runs loop_exact_median).  `cnt` = this select's counter
// set {below, inside, above} (one 128-byte line each), `cnt_stale` the set of the select before the previous one, zeroed here by
// CTA 0 (three sets round-robin, as for the error sums); `list` is double-buffered on the select's parity.
__device__ __noinline__ bool loop_window_median(const LoopSelCtx& c, int mode, float centre, uint32_t wlo, uint32_t whi, unsigned long long* cnt,
                                                unsigned long long* cnt_stale, uint32_t* list, unsigned& epoch, uint32_t& out_bits) {
    const int tid = threadIdx.x, lane = tid & 31;
    auto key_of = [&](int e) -> uint32_t {
        int pos;
        float d;
        if (e < c.cache_cap) {
            pos = __float_as_int(c.s_pp[e].w);
            d = c.s_d2[e];
        } else {
            const int ql = c.K == 1 ? e : e / c.K;
            const long long qi = (((long long)(ql >> c.cshift) * gridDim.x + blockIdx.x) << c.cshift) + (ql & ((1 << c.cshift) - 1));
            const long long pi = c.K == 1 ? qi : qi * c.K + (e - ql * c.K);
            pos = __float_as_int(c.sp_pp[pi].w);
            d = c.md2[pi];
        }
        if (!(pos >= 0 && d < CUDART_INF_F)) return 0xffffffffu;
        return __float_as_uint(mode ? fabsf(d - centre) : d);
    };
    if (tid == 0) {
        *c.s_res = 0;  // this CTA's keys below the window ...
        *c.s_cnt = 0;  // ... and above it
    }
    __syncthreads();
    uint32_t nb = 0, na = 0;
    for (int e0 = 0; e0 < c.n_ent; e0 += kLoopThreads) {
        const int e = e0 + tid;
        const uint32_t k = e < c.n_ent ? key_of(e) : 0xffffffffu;
        const bool counted = k != 0xffffffffu, inside = counted && k >= wlo && k <= whi;
        nb += counted && k < wlo;
        na += counted && k > whi;
        const unsigned bal = __ballot_sync(0xffffffffu, inside);
        if (bal) {  // (warp-uniform)
            unsigned long long base = 0ull;
            if (lane == 0) base = atomicAdd(cnt + 1 * kIsumStride, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned long long slot = base + (unsigned long long)__popc(bal & ((1u << lane) - 1u));
            if (inside && slot < (unsigned long long)kRselCap) __stcg(list + slot, k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
        na += __shfl_xor_sync(0xffffffffu, na, o);
    }
    if (lane == 0) {
        if (nb) atomicAdd(c.s_res, nb);
        if (na) atomicAdd(c.s_cnt, na);
    }
    __syncthreads();
    if (tid == 0) {
        if (*c.s_res) atomicAdd(cnt + 0 * kIsumStride, (unsigned long long)*c.s_res);
    } else if (tid == 32) {
        if (*c.s_cnt) atomicAdd(cnt + 2 * kIsumStride, (unsigned long long)*c.s_cnt);
    }
    grid_barrier(c.bar_counter, epoch);
    if (blockIdx.x == 0 && tid < 3) cnt_stale[tid * kIsumStride] = 0ull;
    const uint32_t n_below = (uint32_t)umin64(__ldcg(cnt + 0 * kIsumStride), 0xffffffffull);
    const uint32_t n_in = (uint32_t)umin64(__ldcg(cnt + 1 * kIsumStride), 0xffffffffull);
    const uint32_t n_above = (uint32_t)umin64(__ldcg(cnt + 2 * kIsumStride), 0xffffffffull);
    const uint32_t total = n_below + n_in + n_above, rank = total >> 1;
    if (total == 0 || n_in > (uint32_t)kRselCap || rank < n_below || rank >= n_below + n_in) return false;  // (the same verdict in every CTA)
    for (uint32_t i = tid; i < n_in; i += kLoopThreads) c.sh[i] = __ldcg(list + i);
    const uint32_t W = whi - wlo;  // < 2^30 (window construction)
    uint32_t r = rank - n_below, prefix = 0;
#pragma unroll 1
    for (int shift = 20; shift >= 0; shift -= 10) {
        if (shift > 0 && (W >> shift) == 0u) continue;
        c.sh2[tid] = 0u;
        if (tid == 0) {
            *c.s_bin = 0;
            *c.s_res = 0;
        }
        __syncthreads();
        for (uint32_t i = tid; i < n_in; i += kLoopThreads) {
            const uint32_t o = c.sh[i] - wlo;
            if (((o ^ prefix) >> (shift + 10)) == 0u) atomicAdd(&c.sh2[(o >> shift) & 1023u], 1u);
        }
        __syncthreads();
        loop_pick<true>(c.sh2, 1024, r, false, 0.f, c.s_bin, c.s_res, c.s_cnt, c.s_warp);
        r = *c.s_res;
        prefix |= *c.s_bin << shift;
        __syncthreads();
    }
    out_bits = wlo + prefix;
    return true;
}

// Product of the weights of every outlier filter EXCEPT the quantile-based one (fast path: that one is
// decided after the barrier).
__device__ __forceinline__ float other_filters_weight(const IcpParams& prm, float d, const float* T, const float4* rn, const float4& fn) {
    float w = 1.f;
    for (int f = 0; f < prm.n_outlier; ++f) {
        const float p = prm.outlier_param[f];
        bool keep = true;
        switch (prm.outlier_kind[f]) {
            case B200ICP_OUTLIER_MAX_DIST: keep = d <= p * p; break;
            case B200ICP_OUTLIER_MIN_DIST: keep = d >= p * p; break;
            case B200ICP_OUTLIER_SURFACE_NORMAL:
                if (rn) keep = surface_normal_keep(T, __ldg(rn), fn, cosf(p));  // (rn is null when either cloud has no normals)
                break;
            default: break;
        }
        w *= keep ? 1.f : 0.f;
    }
    return w;
}

// RobustOutlierFilter's weight of a CANDIDATE pair, recomputed after the barrier from its tuple (ta = (p, dist2 bits), tb = v):
// point2point distances are the tuple's own bits; point2plane (point-to-plane minimiser only: v = (n, (p - q) . n)) divides the
// residual by |n|.  Out of line: the finish paths are short of registers and this runs for a few dozen pairs per iteration.
__device__ __noinline__ float robust_candidate_weight(int mode, float tuning, float approximation, float scale, float4 ta, float4 tb) {
    float dist = ta.w;
    if ((mode >> 12) & 1) {
        const float dot = tb.w / sqrtf(tb.x * tb.x + tb.y * tb.y + tb.z * tb.z);
        dist = dot * dot;
    }
    return robust_weight(mode, tuning, approximation, scale, dist);
}

// The error-minimiser products of one kept pair (same expressions as accumulate_entry).
// MIN 0: v = (n.x, n.y, n.z, (p - q).n);  MIN 1: v = (q.x, q.y, q.z, *).
template <int MIN>
__device__ __forceinline__ void add_pair(float* acc, float w, const float3& p, const float4& v) {
    constexpr int NS = SumLayout<MIN>::N;
    if (MIN == 0) {
        float F[6];
        F[0] = p.y * v.z - p.z * v.y;
        F[1] = p.z * v.x - p.x * v.z;
        F[2] = p.x * v.y - p.y * v.x;
        F[3] = v.x;
        F[4] = v.y;
        F[5] = v.z;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float wf = w * F[c];
#pragma unroll
            for (int r = 0; r <= c; ++r) acc[c * (c + 1) / 2 + r] += wf * F[r];
            acc[21 + c] -= wf * v.w;
        }
    } else if (MIN == 1) {
        const float pv[3] = {p.x, p.y, p.z}, qv[3] = {v.x, v.y, v.z};
        acc[0] += w;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc[1 + c] += w * pv[c];
            acc[4 + c] += w * qv[c];
#pragma unroll
            for (int r = 0; r < 3; ++r) acc[7 + c * 3 + r] += w * qv[r] * pv[c];
        }
    }
    acc[NS - 2] += w;
    acc[NS - 1] += 1.f;
}

template <int MIN, int GK /* lanes per query of the k > 1 search (8, 16, 32); 4 stands for k = 1 */>
__global__ void __launch_bounds__(kLoopThreads, 1)
    icp_loop_kernel(IcpParams prm, GridView g, const float4* __restrict__ nrm, const float4* __restrict__ reading,
                    int32_t* __restrict__ mpos, float* __restrict__ md2, IcpState* __restrict__ gst, uint32_t* __restrict__ hist,
                    double* __restrict__ partials, unsigned* __restrict__ bar_counter, float* __restrict__ trace, int max_iters,
                    int variant_flags, char* __restrict__ fastws, float win_gain, float win_floor, float win_max,
                    float4* __restrict__ sp_pp, float4* __restrict__ sp_nv, float margin_gain, float margin_min, float margin_max, float q_bound,
                    int cache_cap) {
    constexpr int NS = SumLayout<MIN>::N;
    const int kCacheCap = cache_cap;  // (a launch parameter: see kCacheCapMax)
    __shared__ IcpState st;
    __shared__ uint32_t sh[kSel0Bins];   // radix level 0 histogram, then staging / the candidate list
    __shared__ __align__(16) uint32_t sh2[1024];  // local radix levels / the candidates' keys
    __shared__ uint32_t s_warp[kLoopWarps + 1], s_warp2[kLoopWarps];
    __shared__ uint32_t s_bin, s_res, s_cnt, s_stage, s_base;
    __shared__ double s_part[kLoopWarps][NS];
    __shared__ double s_red[kLoopWarps][kAccSlots];
    __shared__ double s_sum[kAccSlots];
    __shared__ float s_scratch[16];
    __shared__ unsigned long long s_isum[kIsumSlots];  // fast path: the published accumulators, read back after the barrier
    __shared__ uint32_t s_tot[4];                   // fast path: this CTA's counts {below, -, above, -}
    __shared__ uint32_t s_limit_bits;               // fast path: dist2 bits of the quantile
    __shared__ float s_Tprev[16];                   // T_iter the bounds L of the match cache refer to
    __shared__ uint32_t s_nlist;
    __shared__ double s_scale[kAccSlots], s_inv_scale[kAccSlots];  // fixed-point scale of each error sum (see kFastIsumOff)
    __shared__ float s_cscale[kAccSlots];                           // ... and the coarser grid of the candidates' products
    __shared__ double s_inv_cscale[kAccSlots];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float4* const s_r4 = reinterpret_cast<float4*>(dyn_smem);
    float4* const s_pp = s_r4 + kCacheCap;   // (x, y, z, bit-cast position) of the matched map point
    float4* const s_nv = s_pp + kCacheCap;   // (normal of the matched point, bound L)
    float* const s_d2 = reinterpret_cast<float*>(s_nv + kCacheCap);
    uint32_t* const s_list = reinterpret_cast<uint32_t*>(s_d2 + kCacheCap);
    unsigned char* const s_big = dyn_smem + loop_cache_bytes(cache_cap);
    uint32_t* const s_fine = reinterpret_cast<uint32_t*>(s_big);    // two-barrier iteration: fine histogram (kFineBins words)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(IcpState) / 4); i += kLoopThreads)
        reinterpret_cast<uint32_t*>(&st)[i] = reinterpret_cast<const uint32_t*>(gst)[i];
    __syncthreads();
    const int nq = st.nq;
    if (tid < kAccSlots) {
        // A-priori magnitude of every sum: |reading point| <= Pb, |map point| <= Qb, |p - q| <= Db, normals nominally unit.  The
        // scale puts (pairs x bound) at 2^50 of the 2^63 range: 4096 x headroom for motion and non-unit normals, and a
        // resolution of 2^-50 of the worst case -- far below the fp32 rounding of the per-warp sums that are added.
        float Db = sqrtf(prm.max_r2);  // inf when the matcher is unbounded
        float Pb = sqrtf(__uint_as_float(st.pmax2_bits)) + 1.f;
        if (Db < 3.0e38f) Pb = fminf(Pb, q_bound + Db);  // a PAIRED reading point lies within maxDist of a map point
        else Db = Pb + q_bound;
        float bound = 1.f;
        if (MIN == 0) {
            if (tid < 21) {
                int c = 0;
                while ((c + 1) * (c + 2) / 2 <= tid) ++c;
                const int r = tid - c * (c + 1) / 2;
                bound = (c < 3 ? Pb : 1.f) * (r < 3 ? Pb : 1.f);
            } else if (tid < 27) {
                bound = (tid - 21 < 3 ? Pb : 1.f) * Db;
            }
        } else if (MIN == 1) {
            if (tid >= 1 && tid <= 3) bound = Pb;
            else if (tid >= 4 && tid <= 6) bound = q_bound;
            else if (tid >= 7 && tid <= 15) bound = Pb * q_bound;
        }
        int e2 = 0;
        frexp((double)nq * (double)prm.knn * fmax((double)bound, 1e-30), &e2);
        const int shift = max(-60, min(60, 50 - e2));
        s_scale[tid] = ldexp(1.0, shift);
        s_inv_scale[tid] = ldexp(1.0, -shift);
        int eb = 0;
        frexp(fmax((double)bound, 1e-30), &eb);  // bound <= 2^eb: one product lands below 2^20 nominally, 2^23 is the accepted limit
        const int cshift = max(-100, min(100, 20 - eb));
        s_cscale[tid] = (float)ldexp(1.0, cshift);
        s_inv_cscale[tid] = ldexp(1.0, -cshift);
    }
    unsigned epoch = 0;
    uint32_t n_runs = 0, n_hist = 0;  // publish/finish rounds (stage-1 histograms) so far: parity selects the double buffer
    const bool use_quantile = prm.quantile_filter >= 0;
    int robust_f = -1;  // the chain's RobustOutlierFilter (at most one); compiled in only in loop_robust.cu's copy of this kernel
    if (kLoopRobust)
        for (int f = 0; f < prm.n_outlier; ++f)
            if (prm.outlier_kind[f] == B200ICP_OUTLIER_ROBUST) robust_f = f;
    bool l2_dirty = false;  // the last level of a three-level median select waits for its zeroing (loop_exact_median)
    // Robust's two medians (of the distances, of their absolute deviations): the last value and the window predicted for the next
    // iteration, the same in every CTA; n_rsel counts the windowed selects (counter set / list buffer in use)
    // (shared memory, written by thread 0: registers are what this kernel is short of)
    __shared__ float rsel_prev[2];
    __shared__ uint32_t rsel_lo[2], rsel_hi[2], rsel_have[2], rsel_valid[2], n_rsel;
    if (tid < 2) {
        rsel_prev[tid] = 0.f;
        rsel_lo[tid] = rsel_hi[tid] = rsel_have[tid] = rsel_valid[tid] = 0u;
        n_rsel = 0u;
    }
    const bool sn_active = prm.rnrm != nullptr && nrm != nullptr;  // SurfaceNormalOutlierFilter has what it needs (else: all ones)
    const int lig = lane & (kLoopG - 1);
    const unsigned gmask = group_mask<kLoopG>(lane);
    const int cshift = loop_chunk_shift(nq, variant_flags);
    const int kChunk = 1 << cshift;
    // This CTA's slice: n_ql reading points (chunks of kChunk consecutive points, dealt round-robin: only the globally last
    // chunk is partial, so local query ql <-> reading point qi_of(ql) is contiguous) and K entries each -- one per neighbour.
    // Entry e = ql * K + j <-> row `pair_of(e)` of the match arrays (ids[i * knn + j] layout).  K = 1: entries are the points.
    const int K = prm.knn;
    int n_ql = 0;
    {
        const long long chunks_total = ((long long)nq + kChunk - 1) >> cshift;
        if ((long long)blockIdx.x < chunks_total) {
            const long long mine = (chunks_total - 1 - blockIdx.x) / gridDim.x + 1;
            n_ql = (int)(mine * kChunk);
            if ((long long)blockIdx.x + (mine - 1) * gridDim.x == chunks_total - 1) n_ql -= (int)(chunks_total * kChunk - nq);
        }
    }
    const int n_ent = n_ql * K;
    auto qi_of = [&](int ql) -> long long { return (((long long)(ql >> cshift) * gridDim.x + blockIdx.x) << cshift) + (ql & (kChunk - 1)); };
    auto pair_of = [&](int e) -> long long { return K == 1 ? qi_of(e) : qi_of(e / K) * K + (e % K); };
    // match cache <- the cold search's matches (bound L = 0: nothing proven yet)
    for (int e = tid; e < n_ent; e += kLoopThreads) {
        const long long pi = pair_of(e);
        const int pos = mpos[pi];
        float4 pt = make_float4(0.f, 0.f, 0.f, 0.f), nn = pt;
        if (pos >= 0) {
            pt = __ldg(g.pts + pos);
            if (MIN == 0 || sn_active) nn = __ldg(nrm + pos);
        }
        pt.w = __int_as_float(pos);
        nn.w = 0.f;
        if (e < kCacheCap) {
            s_pp[e] = pt;
            s_nv[e] = nn;
            s_d2[e] = md2[pi];
        } else {
            sp_pp[pi] = pt;
            sp_nv[pi] = nn;
        }
    }
    for (int ql = tid; ql < n_ql && ql < kCacheCap; ql += kLoopThreads) s_r4[ql] = __ldg(reading + qi_of(ql));
    if (tid < 16) s_Tprev[tid] = st.T[tid];
    __syncthreads();

    for (int it = 0; it < max_iters; ++it) {
        if (st.done) break;  // identical in every CTA
        const bool stamper = blockIdx.x == 0 && tid == 0;
        if (stamper) B200_STAMP(gst, 20);
        B200_CTA_STAMP(partials, 0);
        const bool searched = st.iter > 0;  // iteration 0's matches come from the cold kernel
        unsigned long long t_iter0 = 0;
        if (blockIdx.x == 0 && tid == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_iter0));
            if (it == 0) st.loop_total_ns = t_iter0;  // start mark, turned into a duration at the end
        }
        // ---- V / S over this CTA's entries, 1024 at a time (one block unless nq > 148 * 1024) ------------------------
        const bool verify = !(variant_flags & 32);
        unsigned long long t_search = 0;  // CTA 0, thread 0: time spent in V + S
        if (GK > 4) {
            // k > 1, by blocks of 1024 reading points.  V (thread per point): the cached k matches are still THE k nearest
            // when the farthest of them is closer to the moved query than the bound L (kept in the first entry's s_nv.w)
            // allows any other map point to be.  S (GK lanes per listed point): shell search with one sorted list per group
            // (AccK, knn_device.cuh), pruned from the start by the largest distance to the cached matches -- k distinct
            // real map points -- and widened by the margin so that the new bound L leaves room for the next motions.
            const int ligk = lane & (GK - 1);
            const unsigned gmaskk = group_mask<GK>(lane);
            for (int q0 = 0; searched && q0 < n_ql; q0 += kLoopThreads) {
                unsigned long long t_s0 = 0;
                if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_s0));
                const int ql = q0 + tid;
                bool listed = false;
                if (ql < n_ql) {
                    const long long qq = qi_of(ql);
                    const float4 r4 = ql < kCacheCap ? s_r4[ql] : __ldg(reading + qq);
                    const float3 qn = apply_T(st.T, r4), qo = apply_T(s_Tprev, r4);
                    const float ddx = qn.x - qo.x, ddy = qn.y - qo.y, ddz = qn.z - qo.z;
                    const float delta = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                    const int e0 = ql * K;
                    float maxd = 0.f;
                    int cnt = 0;
                    bool inside = true;
                    for (int j = 0; j < K; ++j) {
                        const int e = e0 + j;
                        const bool cs = e < kCacheCap;
                        const float4 pp = cs ? s_pp[e] : sp_pp[qq * K + j];
                        const int pos = __float_as_int(pp.w);
                        float d2n = CUDART_INF_F;
                        if (pos >= 0) {
                            d2n = dist2_exact(qn.x, qn.y, qn.z, pp);
                            inside = inside && (d2n <= prm.max_r2);  // (false for a NaN query too)
                            maxd = fmaxf(maxd, d2n);
                            cnt += 1;
                        }
                        if (cs) s_d2[e] = d2n;
                        else md2[qq * K + j] = d2n;
                    }
                    float* pL = e0 < kCacheCap ? &s_nv[e0].w : &sp_nv[qq * K].w;
                    const float L = *pL;
                    // fewer than k matches: every other map point was beyond maxDist and has to stay there
                    const float need = cnt == K ? sqrtf(maxd) : sqrtf(prm.max_r2);
                    const float slack = 2e-6f * (need + delta + L);
                    const float Lnew = L - delta - slack;
                    if (verify && inside && need + slack < Lnew) *pL = Lnew;
                    else listed = true;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, listed);
                if (lane == 0) s_warp[warp] = (uint32_t)__popc(bal);
                if (tid == 0) s_nlist = 0u;  // cursor of the S phase
                __syncthreads();
                uint32_t before = 0, n_list = 0;
                block_prefix32(s_warp, lane, warp, before, n_list);
                if (listed) s_list[before + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)ql;
                __syncthreads();
                if (tid == 0 && n_list) atomicAdd(&hist[kHistStat], n_list);
                while (true) {  // batches of 32 / GK consecutive list entries, handed to whole warps dynamically
                    uint32_t i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&s_nlist, 32u / GK);
                    i0 = __shfl_sync(0xffffffffu, i0, 0);
                    if (i0 >= n_list) break;  // warp-uniform
                    const uint32_t i = i0 + (uint32_t)(lane / GK);
                    if (i < n_list) {
                        const int qs = (int)s_list[i];
                        const long long qq = qi_of(qs);
                        const float4 r4 = qs < kCacheCap ? s_r4[qs] : __ldg(reading + qq);
                        const float3 qn = apply_T(st.T, r4), qo = apply_T(s_Tprev, r4);
                        const float ddx = qn.x - qo.x, ddy = qn.y - qo.y, ddz = qn.z - qo.z;
                        const float delta = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                        const float want = margin_gain * delta;
                        const float m = want <= margin_max ? fmaxf(want, margin_min) : 0.f;
                        const int e = qs * K + ligk;
                        const bool mine = ligk < K, cs = e < kCacheCap;
                        const long long pi = qq * K + ligk;
                        float dj = 0.f;
                        if (mine) {
                            const float4 pp = cs ? s_pp[e] : sp_pp[pi];
                            dj = __float_as_int(pp.w) >= 0 ? dist2_exact(qn.x, qn.y, qn.z, pp) : CUDART_INF_F;
                        }
#pragma unroll
                        for (int o = GK / 2; o > 0; o >>= 1) dj = fmaxf(dj, __shfl_xor_sync(gmaskk, dj, o));
                        float bound = CUDART_INF_F;
                        if (dj < CUDART_INF_F) bound = __uint_as_float(__float_as_uint(dj) + 1u);  // next float up: ties with the bound stay accepted
                        bool collected = false;
                        if (dj <= prm.max_r2 && !(variant_flags & 128) && (variant_flags & 0x800000)) {
                            // EXPERIMENT (nn_variant bit 23, off by default: measured slower -- cfg 4 loop 0.96 -> 1.07 ms):
                            // ONE pass over the ball the cached matches bound, collecting what lies inside it (AccCollect: no
                            // cross-lane traffic while scanning), then the k nearest by rank counting
                            constexpr int CS = 4;
                            AccCollect<GK, CS> col;
                            col.init(K, bound, m);
                            search_ball_k<GK, AccCollect<GK, CS>>(g, col, qn.x, qn.y, qn.z, prm.max_r2, ligk, gmaskk);
                            if (!__any_sync(gmaskk, col.ovf)) {  // (a lane out of slots: the sorted-list search below redoes the query)
                                int rank[CS];
                                const int total = col.select(rank, ligk, gmaskk);
                                float rej = col.sd;  // smallest distance tested and not kept: refused at the gate, or ranked k or worse
#pragma unroll
                                for (int c = 0; c < CS; ++c)
                                    if (c < col.cnt && rank[c] >= K) rej = fminf(rej, col.cd[c]);
#pragma unroll
                                for (int o = GK / 2; o > 0; o >>= 1) rej = fminf(rej, __shfl_xor_sync(gmaskk, rej, o));
                                // every point within sqrt(min(bound, maxDist^2)) + m was tested: the others are at least L away
                                float L = fminf(sqrtf(rej), sqrtf(fminf(bound, prm.max_r2)) + m);
                                L = fminf(L - 2e-6f * L, 1.0e18f);
                                auto put = [&](int r, int op, float od) {  // cache entry of rank r
                                    const int e2 = qs * K + r;
                                    const long long pi2 = qq * K + r;
                                    float4 pt = make_float4(0.f, 0.f, 0.f, 0.f), nn = pt;
                                    if (op >= 0) {
                                        pt = __ldg(g.pts + op);
                                        if (MIN == 0 || sn_active) nn = __ldg(nrm + op);
                                    }
                                    pt.w = __int_as_float(op);
                                    nn.w = L;  // (read from the first entry only)
                                    if (e2 < kCacheCap) {
                                        s_pp[e2] = pt;
                                        s_nv[e2] = nn;
                                        s_d2[e2] = od;
                                    } else {
                                        sp_pp[pi2] = pt;
                                        sp_nv[pi2] = nn;
                                        md2[pi2] = od;
                                    }
                                };
#pragma unroll
                                for (int c = 0; c < CS; ++c)
                                    if (c < col.cnt && rank[c] < K) put(rank[c], col.cp[c], col.cd[c]);
                                if (ligk >= total && ligk < K) put(ligk, -1, CUDART_INF_F);  // fewer than k points within maxDist
                                collected = true;
                            }
                        }
                        if (!collected) {
                        AccK<GK, true> acc;
                        acc.init(K, bound, m);
                        // all k cached matches still within maxDist: the ball they bound is searched in one pass
                        if (dj <= prm.max_r2 && !(variant_flags & 128))
                            search_ball_k<GK, AccK<GK, true>>(g, acc, qn.x, qn.y, qn.z, prm.max_r2, ligk, gmaskk);
                        else
                            search_shells<GK, AccK<GK, true>>(g, acc, qn.x, qn.y, qn.z, prm.max_r2, 0, ligk, gmaskk);
                        float gsd = acc.sd;
#pragma unroll
                        for (int o = GK / 2; o > 0; o >>= 1) gsd = fminf(gsd, __shfl_xor_sync(gmaskk, gsd, o));
                        if (mine) {
                            const float od = acc.d;
                            const int op = acc.pos;
                            float4 pt = make_float4(0.f, 0.f, 0.f, 0.f), nn = pt;
                            if (op >= 0) {
                                pt = __ldg(g.pts + op);
                                if (MIN == 0 || sn_active) nn = __ldg(nrm + op);
                            }
                            pt.w = __int_as_float(op);
                            // every point within sqrt(min(k-th, maxDist^2)) + m was tested: the others are at least L away
                            float L = fminf(sqrtf(gsd), sqrtf(fminf(acc.kth, prm.max_r2)) + m);
                            L = fminf(L - 2e-6f * L, 1.0e18f);
                            nn.w = L;  // (read from the first entry only)
                            if (cs) {
                                s_pp[e] = pt;
                                s_nv[e] = nn;
                                s_d2[e] = od;
                            } else {
                                sp_pp[pi] = pt;
                                sp_nv[pi] = nn;
                                md2[pi] = od;
                            }
                        }
                        }
                    }
                    __syncwarp();
                }
                __syncthreads();  // the search results of this block are in the cache (and s_list is free again)
                if (blockIdx.x == 0 && tid == 0) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    t_search += t1 - t_s0;
                }
            }
        } else
        for (int e0 = 0; e0 < n_ent; e0 += kLoopThreads) {
            const int e = e0 + tid;
            const bool have = e < n_ent;
            const bool cached = e < kCacheCap;
            const long long qi = have ? qi_of(e) : 0;
            bool listed = false;
            int cls = 0;
            unsigned long long t_s0 = 0;
            if (searched) {
                if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_s0));
                // V: is the previous match still provably the nearest neighbour?  (shared memory only)
                if (have) {
                    const float4 r4 = cached ? s_r4[e] : __ldg(reading + qi);
                    float4* ppp = cached ? s_pp + e : sp_pp + qi;
                    float4* pnv = cached ? s_nv + e : sp_nv + qi;
                    float* pd2 = cached ? s_d2 + e : md2 + qi;
                    const float3 qn = apply_T(st.T, r4), qo = apply_T(s_Tprev, r4);
                    const float4 pp = *ppp;
                    const int pos = __float_as_int(pp.w);
                    const float L = pnv->w;
                    const bool finite_q = (fabsf(qn.x) < 3.0e38f) && (fabsf(qn.y) < 3.0e38f) && (fabsf(qn.z) < 3.0e38f);
                    if (!finite_q) {  // NaN / inf reading point: never matched
                        ppp->w = __int_as_float(-1);
                        *pd2 = CUDART_INF_F;
                    } else if (!(pos < 0 && prm.max_r2 == CUDART_INF_F)) {  // (unbounded search that found nothing: the map is empty)
                        const float ddx = qn.x - qo.x, ddy = qn.y - qo.y, ddz = qn.z - qo.z;
                        const float delta = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                        const float d2n = pos >= 0 ? dist2_exact(qn.x, qn.y, qn.z, pp) : CUDART_INF_F;
                        const float dn = pos >= 0 ? sqrtf(d2n) : sqrtf(prm.max_r2);
                        const float slack = 2e-6f * (dn + delta + L);
                        const float Lnew = L - delta - slack;  // still a lower bound for every other point, now around the moved query
                        if (verify && (pos < 0 || d2n <= prm.max_r2) && dn + slack < Lnew) {
                            pnv->w = Lnew;
                            *pd2 = d2n;  // inf when there is (still) no neighbour within maxDist
                        } else {
                            listed = true;
                            // cost class by ball radius in cells: beyond the 2 x 2 rows of the search's first phase (> 0.5) the
                            // search gets long; 0 = cheapest
                            const float rc = dn * g.inv_h;
                            cls = (variant_flags & 128) ? 0 : (!(rc <= 1.0f) ? 3 : (rc > 0.5f ? 2 : (rc > 0.3f ? 1 : 0)));
                        }
                    }
                }
                // ordered work list (neighbouring groups get neighbouring queries), longest searches first (four cost classes by
                // ball radius): warps take batches dynamically, so the expensive ones must not be the last to start
                unsigned bal[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) bal[c] = __ballot_sync(0xffffffffu, listed && cls == c);
                if (lane == 0) {  // two packed prefix sums: (class 3 | class 2 << 16), (class 1 | class 0 << 16); <= 1024 entries in all
                    s_warp[warp] = (uint32_t)__popc(bal[3]) | ((uint32_t)__popc(bal[2]) << 16);
                    s_warp2[warp] = (uint32_t)__popc(bal[1]) | ((uint32_t)__popc(bal[0]) << 16);
                }
                if (tid == 0) s_nlist = 0u;  // cursor of the S phase
                __syncthreads();
                uint32_t before_hi = 0, total_hi = 0, before_lo = 0, total_lo = 0;
                block_prefix32(s_warp, lane, warp, before_hi, total_hi);
                block_prefix32(s_warp2, lane, warp, before_lo, total_lo);
                const uint32_t n3 = total_hi & 0xffffu, n2 = total_hi >> 16, n1 = total_lo & 0xffffu, n0 = total_lo >> 16;
                const uint32_t n_list = n3 + n2 + n1 + n0;
                if (listed) {
                    const unsigned lt = (1u << lane) - 1u;
                    uint32_t slot;
                    if (cls == 3) slot = (before_hi & 0xffffu) + __popc(bal[3] & lt);
                    else if (cls == 2) slot = n3 + (before_hi >> 16) + __popc(bal[2] & lt);
                    else if (cls == 1) slot = n3 + n2 + (before_lo & 0xffffu) + __popc(bal[1] & lt);
                    else slot = n3 + n2 + n1 + (before_lo >> 16) + __popc(bal[0] & lt);
                    s_list[slot] = (uint32_t)e;
                }
                __syncthreads();
                if (tid == 0 && n_list) atomicAdd(&hist[kHistStat], n_list);
                if (stamper) B200_STAMP(gst, 17);
                // S: exact ball search for the listed queries, kLoopG lanes each.  The 8 groups of a warp walk the list in
                // lockstep (warp-uniform trip count, reconvergence every round): a warp then issues the LONGEST of its
                // groups' instruction paths per round, not their sum -- this phase is issue-bound, not memory-bound.
                // Batches of 8 consecutive list entries are handed to whole warps dynamically (difficulty is spatially
                // coherent, a static deal leaves the CTA waiting for its unluckiest warp).
                // one listed query, searched by SL lanes (lane_s = this lane's index among them, smask = their mask)
                auto search_entry = [&](auto sl_tag, int es, int lane_s, unsigned smask) {
                    constexpr int SL = decltype(sl_tag)::value;
                    const bool cs = es < kCacheCap;
                    const long long qs = qi_of(es);
                    const float4 r4 = cs ? s_r4[es] : __ldg(reading + qs);
                    float4* ppp = cs ? s_pp + es : sp_pp + qs;
                    float4* pnv = cs ? s_nv + es : sp_nv + qs;
                    float* pd2 = cs ? s_d2 + es : md2 + qs;
                    const float3 qn = apply_T(st.T, r4), qo = apply_T(s_Tprev, r4);
                    const float4 pp = *ppp;
                    const int pos = __float_as_int(pp.w);
                    const float ddx = qn.x - qo.x, ddy = qn.y - qo.y, ddz = qn.z - qo.z;
                    const float delta = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                    const float d2n = pos >= 0 ? dist2_exact(qn.x, qn.y, qn.z, pp) : CUDART_INF_F;
                    const bool valid_prev = pos >= 0 && d2n <= prm.max_r2;
                    const float tau0 = valid_prev ? d2n : prm.max_r2;  // finite: V never lists (no previous match, unbounded maxDist)
                    // margin: worth paying only when the next motion is likely to stay inside it
                    const float want = margin_gain * delta;
                    const float m = want <= margin_max ? fmaxf(want, margin_min) : 0.f;
                    float bd = CUDART_INF_F, sd = CUDART_INF_F;
                    int bp = -1;
                    search_ball4<SL>(g, qn.x, qn.y, qn.z, tau0, m, bd, bp, sd, lane_s, smask);
                    float gbd = bd;
                    int gbp = bp;
#pragma unroll
                    for (int o = SL / 2; o > 0; o >>= 1) {
                        const float od = __shfl_xor_sync(smask, gbd, o);
                        const int op = __shfl_xor_sync(smask, gbp, o);
                        if (od < gbd || (od == gbd && (unsigned)op < (unsigned)gbp)) {
                            gbd = od;
                            gbp = op;
                        }
                    }
                    float gsd = (bp != gbp) ? bd : sd;  // a lane whose best lost is looking at another point (and bd <= sd)
#pragma unroll
                    for (int o = SL / 2; o > 0; o >>= 1) gsd = fminf(gsd, __shfl_xor_sync(smask, gsd, o));
                    const float cover = sqrtf(fminf(gbd, tau0)) + m;  // every lane covered at least this radius
                    if (!(gbd <= prm.max_r2)) {  // nothing within maxDist: whatever was seen beyond it is an "other" point
                        gsd = fminf(gsd, gbd);
                        gbd = CUDART_INF_F;
                        gbp = -1;
                    }
                    if (lane_s == 0) {
                        float L = fminf(sqrtf(gsd), cover);
                        L -= 2e-6f * L;
                        if (gbp < 0 || gbp == pos) {
                            ppp->w = __int_as_float(gbp);
                            pnv->w = L;
                        } else {  // new match: its coordinates and normal go into the cache now
                            float4 pt = __ldg(g.pts + gbp), nn = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (MIN == 0 || sn_active) nn = __ldg(nrm + gbp);
                            pt.w = __int_as_float(gbp);
                            nn.w = L;
                            *ppp = pt;
                            *pnv = nn;
                        }
                        *pd2 = gbd;
                    }
                };
                if (n_list <= (uint32_t)kLoopWarps && !(variant_flags & 0x400000)) {
                    // a handful of queries (converged iterations): one per warp, all 32 lanes scanning -- their latency is what the
                    // other 147 CTAs wait for at the barrier
                    if ((uint32_t)warp < n_list) search_entry(std::integral_constant<int, 32>(), (int)s_list[warp], lane, 0xffffffffu);
                } else
                while (true) {
                    uint32_t i0 = 0;
                    if (lane == 0) i0 = atomicAdd(&s_nlist, 32u / kLoopG);
                    i0 = __shfl_sync(0xffffffffu, i0, 0);
                    if (i0 >= n_list) break;  // warp-uniform
                    const uint32_t i = i0 + (uint32_t)(lane / kLoopG);
                    if (i < n_list) search_entry(std::integral_constant<int, kLoopG>(), (int)s_list[i], lig, gmask);
                    __syncwarp();
                }
                if (stamper) B200_STAMP(gst, 18);
                if (blockIdx.x == 0 && tid == 0) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    t_search += t1 - t_s0;
                }
            }
            if (searched) {
                __syncthreads();  // the search results of this block are in the cache
                if (stamper) B200_STAMP(gst, 19);
            }
        }
        if (stamper) B200_STAMP(gst, 30);
        B200_CTA_STAMP(partials, 1);
        if (searched && tid < 16) s_Tprev[tid] = st.T[tid];  // the bounds now refer to this iteration's query positions
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0 && searched) {
            st.loop_search_ns += t_search;
            st.loop_iters_timed += 1;
        }
        if (stamper) B200_STAMP(gst, 21);
        float qlimit = 0.f;
        bool fast_done = false;
        bool fatal = false;
        // ---- RobustOutlierFilter: this iteration's scale (LPM RobustOutlierFilter::robustFiltering; outlier.cu for the
        //      kernel-per-step path).  mad: sqrt(median |d - median d|), two exact selects of two barriers each; berg: 1.9 sqrt(median d)
        //      in the first iteration, then the recursion; none: nothing to do.  (std stays on the kernel-per-step path.)
        if (robust_f >= 0) {
            const int mode = prm.outlier_mode[robust_f];
            const int estimator = (mode >> 8) & 15, nb = (mode >> 16) & 0x7fff;
            const int iteration = st.iter + 1;  // LPM counts from 1
            if (estimator != B200ICP_SCALE_NONE && (iteration <= nb || nb == 0)) {
                if (estimator == B200ICP_SCALE_BERG && iteration > 1) {
                    if (tid == 0) st.robust_scale = 0.85f * (st.robust_scale - prm.outlier_param[robust_f]) + prm.outlier_param[robust_f];
                } else {
                    LoopSelCtx sc;
                    sc.s_pp = s_pp;
                    sc.s_d2 = s_d2;
                    sc.sp_pp = sp_pp;
                    sc.md2 = md2;
                    sc.sh = sh;
                    sc.sh2 = sh2;
                    sc.s_bin = &s_bin;
                    sc.s_res = &s_res;
                    sc.s_cnt = &s_cnt;
                    sc.s_stage = &s_stage;
                    sc.s_base = &s_base;
                    sc.s_warp = s_warp;
                    sc.hist = hist;
                    sc.bar_counter = bar_counter;
                    sc.n_ent = n_ent;
                    sc.K = K;
                    sc.cshift = cshift;
                    sc.cache_cap = cache_cap;
                    uint32_t med_bits = 0, dev_bits = 0;
                    unsigned long long* const rcnt = reinterpret_cast<unsigned long long*>(fastws + kFastRselCntOff);
                    uint32_t* const rlist = reinterpret_cast<uint32_t*>(fastws + kFastRselListOff);
                    bool have = true;
                    for (int j = 0; j < 2 && have; ++j) {  // 0: median of the distances, 1: of their absolute deviations from it
                        if (j == 1 && estimator != B200ICP_SCALE_MAD) break;
                        uint32_t& bits = j ? dev_bits : med_bits;
                        const float centre = j ? __uint_as_float(med_bits) : 0.f;
                        bool done = false;
                        if (rsel_valid[j] && !(variant_flags & 0x8000000)) {  // one barrier, when the window holds
                            const uint32_t nr = n_rsel;
                            done = loop_window_median(sc, j, centre, rsel_lo[j], rsel_hi[j], rcnt + (size_t)(nr % kIsumBufs) * 4 * kIsumStride,
                                                      rcnt + (size_t)((nr + 2) % kIsumBufs) * 4 * kIsumStride, rlist + (size_t)(nr & 1u) * kRselCap,
                                                      epoch, bits);
                            __syncthreads();
                            if (tid == 0) n_rsel = nr + 1;
                        }
                        if (!done) {
                            uint32_t* h0 = hist;
                            if (j) {  // (the level-0 region may still be being zeroed after the first select)
                                h0 = hist + kHistStage1 + (n_hist & 1u) * kStage1Words;
                                n_hist += 1;
                            }
                            have = loop_exact_median(sc, j, centre, h0, epoch, l2_dirty, bits);
                            if (tid == 0) st.hist_iters += 1;  // (Robust chains: the selects that took two or more barriers)
                        }
                        __syncthreads();
                        if (have && tid == 0) {  // next iteration's window: centred on this value, half-width from its last change
                            const float v = __uint_as_float(bits);
                            rsel_valid[j] = 0u;
                            if (rsel_have[j] && v > 0.f && v < 1.0e30f) {
                                const float a = fmaxf(win_gain * fabsf(v - rsel_prev[j]), 4.f * win_floor * v);
                                if (a <= 2.f * win_max * v) {
                                    rsel_lo[j] = __float_as_uint(fmaxf(v - a, 0.f));
                                    rsel_hi[j] = __float_as_uint(v + a);
                                    rsel_valid[j] = (rsel_hi[j] - rsel_lo[j]) < (1u << 30) ? 1u : 0u;
                                }
                            }
                            rsel_have[j] = 1u;
                            rsel_prev[j] = v;
                        }
                        __syncthreads();
                    }
                    if (!have) {  // LPM: ConvergenceError("no outlier to filter")
                        if (tid == 0) {
                            st.status = B200ICP_ERR_CONVERGENCE;
                            st.done = 1;
                        }
                        __syncthreads();
                        break;
                    }
                    if (tid == 0)
                        st.robust_scale = estimator == B200ICP_SCALE_BERG ? 1.9f * sqrtf(__uint_as_float(med_bits)) : sqrtf(__uint_as_float(dev_bits));
                }
                __syncthreads();
            }
        }
        uint32_t dbg_path = 0, dbg_ncand = 0, dbg_nbelow = 0;  // development record (CTA 0)
        // ---- the rest of the iteration in ONE barrier (stage 0: quantile window predicted from the last limits) or
        //      TWO (stage 1: window = the histogram bin that holds the quantile, found with a histogram pass);
        //      both decisions are identical in every CTA
        for (int stage = 0; stage < 2 && !fast_done; ++stage) {
            uint32_t wlo = 0xffffffffu, whi = 0xffffffffu;
            uint32_t* h1 = nullptr;
            if (use_quantile) {
                if (stage == 0) {
                    if (!searched || !st.win_valid || (variant_flags & (16 | 8))) continue;
                    wlo = st.win_lo;
                    whi = st.win_hi;
                } else {
                    if ((variant_flags & (64 | 8)) || (prm.outlier_kind[prm.quantile_filter] != B200ICP_OUTLIER_TRIMMED_DIST &&
                                                       prm.outlier_kind[prm.quantile_filter] != B200ICP_OUTLIER_MEDIAN_DIST))
                        continue;
                    // level-0 histogram (bits [30:19] of dist2) of this CTA's slice, and the fine one (64 bins per bucket) over the
                    // buckets around the previous limit -> global -> barrier -> bin of the quantile
                    const bool fine_ok = st.have_limit && st.limit > 0.f && st.limit < 1.0e30f && !(variant_flags & 0x20000);
                    const int fb0 = fine_ok ? (int)(__float_as_uint(st.limit) >> kSel0Shift) - kFineSpan : 0;
                    for (int i = tid; i < kSel0Bins; i += kLoopThreads) sh[i] = 0u;
                    if (fine_ok)
                        for (int i = tid; i < kFineBins; i += kLoopThreads) s_fine[i] = 0u;
                    __syncthreads();
                    for (int e = tid; e < n_ent; e += kLoopThreads) {
                        const bool cached = e < kCacheCap;
                        const long long pi = pair_of(e);
                        const int pos = __float_as_int(cached ? s_pp[e].w : sp_pp[pi].w);
                        const float d = cached ? s_d2[e] : md2[pi];
                        if (pos >= 0 && d < CUDART_INF_F) {
                            const uint32_t bits = __float_as_uint(d);
                            const int b0 = (int)(bits >> kSel0Shift);
                            atomicAdd(&sh[b0], 1u);
                            if (fine_ok && b0 >= fb0 && b0 < fb0 + kFineBuckets)
                                atomicAdd(&s_fine[(b0 - fb0) * kFineSub + (int)((bits >> kFineShift) & (kFineSub - 1))], 1u);
                        }
                    }
                    __syncthreads();
                    h1 = hist + kHistStage1 + (n_hist & 1u) * kStage1Words;  // zeroed again by CTA 0 after this stage's second barrier
                    n_hist += 1;
                    for (int i = tid; i < kSel0Bins; i += kLoopThreads)
                        if (sh[i]) atomicAdd(&h1[i], sh[i]);
                    if (fine_ok)
                        for (int i = tid; i < kFineBins; i += kLoopThreads)
                            if (s_fine[i]) atomicAdd(&h1[kSel0Bins + i], s_fine[i]);
                    if (tid == 0) {
                        s_bin = 0;
                        s_res = 0;
                        s_cnt = 0;
                    }
                    if (stamper) B200_STAMP(gst, 24);
                    B200_CTA_STAMP(partials, 2);
                    grid_barrier(bar_counter, epoch);
                    B200_CTA_STAMP(partials, 3);
                    if (stamper) B200_STAMP(gst, 25);
                    const uint32_t total = loop_pick<false>(h1, kSel0Bins, 0u, true, prm.quantile, &s_bin, &s_res, &s_cnt, s_warp);
                    const uint32_t b1 = s_bin, r1 = s_res, c1 = s_cnt;
                    __syncthreads();
                    if (total == 0) {
                        // LPM: ConvergenceError("no outlier to filter"); leave the buffers clean and stop everywhere
                        grid_barrier(bar_counter, epoch);
                        if (blockIdx.x == 0)
                            for (int i = tid; i < kStage1Words; i += kLoopThreads) h1[i] = 0u;
                        if (tid == 0) {
                            st.status = B200ICP_ERR_CONVERGENCE;
                            st.done = 1;
                        }
                        __syncthreads();
                        fatal = true;
                        break;
                    }
                    wlo = b1 << kSel0Shift;
                    whi = wlo | ((1u << kSel0Shift) - 1u);
                    if (fine_ok && (int)b1 >= fb0 && (int)b1 < fb0 + kFineBuckets && c1 > 32u) {
                        // refine: the fine bin of that bucket that holds the quantile (rank r1 inside the bucket)
                        if (tid == 0) {
                            s_bin = 0;
                            s_res = 0;
                            s_cnt = 0;
                        }
                        __syncthreads();
                        loop_pick<false>(h1 + kSel0Bins + ((int)b1 - fb0) * kFineSub, kFineSub, r1, false, 0.f, &s_bin, &s_res, &s_cnt, s_warp);
                        wlo |= s_bin << kFineShift;
                        whi = wlo | ((1u << kFineShift) - 1u);
                        __syncthreads();
                    }
                }
            } else if (stage == 1 || (variant_flags & 16)) {
                continue;
            }
            const int par = (int)(n_runs & 1u);
            unsigned long long* const isum_base = reinterpret_cast<unsigned long long*>(fastws + kFastIsumOff);
            unsigned long long* const my_isum = isum_base + (size_t)(n_runs % kIsumBufs) * kIsumSlots * kIsumStride;
            // (used again in two rounds: every CTA finished reading it before it arrived at THIS round's barrier, and the next
            //  additions to it come after the NEXT round's barrier, which CTA 0 only reaches after this zeroing)
            unsigned long long* const stale_isum = isum_base + (size_t)((n_runs + 2) % kIsumBufs) * kIsumSlots * kIsumStride;
            n_runs += 1;
            float4* my_cand = reinterpret_cast<float4*>(fastws + kFastCandOff) + (size_t)par * kCandCap * 2;
            // Two windows: [wlo, whi] holds the QUANTILE (its rank is counted), [llo, lhi] holds the LIMIT below which a pair is kept.
            // Trimmed: the same window.  Median: limit = factor * median, so the limit's window is the median's scaled (one ulp wider
            // on each side than the rounded products: fp32 multiplication is monotonic).  A pair in either window is a candidate.
            const bool is_median = use_quantile && prm.outlier_kind[prm.quantile_filter] == B200ICP_OUTLIER_MEDIAN_DIST;
            const float mfac = is_median ? prm.outlier_param[prm.quantile_filter] : 1.f;
            uint32_t llo = wlo, lhi = whi;
            if (is_median) {
                const uint32_t a = __float_as_uint(mfac * __uint_as_float(wlo)), b = __float_as_uint(mfac * __uint_as_float(whi));
                llo = a > 0u ? a - 1u : 0u;
                lhi = b < 0x7f800000u ? b + 1u : 0x7f800000u;
            }
            uint32_t c_below = 0, c_above = 0, c_rank = 0;
            if (tid < 4) s_tot[tid] = 0u;
            if (lane < NS) s_part[warp][lane] = 0.0;
            // C: outlier weights + error sums of what is certain + candidate tuples (thread per entry, from the cache)
            for (int e0 = 0; e0 < n_ent; e0 += kLoopThreads) {
                const int e = e0 + tid;
                const bool have = e < n_ent;
                const bool cached = e < kCacheCap;
                const int ql = (GK > 4) ? e / K : e;  // the entry's reading point (k = 1: the entry itself)
                const long long qi = have ? qi_of(ql) : 0;
                const long long pi = (GK > 4) ? qi * K + (e - ql * K) : qi;
                float acc[NS];
#pragma unroll
                for (int i = 0; i < NS; ++i) acc[i] = 0.f;
                bool is_cand = false;
                float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
                if (have) {
                    const float4 pp = cached ? s_pp[e] : sp_pp[pi];
                    const float4 nv = cached ? s_nv[e] : sp_nv[pi];
                    const float d = cached ? s_d2[e] : md2[pi];
                    const int pos = __float_as_int(pp.w);
                    if (pos >= 0 && d < CUDART_INF_F) {
                        const uint32_t bits = __float_as_uint(d);
                        const int cls = bits < llo ? 0 : (bits <= lhi ? 1 : 2);  // kept for sure / candidate for the limit / dropped for sure
                        const bool in_rank = bits >= wlo && bits <= whi;         // candidate for the quantile
                        c_below += bits < wlo;
                        c_above += bits > whi;
                        c_rank += in_rank;
                        if (cls != 2 || in_rank) {
                            const float wo = other_filters_weight(prm, d, st.T, sn_active ? prm.rnrm + qi : nullptr, nv);
                            float3 p = make_float3(CUDART_NAN_F, 0.f, 0.f);
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (wo != 0.f) {
                                p = apply_T(st.T, ql < kCacheCap ? s_r4[ql] : __ldg(reading + qi));
                                if (MIN == 0)
                                    v = make_float4(nv.x, nv.y, nv.z, (prm.min_flags & 1) ? (p.x - pp.x) * nv.x + (p.y - pp.y) * nv.y
                                                                                          : (p.x - pp.x) * nv.x + (p.y - pp.y) * nv.y + (p.z - pp.z) * nv.z);
                                else
                                    v = pp;
                                float w = wo;
                                if (robust_f >= 0) {  // (candidates: recomputed after the barrier, robust_candidate_weight)
                                    const int mode = prm.outlier_mode[robust_f];
                                    float dist = d;
                                    if ((mode >> 12) & 1) {  // point2plane: (n . (p - q))^2, n normalised (the host checked the map has normals)
                                        const float4 nm = (MIN == 0 || sn_active) ? nv : __ldg(nrm + pos);
                                        const float inv = 1.f / sqrtf(nm.x * nm.x + nm.y * nm.y + nm.z * nm.z);
                                        const float dot = (nm.x * inv) * (p.x - pp.x) + (nm.y * inv) * (p.y - pp.y) + (nm.z * inv) * (p.z - pp.z);
                                        dist = dot * dot;
                                    }
                                    w *= robust_weight(mode, prm.outlier_param[robust_f], prm.outlier_param2[robust_f], st.robust_scale, dist);
                                }
                                if (cls == 0 && w != 0.f) add_pair<MIN>(acc, w, p, v);
                                if (w == 0.f) p.x = CUDART_NAN_F;  // (a candidate the robust function drops: counted for the quantile only)
                            }
                            if (cls == 1 || in_rank) {  // a candidate dropped by another filter still takes part in the quantile: p.x = NaN marks it
                                is_cand = true;
                                ta = make_float4(p.x, p.y, p.z, __uint_as_float(bits));
                                tb = v;
                            }
                        }
                    }
                }
                if (use_quantile) {  // the warp's candidates -> the list, at a position the `inside` counter hands out
                    const unsigned bal = __ballot_sync(0xffffffffu, is_cand);
                    if (bal) {  // (warp-uniform)
                        unsigned long long base = 0ull;
                        if (lane == 0) base = atomicAdd(my_isum + kSlotCand * kIsumStride, (unsigned long long)__popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const unsigned long long slot = base + (unsigned long long)__popc(bal & ((1u << lane) - 1u));
                        if (is_cand && slot < (unsigned long long)kCandCap) {
                            __stcg(my_cand + slot * 2, ta);
                            __stcg(my_cand + slot * 2 + 1, tb);
                        }
                    }
                }
                // this block's sums -> the warp's running partial (fp64 from here on)
                float v32[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v32[i] = (i < NS) ? acc[i] : 0.f;
                if (e0 + warp * 32 < n_ent) {  // (warp-uniform: warps past the end of the slice have nothing to add)
                    const float tot = warp_reduce_32slots(v32, lane);
                    if (lane < NS) s_part[warp][lane] += (double)tot;
                }
            }
            // ---- publish, ONE barrier, finish redundantly ---------------------------------------------
            if (stamper) B200_STAMP(gst, 28);
            {
                uint32_t a = c_below, c = c_above, r = c_rank;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    r += __shfl_xor_sync(0xffffffffu, r, o);
                }
                if (lane == 0) {
                    if (a) atomicAdd(&s_tot[0], a);
                    if (c) atomicAdd(&s_tot[2], c);
                    if (r) atomicAdd(&s_tot[1], r);
                }
            }
            __syncthreads();
            if (tid < NS) {
                double v = 0.0;
#pragma unroll
                for (int wv = 0; wv < kLoopWarps; ++wv) v += s_part[wv][tid];
                const double sv = v * s_scale[tid];
                if (!(fabs(sv) < 4.0e18)) atomicAdd(my_isum + kSlotFlag * kIsumStride, 1ull);  // out of range (or NaN): flagged for everyone
                else if (v != 0.0) atomicAdd(my_isum + tid * kIsumStride, (unsigned long long)__double2ll_rn(sv));
            } else if (tid == 32) {
                if (s_tot[0]) atomicAdd(my_isum + kSlotBelow * kIsumStride, (unsigned long long)s_tot[0]);
            } else if (tid == 33) {
                if (s_tot[2]) atomicAdd(my_isum + kSlotAbove * kIsumStride, (unsigned long long)s_tot[2]);
            } else if (tid == 34) {
                if (s_tot[1]) atomicAdd(my_isum + kSlotRank * kIsumStride, (unsigned long long)s_tot[1]);
            }
            if (stamper) B200_STAMP(gst, 22);
            B200_CTA_STAMP(partials, 4);
            grid_barrier(bar_counter, epoch);
            B200_CTA_STAMP(partials, 5);
            if (stamper) B200_STAMP(gst, 23);
            if (l2_dirty) {
                if (blockIdx.x == 0)
                    for (int i = tid; i < 256; i += kLoopThreads) hist[kHistL2 + i] = 0u;
                l2_dirty = false;
            }
            // ---- after the barrier: identical work in every CTA ---------------------------------------
            if (blockIdx.x == 0 && tid < kIsumSlots) stale_isum[tid * kIsumStride] = 0ull;
            // one round trip: the accumulators and, speculatively, the head of the candidate list
            const unsigned long long isum_mine = tid < kIsumSlots ? __ldcg(my_isum + tid * kIsumStride) : 0ull;
            float4 ca[2], cb[2];
            ca[0] = ca[1] = cb[0] = cb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (use_quantile && tid < kSmallCand) {
                ca[0] = __ldcg(my_cand + (size_t)tid * 2);
                cb[0] = __ldcg(my_cand + (size_t)tid * 2 + 1);
            }
            if (tid < kIsumSlots) s_isum[tid] = isum_mine;
            __syncthreads();
            if (stamper) B200_STAMP(gst, 29);
            const uint32_t n_below = (uint32_t)s_isum[kSlotBelow], n_above = (uint32_t)s_isum[kSlotAbove];
            const uint32_t n_cand = (uint32_t)umin64(s_isum[kSlotCand], 0x7fffffffull);  // list entries (either window)
            const uint32_t n_rank = (uint32_t)s_isum[kSlotRank];                          // pairs inside the quantile's window
            const uint32_t total = n_below + n_rank + n_above;
            uint32_t rank = 0;
            bool ok = true;
            if (use_quantile) {
                rank = (prm.quantile == 1.0f) ? (total ? total - 1u : 0u) : (uint32_t)((float)total * prm.quantile);
                if (total && rank >= total) rank = total - 1u;
                ok = total > 0 && n_cand <= (uint32_t)kCandCap && rank >= n_below && rank < n_below + n_rank;
            }
            dbg_path = dbg_path * 10u + (ok ? (stage == 0 ? 1u : 3u) : (stage == 0 ? 2u : 4u));
            dbg_ncand = n_cand;
            dbg_nbelow = n_below;
            // The candidates at or below the limit are summed in FIXED POINT as well (their list order depends on arrival; integer
            // sums do not): each product is rounded to the grid 2^-cshift of its slot -- 2^-20 of the slot's a-priori magnitude
            // bound, a deterministic perturbation of ~0.1 % of the pairs far below the fp32 rounding of the other sums -- so that a
            // warp total is ONE 32-bit REDUX per slot; totals across warps are added as 64-bit integers.
            if (ok && n_cand <= (uint32_t)kSmallCand) {
                // ---- the usual case: <= 128 candidates, one per thread of warps 0..3 ----
                // rank counting: every candidate needs the number of keys below it (`less`) and at or below it (`le`); all 32
                // warps share that work (candidate = tid % 128, an eighth of the keys each)
                const bool have = use_quantile && (uint32_t)tid < n_cand;  // (tid < 128 only)
                const uint32_t mine = __float_as_uint(ca[0].w);
                const bool rank_role = have && mine >= wlo && mine <= whi, sum_role = have && mine >= llo && mine <= lhi;
                if (tid < kSmallCand) {
                    sh2[tid] = rank_role ? mine : 0xffffffffu;  // (padding: neither below nor equal to any distance)
                    sh2[kSmallCand + tid] = 0u;                 // less | le << 16
                }
                __syncthreads();
                if (stamper) B200_STAMP(gst, 1);
                if (use_quantile) {
                    const int c = tid & (kSmallCand - 1), part = tid >> 7;  // 8 parts of 16 keys
                    const uint32_t key = sh2[c];
                    if (key != 0xffffffffu && (uint32_t)(part * 16) < n_cand) {
                        const uint4* keys = reinterpret_cast<const uint4*>(sh2) + part * 4;
                        uint32_t less = 0, le = 0;
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const uint4 o = keys[i4];
                            less += (o.x < key) + (o.y < key) + (o.z < key) + (o.w < key);
                            le += (o.x <= key) + (o.y <= key) + (o.z <= key) + (o.w <= key);
                        }
                        if (less | le) atomicAdd(&sh2[kSmallCand + c], less | (le << 16));
                    }
                }
                __syncthreads();
                if (stamper) B200_STAMP(gst, 2);
                if (tid < kSmallCand) {
                    long long mine_tot = 0ll;
                    if (use_quantile) {
                        const uint32_t r = rank - n_below;
                        const uint32_t packed = sh2[kSmallCand + tid];
                        const uint32_t less = packed & 0xffffu, le = packed >> 16;
                        // the quantile is the key with less <= r < le (every such candidate holds the same bits)
                        if (rank_role && less <= r && r < le) s_limit_bits = mine;
                    }
                    named_bar_sync(1, kSmallCand);
                    if ((uint32_t)(warp * 32) < n_cand && use_quantile) {  // (warp-uniform)
                        // kept: at or below the limit -- the quantile itself (Trimmed) or factor x the median (Median)
                        const uint32_t q_bits = s_limit_bits;
                        const uint32_t limit_bits = is_median ? __float_as_uint(mfac * __uint_as_float(q_bits)) : q_bits;
                        const bool kept = sum_role && mine <= limit_bits && ca[0].x == ca[0].x;
                        float acc2[NS];
#pragma unroll
                        for (int i = 0; i < NS; ++i) acc2[i] = 0.f;
                        if (kept) {
                            float cw = 1.f;
                            if (robust_f >= 0)
                                cw = robust_candidate_weight(prm.outlier_mode[robust_f], prm.outlier_param[robust_f], prm.outlier_param2[robust_f],
                                                             st.robust_scale, make_float4(0.f, 0.f, 0.f, __uint_as_float(mine)), cb[0]);
                            if (cw != 0.f) add_pair<MIN>(acc2, cw, make_float3(ca[0].x, ca[0].y, ca[0].z), cb[0]);
                        }
                        bool range_ok = true;
                        int f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float sv = i < NS ? acc2[i] * s_cscale[i] : 0.f;
                            range_ok = range_ok && (fabsf(sv) < 8388608.f);
                            f[i] = __float2int_rn(sv);
                        }
                        // 31 shuffles for all the slots together (lane l ends with slot l); a REDUX per slot was 2 x slower
                        const int tot = warp_reduce_32slots(f, lane);
                        if (lane < NS) mine_tot = (long long)tot;
                        const bool any_bad = __any_sync(0xffffffffu, !range_ok);
                        if (lane == kSlotFlag) mine_tot = any_bad ? 1ll : 0ll;
                    }
                    reinterpret_cast<long long*>(&s_red[0][0])[warp * kAccSlots + lane] = mine_tot;
                    if (stamper) B200_STAMP(gst, 3);
                    named_bar_sync(1, kSmallCand);
                    if (stamper) B200_STAMP(gst, 4);
                    if (tid < kAccSlots) {
                        // certain pairs + the candidates at or below the limit: integer totals, order-free
                        long long cv = 0ll;
#pragma unroll
                        for (int part = 0; part < kSmallCand / 32; ++part) cv += reinterpret_cast<const long long*>(&s_red[0][0])[part * kAccSlots + tid];
                        const long long iv = (long long)s_isum[tid];
                        s_sum[tid] = (tid == kSlotFlag) ? (double)(iv + cv) : (double)iv * s_inv_scale[tid] + (double)cv * s_inv_cscale[tid];
                    }
                    if (stamper) B200_STAMP(gst, 6);
                }
                __syncthreads();
            } else if (ok) {
                // ---- many candidates (a wide predicted window, or a two-barrier iteration whose bucket could not be refined):
                //      all warps; <= 2 candidates per thread, local radix select over the window ----
                bool have[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t pos = (uint32_t)tid + (uint32_t)j * kLoopThreads;
                    have[j] = pos < n_cand;
                    if (have[j] && pos >= (uint32_t)kSmallCand) {
                        ca[j] = __ldcg(my_cand + (size_t)pos * 2);
                        cb[j] = __ldcg(my_cand + (size_t)pos * 2 + 1);
                    }
                }
                const uint32_t W = whi - wlo;  // < 2^30 (window construction)
                uint32_t r = rank - n_below, prefix = 0;
#pragma unroll 1
                for (int shift = 20; shift >= 0; shift -= 10) {
                    if (shift > 0 && (W >> shift) == 0u) continue;
                    sh2[tid] = 0u;
                    if (tid == 0) {
                        s_bin = 0;
                        s_res = 0;
                    }
                    __syncthreads();
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (have[j] && __float_as_uint(ca[j].w) >= wlo && __float_as_uint(ca[j].w) <= whi) {  // (the quantile's candidates)
                            const uint32_t o = __float_as_uint(ca[j].w) - wlo;
                            if (((o ^ prefix) >> (shift + 10)) == 0u) atomicAdd(&sh2[(o >> shift) & 1023u], 1u);
                        }
                    }
                    __syncthreads();
                    loop_pick<true>(sh2, 1024, r, false, 0.f, &s_bin, &s_res, &s_cnt, s_warp);
                    r = s_res;
                    prefix |= s_bin << shift;
                    __syncthreads();
                }
                const uint32_t q_bits = wlo + prefix;
                if (tid == 0) s_limit_bits = q_bits;
                const uint32_t limit_bits = is_median ? __float_as_uint(mfac * __uint_as_float(q_bits)) : q_bits;
                long long mine_tot = 0ll;
                if ((uint32_t)(warp * 32) < n_cand) {  // (warp-uniform)
                    int f[NS];
#pragma unroll
                    for (int i = 0; i < NS; ++i) f[i] = 0;
                    bool range_ok = true;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (have[j] && __float_as_uint(ca[j].w) >= llo && __float_as_uint(ca[j].w) <= limit_bits && ca[j].x == ca[j].x) {
                            float acc2[NS];
#pragma unroll
                            for (int i = 0; i < NS; ++i) acc2[i] = 0.f;
                            float cw = 1.f;
                            if (robust_f >= 0)
                                cw = robust_candidate_weight(prm.outlier_mode[robust_f], prm.outlier_param[robust_f], prm.outlier_param2[robust_f],
                                                             st.robust_scale, make_float4(0.f, 0.f, 0.f, ca[j].w), cb[j]);
                            if (cw != 0.f) add_pair<MIN>(acc2, cw, make_float3(ca[j].x, ca[j].y, ca[j].z), cb[j]);
#pragma unroll
                            for (int i = 0; i < NS; ++i) {
                                const float sv = acc2[i] * s_cscale[i];
                                range_ok = range_ok && (fabsf(sv) < 8388608.f);
                                f[i] += __float2int_rn(sv);
                            }
                        }
                    }
                    {
                        int f32[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) f32[i] = i < NS ? f[i] : 0;
                        const int tot = warp_reduce_32slots(f32, lane);
                        if (lane < NS) mine_tot = (long long)tot;
                    }
                    const bool any_bad = __any_sync(0xffffffffu, !range_ok);
                    if (lane == kSlotFlag) mine_tot = any_bad ? 1ll : 0ll;
                }
                reinterpret_cast<long long*>(&s_red[0][0])[warp * kAccSlots + lane] = mine_tot;
                __syncthreads();
                if (tid < kAccSlots) {
                    long long cv = 0ll;
                    for (int wv = 0; wv < kLoopWarps; ++wv) cv += reinterpret_cast<const long long*>(&s_red[0][0])[wv * kAccSlots + tid];
                    const long long iv = (long long)s_isum[tid];
                    s_sum[tid] = (tid == kSlotFlag) ? (double)(iv + cv) : (double)iv * s_inv_scale[tid] + (double)cv * s_inv_cscale[tid];
                }
                __syncthreads();
            }
            if (ok) {
                if (use_quantile) qlimit = __uint_as_float(s_limit_bits);
                if (stamper) B200_STAMP(gst, 27);
                if (s_sum[kSlotFlag] != 0.0) {  // a sum left the fixed-point range somewhere: same verdict in every CTA
                    if (tid == 0) {
                        st.status = B200ICP_ERR_NOT_IMPLEMENTED;
                        st.done = 1;
                    }
                    __syncthreads();
                    fatal = true;
                    break;
                }
                if (stamper) B200_STAMP(gst, 13);
                fast_done = true;
                if (tid == 0) {
                    if (stage == 0) st.fast_iters += 1;
                    else st.hist_iters += 1;
                }
            }
            if (h1 && blockIdx.x == 0)  // the stage-1 histograms were read by every CTA before the barrier above
                for (int i = tid; i < kStage1Words; i += kLoopThreads) h1[i] = 0u;
        }
        if (fatal) break;
        if (!fast_done) {
            // general path (window overflow, or forced by nn_variant): it walks mpos / md2 in global memory
            for (int i = tid; i < kSel0Bins; i += kLoopThreads) sh[i] = 0u;
            __syncthreads();
            for (int e = tid; e < n_ent; e += kLoopThreads) {
                const long long pi = pair_of(e);
                const bool cached = e < kCacheCap;
                const int pos = __float_as_int(cached ? s_pp[e].w : sp_pp[pi].w);
                const float d = cached ? s_d2[e] : md2[pi];
                mpos[pi] = pos;
                md2[pi] = d;
                if (use_quantile && pos >= 0 && d < CUDART_INF_F) atomicAdd(&sh[__float_as_uint(d) >> kSel0Shift], 1u);
            }
            __syncthreads();
        }
        // ---- exact quantile of the finite distances (LPM Matches::getDistsQuantile) -------------------
        bool fallback_used = false;
        if (use_quantile && !fast_done) {
            // level 0: histogram of bits [30:19] over this CTA's slice: built by the pre-pass above
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
                for (int e = tid; e < n_ent; e += kLoopThreads) {
                    const uint32_t bits = __float_as_uint(md2[pair_of(e)]);
                    if (bits < 0x7f800000u && (bits >> kSel0Shift) == b1) sh[atomicAdd(&s_stage, 1u)] = bits;
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
                    for (int e = tid; e < n_ent; e += kLoopThreads) {
                        const uint32_t bits = __float_as_uint(md2[pair_of(e)]);
                        if (bits >= 0x7f800000u) continue;
                        if (pass == 1) {
                            if ((bits >> kSel0Shift) == pre) atomicAdd(&sh[(bits >> 8) & 2047u], 1u);
                        } else {
                            if ((bits >> 8) == pre) atomicAdd(&sh[bits & 255u], 1u);
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
        if (!fast_done) {
        float acc[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) acc[i] = 0.f;
        {
            float T[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) T[i] = st.T[i];
            for (int e = tid; e < n_ent; e += kLoopThreads) {
                const long long pi = pair_of(e);
                accumulate_entry<MIN>(acc, prm, T, g, nrm, reading, pi, K, mpos[pi], md2[pi], qlimit, st.robust_scale);
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
        if ((fallback_used || l2_dirty) && blockIdx.x == 0)
            for (int i = tid; i < 256; i += kLoopThreads) hist[kHistL2 + i] = 0u;
        l2_dirty = false;
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
        }  // !fast_done
        if (tid == 32) {  // (warp 1, beside the solve on warp 0)
            {
                // window for the next iteration's one-barrier attempt: centred on this quantile (Trimmed: the limit itself; Median:
                // the median, the limit's window follows from it), half-width from its last change
                const float prev = st.limit;
                st.win_valid = 0;
                if (use_quantile && (prm.outlier_kind[prm.quantile_filter] == B200ICP_OUTLIER_TRIMMED_DIST ||
                                     prm.outlier_kind[prm.quantile_filter] == B200ICP_OUTLIER_MEDIAN_DIST) && st.have_limit &&
                    qlimit > 0.f && qlimit < 1.0e30f) {
                    const float a = fmaxf(win_gain * fabsf(qlimit - prev), win_floor * qlimit);
                    if (a <= win_max * qlimit) {
                        st.win_lo = __float_as_uint(fmaxf(qlimit - a, 0.f));
                        st.win_hi = __float_as_uint(qlimit + a);
                        st.win_valid = (st.win_hi - st.win_lo) < (1u << 30);
                    }
                }
                st.have_limit = use_quantile ? 1 : 0;
                st.limit = qlimit;
            }
        }
        if (tid < 32) {
            finish_warp(prm, &st, s_sum, NS, blockIdx.x == 0 ? trace : nullptr, s_scratch, stamper ? gst : nullptr);
        }
        __syncthreads();
        if (stamper) B200_STAMP(gst, 14);
        B200_CTA_STAMP(partials, 6);
        if (stamper && it < 256 && (variant_flags & 0x80000)) {  // development record (nn_variant bit 19; its L2 round trip is on CTA 0's critical path): {path 0 general / 1 one-barrier / 2 failed attempt, limit, candidates, below, searched so far, next window}
            uint32_t* rec = hist + kHistDebug + it * 8;
            rec[0] = dbg_path;
            rec[1] = __float_as_uint(qlimit);
            rec[2] = dbg_ncand;
            rec[3] = dbg_nbelow;
            rec[4] = __ldcg(hist + kHistStat);
            rec[5] = st.win_valid ? st.win_lo : 0u;
            rec[6] = st.win_valid ? st.win_hi : 0u;
            unsigned long long tn;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn));
            rec[7] = (uint32_t)(tn - t_iter0);
        }
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            st.loop_total_ns = t1 - st.loop_total_ns;
            st.searched_queries = (int)__ldcg(hist + kHistStat);  // every CTA added its share before its last barrier
            hist[kHistStat] = 0u;
        }
        __syncthreads();
        for (int i = tid; i < (int)(sizeof(IcpState) / 4); i += kLoopThreads)
            reinterpret_cast<uint32_t*>(gst)[i] = reinterpret_cast<const uint32_t*>(&st)[i];
    }
}

template <int MIN, int GK>
cudaError_t launch_loop_t(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                          int variant_flags, const float* win3, const float* margin3, int64_t nq, cudaStream_t s) {
    // Function attributes are PER DEVICE: one flag per (instantiation, device ordinal).  Contexts on different GPUs of one
    // process (b200icp_register_batch) each need the opt-in; concurrent contexts may race to set it, which is harmless
    // (same value, idempotent call), hence atomics only to keep the flags themselves well-defined.
    constexpr int kMaxDevices = 64;
    static std::atomic<int> per_sm_of[kMaxDevices];  // 0 = not queried yet, else occupancy + 1
    int dev = 0;
    cudaError_t ed = cudaGetDevice(&dev);
    if (ed != cudaSuccess) return ed;
    const bool cacheable = dev >= 0 && dev < kMaxDevices;
    int per_sm = cacheable ? per_sm_of[dev].load(std::memory_order_acquire) - 1 : -1;
    if (per_sm < 0) {
        cudaError_t ea = cudaFuncSetAttribute(icp_loop_kernel<MIN, GK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLoopDynSmem);
        if (ea != cudaSuccess) return ea;
        int q = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, icp_loop_kernel<MIN, GK>, kLoopThreads, kLoopDynSmem);
        if (e != cudaSuccess) return e;
        per_sm = q;
        if (cacheable) per_sm_of[dev].store(q + 1, std::memory_order_release);
    }
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
    char* fastws = b.fastws;
    float win_gain = win3[0], win_floor = win3[1], win_max = win3[2];
    float4* sp_pp = b.spill_pp;
    float4* sp_nv = b.spill_nv;
    float margin_gain = margin3[0], margin_min = margin3[1], margin_max = margin3[2] * g.view.h;
    // largest norm a map point can have in the (mean-centred) frame of the index: the far corner of the grid's box
    float q_bound = 0.f;
    {
        const float lo[3] = {view.ox, view.oy, view.oz};
        const float ext[3] = {view.nx * view.h, view.ny * view.h, view.nz * view.h};
        double q2 = 0.0;
        for (int d = 0; d < 3; ++d) {
            const double m = std::max(std::fabs((double)lo[d]), std::fabs((double)lo[d] + ext[d]));
            q2 += m * m;
        }
        q_bound = (float)std::sqrt(q2) + 1.f;
    }
    // match cache: as many entries as the busiest CTA owns (chunks dealt round-robin, k entries per point), the rest of the
    // shared memory stays L1; the carve-out hint follows (per function and device, re-set only when the need changes)
    int cache_cap = kCacheCapMax;
    {
        const int cshift = loop_chunk_shift(nq, variant_flags);
        const long long chunks_total = ((long long)nq + (1ll << cshift) - 1) >> cshift;
        const long long per_cta = ((chunks_total + blocks - 1) / blocks) << cshift;
        const long long entries = per_cta * std::max(p.knn, 1);
        cache_cap = (int)std::min<long long>(kCacheCapMax, std::max<long long>(32, (entries + 31) / 32 * 32));
        static const bool full_cache = getenv("B200ICP_FULL_CACHE") != nullptr;  // development: always the largest cache (A/B)
        if (full_cache) cache_cap = kCacheCapMax;
    }
    const size_t dyn_smem = loop_cache_bytes(cache_cap) + kLoopScratchBytes;
    {
        static std::atomic<int> carveout_of[kMaxDevices];  // last hint + 1
        cudaFuncAttributes fa;
        static std::atomic<int> static_smem{-1};
        int ss = static_smem.load(std::memory_order_acquire);
        if (ss < 0) {
            cudaError_t e = cudaFuncGetAttributes(&fa, icp_loop_kernel<MIN, GK>);
            if (e != cudaSuccess) return e;
            ss = (int)fa.sharedSizeBytes;
            static_smem.store(ss, std::memory_order_release);
        }
        const int pct = (int)std::min<size_t>(100, ((size_t)ss + dyn_smem + 1024) * 100 / (228 * 1024) + 1);
        if (!cacheable || carveout_of[dev].load(std::memory_order_acquire) != pct + 1) {
            cudaError_t e = cudaFuncSetAttribute(icp_loop_kernel<MIN, GK>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e != cudaSuccess) return e;
            if (cacheable) carveout_of[dev].store(pct + 1, std::memory_order_release);
        }
    }
    void* args[] = {&prm, &view, &nrm, &reading, &mpos, &md2, &st, &hist, &partials, &bar_counter, &trace, &max_iters, &variant_flags,
                    &fastws, &win_gain, &win_floor, &win_max, &sp_pp, &sp_nv, &margin_gain, &margin_min, &margin_max, &q_bound, &cache_cap};
    return cudaLaunchCooperativeKernel((void*)icp_loop_kernel<MIN, GK>, dim3(blocks), dim3(kLoopThreads), args, dyn_smem, s);
}

}  // namespace

#ifndef B200ICP_LOOP_ROBUST
size_t icp_loop_workspace_bytes() { return kFastBytes; }
void icp_loop_workspace_zero_range(size_t* offset, size_t* bytes) {
    *offset = kFastIsumOff;
    *bytes = kFastBytes - kFastIsumOff;
}
#endif

template <int GK>
cudaError_t launch_loop_g(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                          int variant, const float* win3, const float* margin3, int64_t nq, cudaStream_t s) {
    if (p.minimizer == B200ICP_MIN_POINT_TO_PLANE) return launch_loop_t<0, GK>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
    if (p.minimizer == B200ICP_MIN_POINT_TO_POINT) return launch_loop_t<1, GK>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
    return launch_loop_t<2, GK>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
}

#ifdef B200ICP_LOOP_ROBUST
cudaError_t launch_icp_loop_robust(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                                   int variant, const float* win3, const float* margin3, int64_t nq, cudaStream_t s) {
#else
cudaError_t launch_icp_loop(const IcpParams& p, const GridIndex& g, IcpBuffers& b, unsigned* bar_counter, int max_iters, int n_sms,
                            int variant, const float* win3, const float* margin3, int64_t nq, cudaStream_t s) {
    for (int f = 0; f < p.n_outlier; ++f)
        if (p.outlier_kind[f] == B200ICP_OUTLIER_ROBUST)
            return launch_icp_loop_robust(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
#endif
    if (p.knn < 1 || p.knn > 32 || !b.fastws || !b.spill_pp || !b.spill_nv) return cudaErrorInvalidValue;
    if (p.knn == 1) return launch_loop_g<4>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
    // k > 1: lanes per query >= k (lane j holds the j-th best); a reading too small to fill the SMs gets more lanes per query
    int lanes = p.knn <= 8 ? 8 : (p.knn <= 16 ? 16 : 32);
    while (lanes < 32 && nq * lanes <= (int64_t)n_sms * kLoopThreads) lanes *= 2;
    if (lanes == 8) return launch_loop_g<8>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
    if (lanes == 16) return launch_loop_g<16>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
    return launch_loop_g<32>(p, g, b, bar_counter, max_iters, n_sms, variant, win3, margin3, nq, s);
}

}  // namespace b200
