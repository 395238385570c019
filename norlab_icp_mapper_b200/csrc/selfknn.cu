// selfknn.cu -- self k-NN of a cell-sorted cloud with TMA-staged candidate tiles.
//
// The neighbour search of SurfaceNormalDataPointsFilter{knn} over the whole local map (the `post:` chain of
// /root/reference/examples/config.yaml:26-27, applied at /root/reference/norlab_icp_mapper/Map.cpp:523-525): the queries ARE
// the cell-sorted map points, so every point of a tile of cells has the same candidate set -- the tile plus a one-cell halo
// -- and staging that set in shared memory pays: it is read once from L2 / HBM and reused by every query of the tile
// (~200 queries x ~400 candidates for a surface map), where the per-query shell walk of knn.cu re-reads it per query.
//
// One CTA per tile of 8 x 4 x 4 cells.  Points are sorted by linear cell id with x fastest, so the cells [x0 - 1, x0 + 8] of
// one (y, z) row of the haloed region are ONE contiguous run of float4: <= 36 runs per tile, each moved by one bulk
// asynchronous copy (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier -- the TMA engine does the
// address generation, no register staging, the threads only wait).  Regions larger than a stage (dense 2-D maps, walls)
// go through the two stages alternately, the next chunk's copies in flight while the current one is scanned.
// Every thread then owns one query and scans the staged candidates -- all threads read the same candidate at the same
// time, a shared-memory broadcast -- keeping its k best (distance, position) pairs in registers.  The result is exact when
// the k-th distance stays within the halo (one cell edge); the few queries in sparse regions where it does not are
// listed and finished by the shell walk of knn.cu.  Distances are bit-identical to knn.cu's (same fma order); equal
// distances are ordered by cell-sorted position.
#include "knn_device.cuh"

namespace b200 {
namespace {

constexpr int kTX = 8, kTY = 4, kTZ = 4;        // tile, in cells
constexpr int kRegionRows = (kTY + 2) * (kTZ + 2);
constexpr int kTileThreads = 256;
constexpr int kStagePts = 1280;                 // points per stage (20 KB); two stages
constexpr int kMaxSegs = 96;                    // runs (split at stage boundaries) a tile can be cut into; more -> shell walk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
// bulk asynchronous copy global -> shared (the 1-D form of the TMA): bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct Seg {
    uint32_t src;   // first point (cell-sorted position)
    uint32_t len;   // points
    uint32_t dst;   // offset inside its stage
    uint32_t row;   // region row (y, z) the run belongs to
};

template <int KMAX>
__global__ void __launch_bounds__(kTileThreads) selfknn_tile_kernel(GridView g, int k, int32_t* __restrict__ out_ids, float* __restrict__ out_d2,
                                                                    uint32_t* __restrict__ fb_list, unsigned* __restrict__ fb_count,
                                                                    unsigned fb_capacity) {
    __shared__ __align__(128) float4 s_buf[2][kStagePts];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_rs[kRegionRows], s_re[kRegionRows];  // haloed region: run of every (y, z) row
    __shared__ uint32_t s_qs[kTY * kTZ], s_qoff[kTY * kTZ + 1];  // the tile's own points: run start and exclusive prefix of the run lengths
    __shared__ Seg s_seg[kMaxSegs];
    __shared__ uint32_t s_chunk_first[kMaxSegs + 1];
    __shared__ int s_nseg, s_nchunk;

    const int tid = threadIdx.x;
    const int tiles_x = (g.nx + kTX - 1) / kTX, tiles_y = (g.ny + kTY - 1) / kTY;
    const int t = blockIdx.x;
    const int tx0 = (t % tiles_x) * kTX, ty0 = ((t / tiles_x) % tiles_y) * kTY, tz0 = (t / (tiles_x * tiles_y)) * kTZ;

    if (tid < kRegionRows) {
        const int y = ty0 - 1 + tid % (kTY + 2), z = tz0 - 1 + tid / (kTY + 2);
        uint32_t s = 0, e = 0;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
            const uint32_t* row = g.cell_start + ((size_t)z * g.ny + y) * (size_t)g.nx;
            s = __ldg(row + max(tx0 - 1, 0));
            e = __ldg(row + min(tx0 + kTX + 1, g.nx));
            const int iy = y - ty0, iz = z - tz0;
            if (iy >= 0 && iy < kTY && iz >= 0 && iz < kTZ) {
                const uint32_t qs = __ldg(row + tx0), qe = __ldg(row + min(tx0 + kTX, g.nx));
                s_qs[iz * kTY + iy] = qs;
                s_qoff[iz * kTY + iy] = qe - qs;  // (lengths for now)
            }
        } else {
            const int iy = y - ty0, iz = z - tz0;
            if (iy >= 0 && iy < kTY && iz >= 0 && iz < kTZ) {
                s_qs[iz * kTY + iy] = 0u;
                s_qoff[iz * kTY + iy] = 0u;
            }
        }
        s_rs[tid] = s;
        s_re[tid] = e;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int r = 0; r < kTY * kTZ; ++r) {
            const uint32_t len = s_qoff[r];
            s_qoff[r] = acc;
            acc += len;
        }
        s_qoff[kTY * kTZ] = acc;
        // cut the region's runs into stage-sized chunks (nothing to do for an empty tile)
        int nseg = 0, chunk = 0;
        uint32_t fill = 0;
        s_chunk_first[0] = 0;
        bool overflow = false;
        for (int r = 0; r < kRegionRows && !overflow && acc > 0; ++r) {
            uint32_t s = s_rs[r];
            const uint32_t e = s_re[r];
            while (s < e) {
                if (fill == (uint32_t)kStagePts) {
                    ++chunk;
                    s_chunk_first[chunk] = (uint32_t)nseg;
                    fill = 0;
                }
                if (nseg == kMaxSegs) {
                    overflow = true;
                    break;
                }
                const uint32_t len = min(e - s, (uint32_t)kStagePts - fill);
                s_seg[nseg++] = Seg{s, len, fill, (uint32_t)r};
                fill += len;
                s += len;
            }
        }
        s_chunk_first[chunk + 1] = (uint32_t)nseg;
        s_nseg = overflow ? -1 : nseg;
        s_nchunk = chunk + 1;
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t nQ = s_qoff[kTY * kTZ];
    if (nQ == 0) return;  // (an empty tile: most of them, the map is a surface)
    const int nseg = s_nseg, nchunk = s_nchunk;
    if (nseg < 0) {  // a region cut into too many pieces (cannot happen with <= 36 rows unless it is huge): everything to the shell walk
        for (uint32_t q = tid; q < nQ; q += kTileThreads) {
            int r = 0;
            while (s_qoff[r + 1] <= q) ++r;
            const unsigned slot = atomicAdd(fb_count, 1u);
            if (slot < fb_capacity) fb_list[slot] = s_qs[r] + (q - s_qoff[r]);
        }
        return;
    }
    // region box in grid units (cells): a query's distance to its faces bounds what the staged set can prove
    const float bx0 = (float)(tx0 - 1), bx1 = (float)(tx0 + kTX + 1), by0 = (float)(ty0 - 1), by1 = (float)(ty0 + kTY + 1);
    const float bz0 = (float)(tz0 - 1), bz1 = (float)(tz0 + kTZ + 1);
    unsigned phase[2] = {0u, 0u};
    auto issue = [&](int c) {
        // warp 0: one lane arms the barrier with the chunk's byte count, then the lanes issue the copies
        if (tid < 32) {
            const uint32_t f = s_chunk_first[c], l = s_chunk_first[c + 1];
            if (tid == 0) {
                uint32_t pts = 0;
                for (uint32_t i = f; i < l; ++i) pts += s_seg[i].len;
                mbar_expect_tx(&s_bar[c & 1], pts * (uint32_t)sizeof(float4));
            }
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was read through the generic proxy until the last barrier
            for (uint32_t i = f + (uint32_t)tid; i < l; i += 32u) {
                const Seg sg = s_seg[i];
                tma_load_1d(&s_buf[c & 1][sg.dst], g.pts + sg.src, sg.len * (uint32_t)sizeof(float4), &s_bar[c & 1]);
            }
        }
    };
    for (uint32_t q0 = 0; q0 < nQ; q0 += kTileThreads) {  // (one round unless the tile holds more than 256 points)
        const uint32_t q = q0 + (uint32_t)tid;
        const bool have = q < nQ;
        uint32_t qpos = 0;
        float4 qp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) {
            int r = 0;
            while (s_qoff[r + 1] <= q) ++r;
            qpos = s_qs[r] + (q - s_qoff[r]);
            qp = __ldg(g.pts + qpos);
        }
        float bd[KMAX];  // the k best so far, (distance, position) ascending; slots >= k stay (+inf, -1)
        int bp[KMAX];
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
            bd[i] = CUDART_INF_F;
            bp[i] = -1;
        }
        // What the staged set can prove: every point within `room` of the query is in the haloed region (faces on the grid's own
        // boundary do not count: nothing lies beyond them).  Farther candidates are never looked at twice: the gate a candidate
        // has to pass starts at room^2 and tightens to the k-th distance once k points are in.
        const float ux = (qp.x - g.ox) * g.inv_h, uy = (qp.y - g.oy) * g.inv_h, uz = (qp.z - g.oz) * g.inv_h;
        float room = CUDART_INF_F;
        if (tx0 - 1 > 0) room = fminf(room, ux - bx0);
        if (tx0 + kTX + 1 < g.nx) room = fminf(room, bx1 - ux);
        if (ty0 - 1 > 0) room = fminf(room, uy - by0);
        if (ty0 + kTY + 1 < g.ny) room = fminf(room, by1 - uy);
        if (tz0 - 1 > 0) room = fminf(room, uz - bz0);
        if (tz0 + kTZ + 1 < g.nz) room = fminf(room, bz1 - uz);
        if (room < CUDART_INF_F) room = fmaxf(room - g.slack - 1e-5f * (fabsf(ux) + fabsf(uy) + fabsf(uz)), 0.f) * g.h;
        const float gate2 = have ? room * room : -1.f;  // (inf stays inf; a lane without a query lets nothing pass)
        float kth_d = gate2;
        int kth_p = -1;  // (as unsigned: after every position, so a candidate AT the gate passes)
        issue(0);
        for (int c = 0; c < nchunk; ++c) {
            if (c + 1 < nchunk) issue(c + 1);  // the next chunk goes into the other stage, which everybody left at the end of round c - 1
            mbar_wait(&s_bar[c & 1], phase[c & 1]);
            phase[c & 1] ^= 1u;
            const uint32_t f = s_chunk_first[c], l = s_chunk_first[c + 1];
            for (uint32_t i = f; i < l; ++i) {
                const Seg sg = s_seg[i];
                // a run belongs to one (y, z) row of cells: skipped when no query of the warp can still use anything from that row
                const float ry = (float)(ty0 - 1 + (int)(sg.row % (kTY + 2))), rz = (float)(tz0 - 1 + (int)(sg.row / (kTY + 2)));
                const float gy = fmaxf(fmaxf(ry - uy, uy - (ry + 1.f)), 0.f), gz = fmaxf(fmaxf(rz - uz, uz - (rz + 1.f)), 0.f);
                const float gap = fmaxf(sqrtf(gy * gy + gz * gz) - g.slack - 1e-5f * (fabsf(uy) + fabsf(uz)), 0.f) * g.h;
                if (!__any_sync(0xffffffffu, gap * gap <= kth_d)) continue;
                const float4* src = &s_buf[c & 1][sg.dst];
                for (uint32_t j = 0; j < sg.len; ++j) {
                    const float dd = dist2_exact(qp.x, qp.y, qp.z, src[j]);  // (every thread reads the same point: a broadcast)
                    const int pos = (int)(sg.src + j);
                    if (dd < kth_d || (dd == kth_d && (unsigned)pos < (unsigned)kth_p)) {
                        // bubble the newcomer down the sorted list
                        float cd = dd;
                        int cp = pos;
#pragma unroll
                        for (int s = 0; s < KMAX; ++s) {
                            const bool lt = cd < bd[s] || (cd == bd[s] && (unsigned)cp < (unsigned)bp[s]);
                            if (lt && s < k) {
                                const float td = bd[s];
                                const int tp = bp[s];
                                bd[s] = cd;
                                bp[s] = cp;
                                cd = td;
                                cp = tp;
                            }
                            if (s == k - 1 && bd[s] <= gate2) {  // (k points in: the k-th distance is the gate from here on)
                                kth_d = bd[s];
                                kth_p = bp[s];
                            }
                        }
                    }
                }
            }
            __syncthreads();  // everybody is done with this stage before the copies of round c + 2 overwrite it
        }
        if (have) {
            // exact iff k points were found inside the gate (or the region has no inner face at all: then it holds every point)
            float kthf = CUDART_INF_F;
#pragma unroll
            for (int s = 0; s < KMAX; ++s)
                if (s == k - 1) kthf = bd[s];
            const bool exact = room == CUDART_INF_F || kthf < CUDART_INF_F;
            if (!exact) {
                const unsigned slot = atomicAdd(fb_count, 1u);
                if (slot < fb_capacity) fb_list[slot] = qpos;
            }
            // (written in both cases: the shell walk overwrites the rows it redoes)
#pragma unroll
            for (int s = 0; s < KMAX; ++s) {
                if (s < k) {
                    out_ids[(size_t)qpos * k + s] = bp[s];
                    out_d2[(size_t)qpos * k + s] = bd[s];
                }
            }
        }
    }
}

// rows of the redone queries back to their places
__global__ void __launch_bounds__(256) selfknn_scatter_kernel(const uint32_t* __restrict__ list, const unsigned* __restrict__ n_list, unsigned capacity, int k,
                                                              const int32_t* __restrict__ ids, const float* __restrict__ d2,
                                                              int32_t* __restrict__ out_ids, float* __restrict__ out_d2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n = min((long long)*n_list, (long long)capacity);
    if (i >= n * k) return;
    const long long row = i / k, c = i - row * k;
    const size_t dst = (size_t)list[row] * k + c;
    out_ids[dst] = ids[i];
    out_d2[dst] = d2[i];
}

}  // namespace

int selfknn_tiles(const GridView& g) {
    const long long tiles = (long long)((g.nx + kTX - 1) / kTX) * ((g.ny + kTY - 1) / kTY) * ((g.nz + kTZ - 1) / kTZ);
    return tiles > 0x7fffffffLL ? -1 : (int)tiles;
}

cudaError_t launch_selfknn_tiles(const GridView& g, int k, int32_t* out_ids, float* out_d2, uint32_t* fb_list, unsigned* fb_count,
                                 unsigned fb_capacity, cudaStream_t s) {
    const int tiles = selfknn_tiles(g);
    if (tiles <= 0 || k < 1 || k > 16) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(fb_count, 0, sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    if (k <= 8) selfknn_tile_kernel<8><<<tiles, kTileThreads, 0, s>>>(g, k, out_ids, out_d2, fb_list, fb_count, fb_capacity);
    else selfknn_tile_kernel<16><<<tiles, kTileThreads, 0, s>>>(g, k, out_ids, out_d2, fb_list, fb_count, fb_capacity);
    return cudaGetLastError();
}

cudaError_t launch_selfknn_scatter(const uint32_t* list, const unsigned* n_list, unsigned capacity, int k, const int32_t* ids, const float* d2,
                                   int32_t* out_ids, float* out_d2, cudaStream_t s) {
    const long long threads = (long long)capacity * k;
    if (threads <= 0) return cudaSuccess;
    selfknn_scatter_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(list, n_list, capacity, k, ids, d2, out_ids, out_d2);
    return cudaGetLastError();
}

}  // namespace b200
