// icp_device.cuh -- device-side pieces of one ICP iteration shared by icp.cu (one kernel per step)
// and loop.cu (persistent loop kernel): rigid apply, the one-warp 6x6 solve, AngleAxis / SVD
// updates, T_iter composition and the transformation checkers.
#pragma once
#include "common.cuh"

namespace b200 {
namespace {

struct Mat4 {
    float m[16];
};

__device__ __forceinline__ float3 apply_T(const float* __restrict__ T, const float4& r) {
    float3 o;
    o.x = __fadd_rn(__fmaf_rn(T[8], r.z, __fmaf_rn(T[4], r.y, __fmul_rn(T[0], r.x))), T[12]);
    o.y = __fadd_rn(__fmaf_rn(T[9], r.z, __fmaf_rn(T[5], r.y, __fmul_rn(T[1], r.x))), T[13]);
    o.z = __fadd_rn(__fmaf_rn(T[10], r.z, __fmaf_rn(T[6], r.y, __fmul_rn(T[2], r.x))), T[14]);
    return o;
}

// ---- small dense linear algebra on one thread ----------------------------------------------------
__device__ __noinline__ void jacobi_eig_f64(double* A, int n, double* V, double* w) {
    for (int i = 0; i < n * n; ++i) V[i] = 0.0;
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += A[q * n + p] * A[q * n + p];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[q * n + p];
                if (fabs(apq) < 1e-300) continue;
                const double app = A[p * n + p], aqq = A[q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[p * n + k], akq = A[q * n + k];
                    A[p * n + k] = c * akp - s * akq;
                    A[q * n + k] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[k * n + p], aqk = A[k * n + q];
                    A[k * n + p] = c * apk - s * aqk;
                    A[k * n + q] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[p * n + k], vkq = V[q * n + k];
                    V[p * n + k] = c * vkp - s * vkq;
                    V[q * n + k] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// Minimum-norm least-squares fallback (rank-deficient A), kept out of line: it is never on the
// common path and must not inflate the register allocation of the accumulate kernel.
__device__ __noinline__ void solve_min_norm(const float* A_in, const float* b, float* x, int n) {
    double Ad[36], V[36], w[6];
    for (int i = 0; i < n * n; ++i) Ad[i] = (double)A_in[i];
    jacobi_eig_f64(Ad, n, V, w);
    double wmax = 0.0;
    for (int i = 0; i < n; ++i) wmax = fmax(wmax, fabs(w[i]));
    double xd[6] = {0, 0, 0, 0, 0, 0};
    for (int e = 0; e < n; ++e) {
        if (!(fabs(w[e]) > 1e-6 * wmax)) continue;
        double proj = 0.0;
        for (int i = 0; i < n; ++i) proj += V[e * n + i] * (double)b[i];
        proj /= w[e];
        for (int i = 0; i < n; ++i) xd[i] += proj * V[e * n + i];
    }
    for (int i = 0; i < n; ++i) x[i] = (float)xd[i];
}

__device__ void quat_from_T(const float* T, float* q /* w x y z */) {
    float R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = T[c * 4 + r];
    float t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0.f) {
        t = sqrtf(t + 1.f);
        q[0] = 0.5f * t;
        t = 0.5f / t;
        q[1] = (R[2][1] - R[1][2]) * t;
        q[2] = (R[0][2] - R[2][0]) * t;
        q[3] = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(R[i][i] - R[j][j] - R[k][k] + 1.f);
        q[1 + i] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (R[k][j] - R[j][k]) * t;
        q[1 + j] = (R[j][i] + R[i][j]) * t;
        q[1 + k] = (R[k][i] + R[i][k]) * t;
    }
}

__device__ float quat_angular_distance(const float* a, const float* b) {
    const float bw = b[0], bx = -b[1], by = -b[2], bz = -b[3];
    const float w = a[0] * bw - a[1] * bx - a[2] * by - a[3] * bz;
    const float x = a[0] * bx + a[1] * bw + a[2] * bz - a[3] * by;
    const float y = a[0] * by + a[2] * bw + a[3] * bx - a[1] * bz;
    const float z = a[0] * bz + a[3] * bw + a[1] * by - a[2] * bx;
    return 2.f * atan2f(sqrtf(x * x + y * y + z * z), fabsf(w));
}

// LPM PointToPlaneErrorMinimizer::compute_in_place tail: solve, AngleAxis / Rotation2D.
// A (entries rounded to fp32 like the reference's float matrices) is factored and solved in fp32 the way
// upstream does it -- Eigen's unblocked LLT, then the two triangular solves -- by every lane of the warp
// redundantly (no communication; x ends up identical in all lanes).  "A is invertible" is restated as "every
// pivot of the factorisation exceeds 6 eps max|diag|"; the rank-deficient case falls back to the minimum-norm
// solution on lane 0.  (An fp64 Gauss-Jordan was measured at 2.2 us per iteration on this dependent chain;
// the 6x6 system does not need it: upstream itself solves in fp32.)
// Unknowns that are NOT solved for (bit i = unknown i of [alpha, beta, gamma, tx, ty, tz]): a 2-D cloud (z = 0 embedding: their
// rows are empty anyway) and force2D leave roll, pitch and tz out -- LPM solves the 3 x 3 system of [cross_z; n_x; n_y] --,
// force4DOF leaves roll and pitch out (4 x 4 system of [cross_z; n_x; n_y; n_z]).  LPM ErrorMinimizers/PointToPlane.cpp.
__device__ __forceinline__ int solve_mask(int dim, int min_flags) {
    return (dim == 2 || (min_flags & 1)) ? 0x23 : ((min_flags & 2) ? 0x03 : 0);
}

__device__ __forceinline__ void solve6_warp(const double* S, int dim, int lane, float* x /*[6]*/, float* s_x /*smem[8]*/, int min_flags = 0, IcpState* stamp_to = nullptr) {
    if (stamp_to && lane == 0) B200_STAMP(stamp_to, 5);
    float L[21], bv[6];  // lower triangle, L[i * (i + 1) / 2 + j] = (i, j), j <= i
#pragma unroll
    for (int i = 0; i < 21; ++i) L[i] = (float)S[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) bv[i] = (float)S[21 + i];
    const int mask = solve_mask(dim, min_flags);
    if (mask) {  // excluded unknowns: decoupled rows / columns with a unit diagonal and a zero right-hand side -> x = 0 there
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j)
                if (((mask >> i) | (mask >> j)) & 1) L[i * (i + 1) / 2 + j] = (i == j) ? 1.f : 0.f;
            if ((mask >> i) & 1) bv[i] = 0.f;
        }
    }
    float maxdiag = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) maxdiag = fmaxf(maxdiag, fabsf(L[i * (i + 1) / 2 + i]));
    const float thr = 6.0f * 1.1920929e-7f * maxdiag;
    if (stamp_to && lane == 0) B200_STAMP(stamp_to, 7);
    bool ok = true;
    float invd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float d = L[k * (k + 1) / 2 + k];
#pragma unroll
        for (int j = 0; j < k; ++j) d -= L[k * (k + 1) / 2 + j] * L[k * (k + 1) / 2 + j];
        if (!(d > thr)) ok = false;
        // only 1 / L_kk is ever used below: one correctly rounded reciprocal square root instead of an IEEE square root followed
        // by an IEEE division (this chain of six is on the critical path of every ICP iteration)
        invd[k] = __frsqrt_rn(d);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            float v = L[i * (i + 1) / 2 + k];
#pragma unroll
            for (int j = 0; j < k; ++j) v -= L[i * (i + 1) / 2 + j] * L[k * (k + 1) / 2 + j];
            L[i * (i + 1) / 2 + k] = v * invd[k];
        }
    }
    if (stamp_to && lane == 0) B200_STAMP(stamp_to, 8);
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // L y = b
        float v = bv[i];
#pragma unroll
        for (int j = 0; j < i; ++j) v -= L[i * (i + 1) / 2 + j] * y[j];
        y[i] = v * invd[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {  // L^T x = y
        float v = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) v -= L[j * (j + 1) / 2 + i] * x[j];
        x[i] = v * invd[i];
    }
    if (stamp_to && lane == 0) B200_STAMP(stamp_to, 9);
    bool bad = !ok;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (isnan(x[i])) bad = true;
    if (bad) {  // rare: rank-deficient normal equations
        if (lane == 0) {
            float A6[36], b6[6], xs[6] = {0, 0, 0, 0, 0, 0};
            for (int c = 0; c < 6; ++c)
                for (int rr = 0; rr <= c; ++rr) {
                    const float v = (float)S[c * (c + 1) / 2 + rr];
                    A6[c * 6 + rr] = v;
                    A6[rr * 6 + c] = v;
                }
            for (int i = 0; i < 6; ++i) b6[i] = (float)S[21 + i];
            if (!mask) {
                solve_min_norm(A6, b6, xs, 6);
            } else {  // the sub-system of the unknowns that are solved for
                float As[36], bs[6], xr[6];
                int id[6], n = 0;
                for (int i = 0; i < 6; ++i)
                    if (!((mask >> i) & 1)) id[n++] = i;
                for (int c = 0; c < n; ++c) {
                    for (int rr = 0; rr < n; ++rr) As[c * n + rr] = A6[id[c] * 6 + id[rr]];
                    bs[c] = b6[id[c]];
                }
                solve_min_norm(As, bs, xr, n);
                for (int c = 0; c < n; ++c) xs[id[c]] = xr[c];
            }
            for (int i = 0; i < 6; ++i) s_x[i] = xs[i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = s_x[i];
    }
}

// x -> dT (R column-major in dT[0..8], t in dT[9..11]); Eigen AngleAxis(|r|, r/|r|) / Rotation2D.
__device__ __forceinline__ void delta_from_x(const float* x, int dim, float* dT, int min_flags = 0) {
    dT[0] = 1.f; dT[1] = 0.f; dT[2] = 0.f;
    dT[3] = 0.f; dT[4] = 1.f; dT[5] = 0.f;
    dT[6] = 0.f; dT[7] = 0.f; dT[8] = 1.f;
    if (dim == 3 && !(min_flags & 3)) {
        const float nrm2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
        const float ang = sqrtf(nrm2);
        float ax0 = x[0], ax1 = x[1], ax2 = x[2];
        if (nrm2 > 0.f) {
            ax0 = x[0] / ang;
            ax1 = x[1] / ang;
            ax2 = x[2] / ang;
        }
        float s, c;
        sincosf(ang, &s, &c);
        const float sx = s * ax0, sy = s * ax1, sz = s * ax2;
        const float cx = (1.f - c) * ax0, cy = (1.f - c) * ax1, cz = (1.f - c) * ax2;
        float R[9];  // column-major
        float tmp = cx * ax1;
        R[3] = tmp - sz;  // (0,1)
        R[1] = tmp + sz;  // (1,0)
        tmp = cx * ax2;
        R[6] = tmp + sy;  // (0,2)
        R[2] = tmp - sy;  // (2,0)
        tmp = cy * ax2;
        R[7] = tmp - sx;  // (1,2)
        R[5] = tmp + sx;  // (2,1)
        R[0] = cx * ax0 + c;
        R[4] = cy * ax1 + c;
        R[8] = cz * ax2 + c;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if (isnan(R[i])) bad = true;
        if (!bad) {
#pragma unroll
            for (int i = 0; i < 9; ++i) dT[i] = R[i];
        }
        dT[9] = x[3];
        dT[10] = x[4];
        dT[11] = x[5];
    } else {  // rotation about z only: theta = x[2], t = (x[3], x[4], x[5]) -- 2-D clouds and force2D (x[5] = 0: Rotation2D
              // written into the top-left corner of an identity), force4DOF (AngleAxis(x[2], unitZ) + the full translation)
        const float s = sinf(x[2]), c = cosf(x[2]);
        dT[0] = c;
        dT[1] = s;
        dT[3] = -s;
        dT[4] = c;
        dT[9] = x[3];
        dT[10] = x[4];
        dT[11] = x[5];
    }
}

// Symmetric 3x3 eigen-decomposition for the point-to-point SVD: cyclic Jacobi in fp64 like jacobi_eig_f64, but the
// sweep stops at a RELATIVE off-diagonal level (1e-30 of the squared diagonal: far below fp64 resolution) instead of
// running on until the off-diagonals underflow, and a rotation costs one sqrt, one division and one rsqrt.  This
// serial chain runs once per ICP iteration on one thread: it was ~20 us of every point-to-point iteration.
__device__ __noinline__ void jacobi_eig3_f64(double* A, double* V, double* w) {
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        const double dia = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
        if (!(off > 1e-30 * dia) || off < 1e-300) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[q * 3 + p];
                if (fabs(apq) < 1e-300) continue;
                const double app = A[p * 3 + p], aqq = A[q * 3 + q];
                const double d = aqq - app;
                // t = sign(theta) / (|theta| + sqrt(theta^2 + 1)), theta = d / (2 apq), without forming theta
                const double t = (d >= 0 ? 2.0 : -2.0) * apq / (fabs(d) + sqrt(d * d + 4.0 * apq * apq));
                const double c = rsqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[p * 3 + k], akq = A[q * 3 + k];
                    A[p * 3 + k] = c * akp - sn * akq;
                    A[q * 3 + k] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[k * 3 + p], aqk = A[k * 3 + q];
                    A[k * 3 + p] = c * apk - sn * aqk;
                    A[k * 3 + q] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[p * 3 + k], vkq = V[q * 3 + k];
                    V[p * 3 + k] = c * vkp - sn * vkq;
                    V[q * 3 + k] = sn * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i * 3 + i];
}

// LPM PointToPointErrorMinimizer::compute_in_place: weighted centroids, SVD of the cross-covariance.
__device__ __noinline__ void delta_point_to_point(const double* S, int dim, float* dT) {
    for (int i = 0; i < 16; ++i) dT[i] = (i % 5 == 0) ? 1.f : 0.f;
    const double sw = S[0];
    double mp[3], mq[3];
    for (int d = 0; d < 3; ++d) {
        mp[d] = S[1 + d] / sw;
        mq[d] = S[4 + d] / sw;
    }
    double M[9];
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) M[c * dim + r] = S[7 + c * 3 + r] - sw * mq[r] * mp[c];
    if (dim == 2) {
        // R = U diag(1, det(U V^T)) V^T of a 2x2 M in closed form: the rotation that maximises tr(R^T M),
        // (cos, sin) parallel to (M00 + M11, M10 - M01).  Degenerate M (both zero): identity, like a zero SVD.
        const double a = M[0] + M[3], b = M[1] - M[2];  // M[c * 2 + r]
        const double hyp = sqrt(a * a + b * b);
        double cs = 1.0, sn = 0.0;
        if (hyp > 0.0) {
            cs = a / hyp;
            sn = b / hyp;
        }
        dT[0] = (float)cs;
        dT[1] = (float)sn;
        dT[4] = (float)-sn;
        dT[5] = (float)cs;
        for (int r = 0; r < 2; ++r) {
            float acc = 0.f;
            for (int c = 0; c < 2; ++c) acc += dT[c * 4 + r] * (float)mp[c];
            dT[12 + r] = (float)mq[r] - acc;
        }
        return;
    }
    double MtM[9], V[9], ev[3];
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) {
            double a = 0;
            for (int j = 0; j < dim; ++j) a += M[r * dim + j] * M[c * dim + j];
            MtM[c * dim + r] = a;
        }
    jacobi_eig3_f64(MtM, V, ev);
    int order[3] = {0, 1, 2};
    for (int a = 0; a < dim; ++a)
        for (int bb = a + 1; bb < dim; ++bb)
            if (ev[order[bb]] > ev[order[a]]) {
                const int tt = order[a];
                order[a] = order[bb];
                order[bb] = tt;
            }
    double U[9], Vs[9];
    for (int e = 0; e < dim; ++e)
        for (int r = 0; r < dim; ++r) Vs[e * dim + r] = V[order[e] * dim + r];
    for (int e = 0; e < dim; ++e) {
        double u[3] = {0, 0, 0}, nn = 0;
        for (int r = 0; r < dim; ++r) {
            for (int c = 0; c < dim; ++c) u[r] += M[c * dim + r] * Vs[e * dim + c];
            nn += u[r] * u[r];
        }
        nn = sqrt(nn);
        for (int r = 0; r < dim; ++r) U[e * dim + r] = (nn > 0) ? u[r] / nn : 0.0;
    }
    if (dim == 3) {
        double* u0 = U;
        double* u1 = U + 3;
        double* u2 = U + 6;
        const double cx[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
        const double sgn = (cx[0] * u2[0] + cx[1] * u2[1] + cx[2] * u2[2]) < 0 ? -1.0 : 1.0;
        for (int r = 0; r < 3; ++r) u2[r] = sgn * cx[r];
    } else {
        double* u0 = U;
        double* u1 = U + 2;
        const double px[2] = {-u0[1], u0[0]};
        const double sgn = (px[0] * u1[0] + px[1] * u1[1]) < 0 ? -1.0 : 1.0;
        u1[0] = sgn * px[0];
        u1[1] = sgn * px[1];
    }
    double R[9];
    for (int pass = 0; pass < 2; ++pass) {
        for (int r = 0; r < dim; ++r)
            for (int c = 0; c < dim; ++c) {
                double a = 0;
                for (int e = 0; e < dim; ++e) a += U[e * dim + r] * Vs[e * dim + c];
                R[c * dim + r] = a;
            }
        const double det = (dim == 2) ? R[0] * R[3] - R[2] * R[1]
                                      : R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
                                            R[6] * (R[1] * R[5] - R[4] * R[2]);
        if (det >= 0) break;
        for (int c = 0; c < dim; ++c) Vs[(dim - 1) * dim + c] = -Vs[(dim - 1) * dim + c];
    }
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) dT[c * 4 + r] = (float)R[c * dim + r];
    for (int r = 0; r < dim; ++r) {
        float a = 0.f;
        for (int c = 0; c < dim; ++c) a += dT[c * 4 + r] * (float)mp[c];
        dT[12 + r] = (float)mq[r] - a;
    }
}

// Tail of one iteration, run by lane 0: T_iter = dT * T_iter, bookkeeping, then the checkers
// (LPM TransformationCheckersImpl.cpp: Counter, Differential, Bound).
__device__ __forceinline__ void finish_iteration(const IcpParams& prm, IcpState* st, const float* dT12, double pairs, double wsum,
                                              float* trace) {
    const double denom = (double)prm.knn * (double)st->nq;
    float T[16];
    {
        float old[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) old[i] = st->T[i];
        // T = [R t; 0 1] * old   (same accumulation order as the oracle's 4x4 product)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) acc += dT12[k * 3 + r] * old[c * 4 + k];
                acc += dT12[9 + r] * old[c * 4 + 3];
                T[c * 4 + r] = acc;
            }
            T[c * 4 + 3] = old[c * 4 + 3];
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) st->T[i] = T[i];
    st->pairs = (long long)pairs;
    st->used_ratio = (float)pairs / (float)denom;
    st->overlap = (float)wsum / (float)denom;
    const int it = st->iter;
    if (trace) {
#pragma unroll
        for (int i = 0; i < 16; ++i) trace[(size_t)it * 16 + i] = T[i];
    }
    st->iter = it + 1;

    // LPM TransformationCheckers::check runs the checkers in their YAML order; the Counter signals its limit by THROWING
    // MaxNumIterationsReached (caught by ICP.cpp), which skips the checkers listed after it for that iteration.  counter_after:
    // bit 0 = Differential is listed before the Counter, bit 1 = Bound is (0 = LPM's setDefault order: Counter first).
    bool iterate = true;
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1 && prm.max_iteration_count > 0) {
            st->counter += 1;
            if (st->counter >= prm.max_iteration_count) {
                st->max_iter_reached = 1;
                st->done = 1;
                return;
            }
        }
        if (prm.use_differential && (((prm.counter_after & 1) != 0) == (phase == 0))) {
            const int smooth = min(max(prm.smooth_length, 1), 7);
            const int slot = st->dcount % 8;
            quat_from_T(T, st->dq[slot]);
            for (int d = 0; d < 3; ++d) st->dt[slot][d] = T[12 + d];
            st->dcount += 1;
            float vr = 0.f, vt = 0.f;
            if (st->dcount > smooth) {
                for (int j = st->dcount - 1; j >= st->dcount - smooth; --j) {
                    const int a = j % 8, b = (j - 1) % 8;
                    vr += fabsf(quat_angular_distance(st->dq[a], st->dq[b]));
                    const float dx = st->dt[a][0] - st->dt[b][0], dy = st->dt[a][1] - st->dt[b][1], dz = st->dt[a][2] - st->dt[b][2];
                    vt += fabsf(sqrtf(dx * dx + dy * dy + dz * dz));
                }
                vr /= (float)smooth;
                vt /= (float)smooth;
                if (vr < prm.min_diff_rot_err && vt < prm.min_diff_trans_err) iterate = false;
            }
            if (isnan(vr) || isnan(vt)) {
                st->status = B200ICP_ERR_NAN;
                st->done = 1;
                return;
            }
        }
        if (prm.use_bound && (((prm.counter_after & 2) != 0) == (phase == 0))) {
            float qc[4];
            quat_from_T(T, qc);
            const float vr = quat_angular_distance(qc, st->bq0);
            float vt = 0.f;
            for (int d = 0; d < 3; ++d) vt += (T[12 + d] - st->bt0[d]) * (T[12 + d] - st->bt0[d]);
            vt = sqrtf(vt);
            if (isnan(vr) || isnan(vt)) {
                st->status = B200ICP_ERR_NAN;
                st->done = 1;
                return;
            }
            if (vr > prm.max_rotation_norm || vt > prm.max_translation_norm) {
                st->status = B200ICP_ERR_BOUND;
                st->done = 1;
                return;
            }
        }
    }
    if (prm.max_iteration_count <= 0 && !prm.use_differential) iterate = false;
    if (!iterate) {
        st->done = 1;
        return;
    }
    // the next transformations.apply(stepReading, T_iter) checks orthonormality
    const float det = T[0] * (T[5] * T[10] - T[9] * T[6]) - T[4] * (T[1] * T[10] - T[9] * T[2]) + T[8] * (T[1] * T[6] - T[5] * T[2]);
    if (fabsf(1.f - det) > 1e-3f) {
        st->status = B200ICP_ERR_TRANSFORM;
        st->done = 1;
    }
}

// Everything after the sums, on warp 0 of the last block.
__device__ __forceinline__ void finish_warp(const IcpParams& prm, IcpState* st, const double* S, int n_sums, float* trace,
                                            float* s_scratch /* smem[16] */, IcpState* stamp_to = nullptr) {
    const int lane = threadIdx.x & 31;
    const double pairs = S[n_sums - 1];
    const double wsum = S[n_sums - 2];
    if (pairs == 0.0) {  // LPM: ConvergenceError("ErrorMnimizer: no point to minimize")
        if (lane == 0) {
            st->status = B200ICP_ERR_CONVERGENCE;
            st->done = 1;
        }
        return;
    }
    float dT[12];
    if (prm.minimizer == B200ICP_MIN_POINT_TO_PLANE) {
        float x[6];
        solve6_warp(S, prm.dim, lane, x, s_scratch, prm.min_flags, stamp_to);
        if (stamp_to && lane == 0) B200_STAMP(stamp_to, 15);
        delta_from_x(x, prm.dim, dT, prm.min_flags);
        if (stamp_to && lane == 0) B200_STAMP(stamp_to, 16);
    } else {
        if (lane == 0) {
            float d16[16];
            for (int i = 0; i < 16; ++i) d16[i] = (i % 5 == 0) ? 1.f : 0.f;
            if (prm.minimizer == B200ICP_MIN_POINT_TO_POINT) delta_point_to_point(S, prm.dim, d16);
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) s_scratch[c * 3 + r] = d16[c * 4 + r];
            for (int r = 0; r < 3; ++r) s_scratch[9 + r] = d16[12 + r];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 12; ++i) dT[i] = s_scratch[i];
    }
    if (lane == 0) finish_iteration(prm, st, dT, pairs, wsum, trace);
}

// ---- ErrorElements + error minimiser sums --------------------------------------------------------
// MIN: 0 point-to-plane (29 sums), 1 point-to-point (18 sums), 2 identity (2 sums).
template <int MIN>
struct SumLayout;
template <>
struct SumLayout<0> {
    static constexpr int N = 29;
};
template <>
struct SumLayout<1> {
    static constexpr int N = 18;
};
template <>
struct SumLayout<2> {
    static constexpr int N = 2;
};

// Sum 32 per-lane values v[0..31] across the 32 lanes of a warp with 31 shuffles (butterfly
// transpose): on return every lane holds the warp total of ONE slot, namely slot `lane_slot(lane)`.
// fp32: pairwise tree over the lanes, deterministic; int: exact.
template <typename V>
__device__ __forceinline__ V warp_reduce_32slots(V (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const V send = upper ? v[j] : v[j + half];
            const V keep = upper ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}
// slot held by `lane` after warp_reduce_32slots: bit b of the lane selects the upper half at level b
__device__ __forceinline__ int lane_slot(int lane) { return lane; }

// LPM SurfaceNormalOutlierFilter: keep when the reading normal (rotated by T_iter like the reading) and the reference normal,
// both normalised, have a dot product >= eps = cos(maxAngle).
__device__ __forceinline__ bool surface_normal_keep(const float* T, const float4& rn, const float4& fn, float eps) {
    const float rx = __fmaf_rn(T[8], rn.z, __fmaf_rn(T[4], rn.y, __fmul_rn(T[0], rn.x)));
    const float ry = __fmaf_rn(T[9], rn.z, __fmaf_rn(T[5], rn.y, __fmul_rn(T[1], rn.x)));
    const float rz = __fmaf_rn(T[10], rn.z, __fmaf_rn(T[6], rn.y, __fmul_rn(T[2], rn.x)));
    const float lr = sqrtf(rx * rx + ry * ry + rz * rz), lf = sqrtf(fn.x * fn.x + fn.y * fn.y + fn.z * fn.z);
    const float dot = (rx / lr) * (fn.x / lf) + (ry / lr) * (fn.y / lf) + (rz / lr) * (fn.z / lf);
    return !(dot < eps);
}

// One (reading point, neighbour) entry of ErrorElements: outlier weights, then the error-minimiser
// products.  MIN: 0 point-to-plane (29 sums), 1 point-to-point (18 sums), 2 identity (2 sums).
// LPM RobustOutlierFilter::robustFiltering: the M-estimator weight of one match.  dist = squared distance (point2point) or
// squared distance along the map point's unit normal (point2plane); e2 = dist / scale^2; k = tuning (berg: the constant of
// Bergstrom & Edlund 2014 for cauchy / tukey / huber, the tuning parameter being the target scale there).
__device__ __forceinline__ float robust_weight(int mode, float tuning, float approximation, float scale, float dist) {
    const int fct = mode & 255, est = (mode >> 8) & 15;
    float k = tuning;
    if (est == B200ICP_SCALE_BERG) k = fct == B200ICP_ROBUST_CAUCHY ? 4.3040f : (fct == B200ICP_ROBUST_TUKEY ? 7.0589f : (fct == B200ICP_ROBUST_HUBER ? 2.0138f : tuning));
    const float s = est == B200ICP_SCALE_NONE ? 1.f : scale;
    const float e2 = dist / (s * s);
    const float k2 = k * k;
    float w = 1.f;
    switch (fct) {
        case B200ICP_ROBUST_CAUCHY: w = 1.f / (1.f + e2 / k2); break;
        case B200ICP_ROBUST_WELSCH: w = expf(-e2 / k2); break;
        case B200ICP_ROBUST_SC: w = e2 >= k ? 4.f * k2 / ((k + e2) * (k + e2)) : 1.f; break;
        case B200ICP_ROBUST_GM: w = k2 / ((k + e2) * (k + e2)); break;
        case B200ICP_ROBUST_TUKEY: w = e2 >= k2 ? 0.f : (1.f - e2 / k2) * (1.f - e2 / k2); break;
        case B200ICP_ROBUST_HUBER: w = e2 >= k2 ? k / sqrtf(e2) : 1.f; break;
        case B200ICP_ROBUST_L1: w = 1.f / sqrtf(e2); break;
        case B200ICP_ROBUST_STUDENT: w = powf(1.f + e2 / k, -(k + 3.f) * 0.5f) * (k + 3.f) / (k + e2); break;
        default: break;
    }
    if (w <= 0.f) w = 0.f;  // (upstream's 1e-50 floor is 0 in fp32)
    const float a2 = approximation * approximation;
    if (a2 != CUDART_INF_F && e2 >= a2) w = 0.f;
    return w;
}

template <int MIN>
__device__ __forceinline__ void accumulate_entry(float* acc, const IcpParams& prm, const float* T, const GridView& g,
                                                 const float4* __restrict__ nrm, const float4* __restrict__ reading, long long e,
                                                 int K, int pos, float d, float qlimit, float robust_scale = 1.f) {
    constexpr int NS = SumLayout<MIN>::N;
    if (pos < 0) return;
    if (d == CUDART_INF_F) return;
    float w = 1.f;
    for (int f = 0; f < prm.n_outlier; ++f) {
        const float p = prm.outlier_param[f];
        bool keep = true;
        switch (prm.outlier_kind[f]) {
            case B200ICP_OUTLIER_TRIMMED_DIST:
            case B200ICP_OUTLIER_VAR_TRIMMED_DIST: keep = d <= qlimit; break;
            case B200ICP_OUTLIER_MEDIAN_DIST: keep = d <= p * qlimit; break;
            case B200ICP_OUTLIER_MAX_DIST: keep = d <= p * p; break;
            case B200ICP_OUTLIER_MIN_DIST: keep = d >= p * p; break;
            case B200ICP_OUTLIER_SURFACE_NORMAL:
                if (prm.rnrm && nrm) keep = surface_normal_keep(T, __ldg(prm.rnrm + ((K == 1) ? e : e / K)), __ldg(nrm + pos), cosf(p));
                break;
        }
        w *= keep ? 1.f : 0.f;
        if (prm.outlier_kind[f] == B200ICP_OUTLIER_ROBUST) {
            float dist = d;
            if ((prm.outlier_mode[f] >> 12) & 1) {  // point2plane: (n . (p - q))^2, n normalised (nrm is there: checked by the host)
                const float3 pr = apply_T(T, __ldg(reading + ((K == 1) ? e : e / K)));
                const float4 qm = __ldg(g.pts + pos), nm = __ldg(nrm + pos);
                const float inv = 1.f / sqrtf(nm.x * nm.x + nm.y * nm.y + nm.z * nm.z);
                const float dot = (nm.x * inv) * (pr.x - qm.x) + (nm.y * inv) * (pr.y - qm.y) + (nm.z * inv) * (pr.z - qm.z);
                dist = dot * dot;
            }
            w *= robust_weight(prm.outlier_mode[f], p, prm.outlier_param2[f], robust_scale, dist);
        }
    }
    if (w == 0.f) return;
    const long long i = (K == 1) ? e : e / K;
    const float3 p = apply_T(T, __ldg(reading + i));
    const float4 q = __ldg(g.pts + pos);
    if (MIN == 0) {
        const float4 n = __ldg(nrm + pos);
        float F[6];
        F[0] = p.y * n.z - p.z * n.y;
        F[1] = p.z * n.x - p.x * n.z;
        F[2] = p.x * n.y - p.y * n.x;
        F[3] = n.x;
        F[4] = n.y;
        F[5] = n.z;
        // force2D: the clouds are cut down to x, y before the residual is formed (LPM PointToPlane.cpp), z does not count
        const float dot = (prm.min_flags & 1) ? (p.x - q.x) * n.x + (p.y - q.y) * n.y : (p.x - q.x) * n.x + (p.y - q.y) * n.y + (p.z - q.z) * n.z;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float wf = w * F[c];
#pragma unroll
            for (int r = 0; r <= c; ++r) acc[c * (c + 1) / 2 + r] += wf * F[r];
            acc[21 + c] -= wf * dot;
        }
    } else if (MIN == 1) {
        const float pv[3] = {p.x, p.y, p.z}, qv[3] = {q.x, q.y, q.z};
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

}  // namespace
}  // namespace b200
