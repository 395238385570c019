/*
 * icp_oracle.c -- CPU restatement of the ICP hot path that norlab_icp_mapper reaches through
 * libpointmatcher / libnabo.  TEST INFRASTRUCTURE ONLY (see icp_oracle.h): the checker for the
 * CUDA path and the timed CPU baseline.  PARITY UNPINNED (no upstream source, tests or golden
 * vectors available here).
 *
 * Citations: "ref:" lines name the call site under /root/reference that reaches the restated
 * upstream routine; "LPM"/"NABO" name the upstream file of libpointmatcher 1.4.x / libnabo whose
 * published algorithm is restated.
 *
 * Floating-point conventions shared with the CUDA path so that neighbour distances can be
 * compared bit for bit (upstream leaves both to the compiler's contraction choices):
 *   - rigid apply:  x' = fma(T02,z, fma(T01,y, T00*x)) + T03   (k-order of a column-major GEMM)
 *   - distance:     d2 = fma(dz,dz, fma(dy,dy, dx*dx))
 *   - map mean:     exact fixed-point sum (2^-16 m) / N, rounded to fp32 (order independent)
 */
#define _GNU_SOURCE
#include "icp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_BUCKET 8 /* NABO: KDTreeMatcher creates the tree with the default bucketSize 8 */
#define ORC_MAXK 64

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int32_t orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ============================================================================================
 * libnabo kd-tree (NABO nabo/kdtree_cpu.cpp, KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt):
 * points in leaves, split on the dimension of largest extent, at the median by count
 * (nth_element), bucket size 8; search = depth-first with incremental "rd" lower bound
 * (Arya & Mount), pruning `rd <= maxRadius2 && rd*(1+eps)^2 < worst`, eps = 0 here.
 * ============================================================================================ */
typedef struct {
    int32_t dim;   /* split dimension, -1 for a leaf */
    float cut;     /* split value */
    int32_t right; /* index of right child (left child = this + 1) */
    int32_t start; /* leaf: first bucket entry */
    int32_t count; /* leaf: entries */
} orc_node;

struct orc_kdtree {
    int32_t dim;
    int64_t n;
    float* pts;    /* bucket-ordered coordinates, 4 floats per point (unused = 0) */
    int32_t* idx;  /* bucket-ordered original indices */
    orc_node* nodes;
    int64_t n_nodes, cap_nodes;
};

typedef struct {
    const float* src;
    int32_t stride;
    int32_t* perm;
} build_ctx;

static inline float key_of(const build_ctx* c, int32_t id, int d) {
    return c->src[(int64_t)id * c->stride + d];
}

/* nth_element on perm[lo,hi) by coordinate d: after return, perm[nth] holds the element that a
 * full sort would put there, everything before is <=, everything after is >=. */
static void nth_element_dim(const build_ctx* c, int64_t lo, int64_t hi, int64_t nth, int d) {
    int32_t* p = c->perm;
    while (hi - lo > 1) {
        /* median of three pivot */
        int64_t mid = lo + (hi - lo) / 2;
        float a = key_of(c, p[lo], d), b = key_of(c, p[mid], d), e = key_of(c, p[hi - 1], d);
        float pivot = (a < b) ? ((b < e) ? b : (a < e ? e : a)) : ((a < e) ? a : (b < e ? e : b));
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (key_of(c, p[i], d) < pivot) ++i;
            while (key_of(c, p[j], d) > pivot) --j;
            if (i <= j) {
                int32_t t = p[i];
                p[i] = p[j];
                p[j] = t;
                ++i;
                --j;
            }
        }
        /* [lo, j] <= pivot, [i, hi) >= pivot, (j, i) == pivot */
        if (nth <= j)
            hi = j + 1;
        else if (nth >= i)
            lo = i;
        else
            return;
    }
}

static int32_t new_node(orc_kdtree* t) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
        t->nodes = (orc_node*)realloc(t->nodes, (size_t)t->cap_nodes * sizeof(orc_node));
    }
    return (int32_t)t->n_nodes++;
}

static int32_t build_rec(orc_kdtree* t, const build_ctx* c, int64_t lo, int64_t hi) {
    const int32_t me = new_node(t);
    const int64_t count = hi - lo;
    if (count <= ORC_BUCKET) {
        t->nodes[me].dim = -1;
        t->nodes[me].cut = 0.f;
        t->nodes[me].right = -1;
        t->nodes[me].start = (int32_t)lo;
        t->nodes[me].count = (int32_t)count;
        return me;
    }
    /* bounds of this node's points -> dimension of largest extent */
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = lo; i < hi; ++i)
        for (int d = 0; d < t->dim; ++d) {
            float v = key_of(c, c->perm[i], d);
            if (v < mn[d]) mn[d] = v;
            if (v > mx[d]) mx[d] = v;
        }
    int cd = 0;
    for (int d = 1; d < t->dim; ++d)
        if (mx[d] - mn[d] > mx[cd] - mn[cd]) cd = d;
    const int64_t right_count = count / 2, left_count = count - right_count;
    nth_element_dim(c, lo, hi, lo + left_count, cd);
    const float cut = key_of(c, c->perm[lo + left_count], cd);
    (void)build_rec(t, c, lo, lo + left_count); /* left child is me + 1 */
    const int32_t r = build_rec(t, c, lo + left_count, hi);
    t->nodes[me].dim = cd;
    t->nodes[me].cut = cut;
    t->nodes[me].right = r;
    t->nodes[me].start = 0;
    t->nodes[me].count = 0;
    return me;
}

orc_kdtree* orc_kdtree_build(const float* pts, int32_t stride, int64_t n, int32_t dim) {
    if (!pts || n <= 0 || dim < 1 || dim > 3 || n > INT32_MAX) return NULL;
    orc_kdtree* t = (orc_kdtree*)calloc(1, sizeof(orc_kdtree));
    t->dim = dim;
    t->n = n;
    build_ctx c;
    c.src = pts;
    c.stride = stride;
    c.perm = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) c.perm[i] = (int32_t)i;
    build_rec(t, &c, 0, n);
    t->pts = (float*)calloc((size_t)n * 4, sizeof(float));
    t->idx = c.perm;
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) t->pts[i * 4 + d] = pts[(int64_t)c.perm[i] * stride + d];
    return t;
}

void orc_kdtree_free(orc_kdtree* t) {
    if (!t) return;
    free(t->pts);
    free(t->idx);
    free(t->nodes);
    free(t);
}

static inline float dist2_fma(const float* a, const float* b) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

typedef struct {
    const orc_kdtree* t;
    float q[4];
    float off[3];
    float max_r2;
    int k;
    float hd[ORC_MAXK]; /* NABO IndexHeapBruteForceVector: kept sorted ascending by insertion */
    int32_t hi[ORC_MAXK];
} search_ctx;

static void search_rec(search_ctx* s, int32_t ni, float rd) {
    const orc_node* nd = &s->t->nodes[ni];
    if (nd->dim < 0) {
        const float* p = s->t->pts + (int64_t)nd->start * 4;
        for (int j = 0; j < nd->count; ++j, p += 4) {
            const float d = dist2_fma(s->q, p);
            /* NABO: (dist <= maxRadius2) && (dist < heap.headValue()) ; ALLOW_SELF_MATCH set */
            if (d <= s->max_r2 && d < s->hd[s->k - 1]) {
                int i = s->k - 1;
                for (; i > 0; --i) {
                    if (s->hd[i - 1] > d) {
                        s->hd[i] = s->hd[i - 1];
                        s->hi[i] = s->hi[i - 1];
                    } else
                        break;
                }
                s->hd[i] = d;
                s->hi[i] = s->t->idx[nd->start + j];
            }
        }
        return;
    }
    const int cd = nd->dim;
    const float old_off = s->off[cd];
    const float new_off = s->q[cd] - nd->cut;
    const int32_t near_child = (new_off > 0.f) ? nd->right : ni + 1;
    const int32_t far_child = (new_off > 0.f) ? ni + 1 : nd->right;
    search_rec(s, near_child, rd);
    rd += -old_off * old_off + new_off * new_off;
    if (rd <= s->max_r2 && rd < s->hd[s->k - 1]) {
        s->off[cd] = new_off;
        search_rec(s, far_child, rd);
        s->off[cd] = old_off;
    }
}

void orc_kdtree_knn(const orc_kdtree* t, const float* q, int32_t qstride, int64_t nq, int32_t k,
                    float max_radius, int32_t* ids, float* d2, int32_t nthreads) {
    orc_kdtree_knn_ex(t, q, qstride, nq, k, max_radius, NULL, 0, ids, d2, nthreads);
}

/* max_radii (optional, one per query): LPM KDTreeMatcher with a `maxSearchDist` descriptor on the reading -- libnabo's knn
 * overload with a vector of radii; replaces max_radius.  strict: accept dist2 < r^2 instead of <= (b200icp_config::conventions bit 0). */
void orc_kdtree_knn_ex(const orc_kdtree* t, const float* q, int32_t qstride, int64_t nq, int32_t k, float max_radius,
                       const float* max_radii, int32_t strict, int32_t* ids, float* d2, int32_t nthreads) {
    if (!t || k < 1 || k > ORC_MAXK) return;
    const float max_r2_all = isinf(max_radius) ? INFINITY : max_radius * max_radius;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        search_ctx s;
        s.t = t;
        s.k = k;
        float max_r2 = max_r2_all;
        if (max_radii) max_r2 = isinf(max_radii[i]) ? INFINITY : max_radii[i] * max_radii[i];
        if (strict && !isinf(max_r2)) max_r2 = nextafterf(max_r2, 0.f); /* d < r2  <=>  d <= the next float below r2 */
        s.max_r2 = max_r2;
        s.q[0] = s.q[1] = s.q[2] = s.q[3] = 0.f;
        for (int d = 0; d < t->dim; ++d) s.q[d] = q[i * qstride + d];
        s.off[0] = s.off[1] = s.off[2] = 0.f;
        for (int j = 0; j < k; ++j) {
            s.hd[j] = INFINITY;
            s.hi[j] = -1;
        }
        search_rec(&s, 0, 0.f);
        for (int j = 0; j < k; ++j) {
            ids[i * k + j] = (s.hd[j] == INFINITY) ? -1 : s.hi[j];
            d2[i * k + j] = s.hd[j];
        }
    }
}

/* ============================================================================================
 * small dense linear algebra (fp32 like upstream's Eigen float path; fp64 only in fallbacks)
 * ============================================================================================ */
static void mat4_identity(float* M, int n) {
    for (int i = 0; i < n * n; ++i) M[i] = 0.f;
    for (int i = 0; i < n; ++i) M[i * n + i] = 1.f;
}
/* C = A*B, column-major n x n */
static void matn_mul(const float* A, const float* B, float* C, int n) {
    float tmp[16];
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) {
            float acc = 0.f;
            for (int k = 0; k < n; ++k) acc += A[k * n + r] * B[c * n + k];
            tmp[c * n + r] = acc;
        }
    memcpy(C, tmp, sizeof(float) * (size_t)(n * n));
}

static float det_rot(const float* T, int dim) {
    const int n = dim + 1;
#define M_(r, c) T[(c) * n + (r)]
    if (dim == 2) return M_(0, 0) * M_(1, 1) - M_(0, 1) * M_(1, 0);
    return M_(0, 0) * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1)) -
           M_(0, 1) * (M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0)) +
           M_(0, 2) * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
#undef M_
}

/* LPM TransformationsImpl.cpp RigidTransformation::compute: features <- T*features, `normals`
 * descriptor <- R*normals; TransformationError if |1 - det R| > 1e-3. Points are `rows` floats
 * apart; T is (dim+1)x(dim+1) column-major. */
static void apply_T_point(const float* T, int dim, const float* p, float* out) {
    const int n = dim + 1;
    if (dim == 3) {
        for (int r = 0; r < 3; ++r)
            out[r] = fmaf(T[2 * n + r], p[2], fmaf(T[1 * n + r], p[1], T[0 * n + r] * p[0])) +
                     T[3 * n + r];
    } else {
        for (int r = 0; r < 2; ++r) out[r] = fmaf(T[1 * n + r], p[1], T[0 * n + r] * p[0]) + T[2 * n + r];
    }
}
static void apply_R_vec(const float* T, int dim, const float* v, float* out) {
    const int n = dim + 1;
    if (dim == 3) {
        for (int r = 0; r < 3; ++r)
            out[r] = fmaf(T[2 * n + r], v[2], fmaf(T[1 * n + r], v[1], T[0 * n + r] * v[0]));
    } else {
        for (int r = 0; r < 2; ++r) out[r] = fmaf(T[1 * n + r], v[1], T[0 * n + r] * v[0]);
    }
}

int32_t orc_transform(float* features, int32_t rows, float* normals, int64_t n, const float* T) {
    const int dim = rows - 1;
    if (!features || !T || (dim != 2 && dim != 3) || n < 0) return B200ICP_ERR_INVALID_ARG;
    if (fabsf(1.f - det_rot(T, dim)) > 1e-3f) return B200ICP_ERR_TRANSFORM;
    for (int64_t i = 0; i < n; ++i) {
        float o[3];
        apply_T_point(T, dim, features + i * rows, o);
        for (int d = 0; d < dim; ++d) features[i * rows + d] = o[d];
        if (normals) {
            apply_R_vec(T, dim, normals + i * dim, o);
            for (int d = 0; d < dim; ++d) normals[i * dim + d] = o[d];
        }
    }
    return B200ICP_OK;
}

/* Eigen LLT (unblocked, lower) in fp32, column-major n x n, in place. Returns 0 on success. */
static int llt_f32(float* A, int n) {
    for (int k = 0; k < n; ++k) {
        float x = A[k * n + k];
        for (int j = 0; j < k; ++j) x -= A[j * n + k] * A[j * n + k];
        if (!(x > 0.f)) return k + 1;
        x = sqrtf(x);
        A[k * n + k] = x;
        for (int i = k + 1; i < n; ++i) {
            float v = A[k * n + i];
            for (int j = 0; j < k; ++j) v -= A[j * n + i] * A[j * n + k];
            A[k * n + i] = v / x;
        }
    }
    return 0;
}

/* symmetric Jacobi eigen-decomposition in fp64 (fallback paths and surface normals) */
static void jacobi_eig_f64(double* A, int n, double* V, double* w) {
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

/* LPM ErrorMinimizers/PointToPlane.cpp solvePossiblyUnderdeterminedLinearSystem: when A is
 * invertible x = A.llt().solve(b); otherwise the minimum-norm least-squares solution.  The
 * invertibility test (upstream: fullPivHouseholderQr(A).isInvertible()) is restated as "the fp32
 * Cholesky succeeds with every pivot above n*eps*max|diag|"; the rank-deficient branch is the
 * fp64 pseudo-inverse (what upstream's reduced QR / JacobiSVD chain converges to). */
static void solve_normal_eq(const float* A_in, const float* b, float* x, int n) {
    float L[36];
    memcpy(L, A_in, sizeof(float) * (size_t)(n * n));
    float maxdiag = 0.f;
    for (int i = 0; i < n; ++i) maxdiag = fmaxf(maxdiag, fabsf(A_in[i * n + i]));
    int ok = (llt_f32(L, n) == 0);
    if (ok) {
        const float thr = (float)n * FLT_EPSILON * maxdiag;
        for (int i = 0; i < n; ++i)
            if (!(L[i * n + i] * L[i * n + i] > thr)) ok = 0;
    }
    if (ok) {
        float y[6];
        for (int i = 0; i < n; ++i) { /* L y = b */
            float v = b[i];
            for (int j = 0; j < i; ++j) v -= L[j * n + i] * y[j];
            y[i] = v / L[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) { /* L^T x = y */
            float v = y[i];
            for (int j = i + 1; j < n; ++j) v -= L[i * n + j] * x[j];
            x[i] = v / L[i * n + i];
        }
        int bad = 0;
        for (int i = 0; i < n; ++i)
            if (isnan(x[i])) bad = 1;
        if (!bad) return;
    }
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

/* Eigen Quaternion(Matrix3) (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl). */
static void quat_from_R(const float* T, int dim, float q[4] /* w x y z */) {
    const int n = dim + 1;
    float R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int r = 0; r < dim; ++r)
        for (int c = 0; c < dim; ++c) R[r][c] = T[c * n + r];
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
/* Eigen QuaternionBase::angularDistance: d = a * conj(b); 2*atan2(|d.vec|, |d.w|) */
static float quat_angular_distance(const float a[4], const float b[4]) {
    const float bw = b[0], bx = -b[1], by = -b[2], bz = -b[3];
    const float w = a[0] * bw - a[1] * bx - a[2] * by - a[3] * bz;
    const float x = a[0] * bx + a[1] * bw + a[2] * bz - a[3] * by;
    const float y = a[0] * by + a[2] * bw + a[3] * bx - a[1] * bz;
    const float z = a[0] * bz + a[3] * bw + a[1] * by - a[2] * bx;
    return 2.f * atan2f(sqrtf(x * x + y * y + z * z), fabsf(w));
}

/* ============================================================================================
 * PM::ICPSequence
 * ============================================================================================ */
struct orc_icp {
    b200icp_config cfg;
    int dim;
    int64_t n_map;
    float* map;     /* 4 floats per point, mean-centred */
    float* normals; /* 3 floats per point, or NULL */
    float mean[3];
    orc_kdtree* tree;
    float last_var_ratio; /* VarTrimmed: the tuned ratio of the last iteration (diagnostic) */
    float last_robust_scale; /* RobustOutlierFilter: the scale used by the last iteration (diagnostic) */
    float* reading_normals; /* dim floats per reading point for the NEXT register call (SurfaceNormalOutlierFilter), or NULL */
    int64_t n_reading_normals;
    float* reading_max_dist; /* `maxSearchDist` descriptor of the reading for the NEXT register call (one radius per point), or NULL */
    int64_t n_reading_max_dist;
    char err[256];
};

orc_icp* orc_icp_create(const b200icp_config* cfg) {
    if (!cfg || (cfg->dim != 2 && cfg->dim != 3) || cfg->knn < 1 || cfg->knn > ORC_MAXK) return NULL;
    if ((cfg->minimizer_flags & 3) == 3) return NULL; /* LPM ConfigurationError: force2D together with force4DOF */
    orc_icp* o = (orc_icp*)calloc(1, sizeof(orc_icp));
    o->cfg = *cfg;
    o->dim = cfg->dim;
    return o;
}
void orc_icp_destroy(orc_icp* o) {
    if (!o) return;
    free(o->map);
    free(o->normals);
    free(o->reading_normals);
    free(o->reading_max_dist);
    orc_kdtree_free(o->tree);
    free(o);
}
const char* orc_icp_last_error(const orc_icp* o) { return o ? o->err : "null oracle"; }
float orc_icp_last_var_ratio(const orc_icp* o) { return o ? o->last_var_ratio : 0.f; }
float orc_icp_last_robust_scale(const orc_icp* o) { return o ? o->last_robust_scale : 0.f; }
/* the `normals` descriptor of the reading handed to the next orc_icp_register (dim x n, column-major); NULL clears it */
int32_t orc_icp_set_reading_normals(orc_icp* o, const float* normals, int64_t n) {
    if (!o || n < 0) return B200ICP_ERR_INVALID_ARG;
    free(o->reading_normals);
    o->reading_normals = NULL;
    o->n_reading_normals = 0;
    if (normals && n > 0) {
        o->reading_normals = (float*)malloc((size_t)n * o->dim * sizeof(float));
        memcpy(o->reading_normals, normals, (size_t)n * o->dim * sizeof(float));
        o->n_reading_normals = n;
    }
    return B200ICP_OK;
}
/* the `maxSearchDist` descriptor of the reading handed to the next orc_icp_register / orc_icp_match (LPM KDTreeMatcher: when the
 * reading carries it, every point is searched within its own radius and the matcher's maxDist is not used); NULL clears it */
int32_t orc_icp_set_reading_max_search_dist(orc_icp* o, const float* radii, int64_t n) {
    if (!o || n < 0) return B200ICP_ERR_INVALID_ARG;
    free(o->reading_max_dist);
    o->reading_max_dist = NULL;
    o->n_reading_max_dist = 0;
    if (radii && n > 0) {
        o->reading_max_dist = (float*)malloc((size_t)n * sizeof(float));
        memcpy(o->reading_max_dist, radii, (size_t)n * sizeof(float));
        o->n_reading_max_dist = n;
    }
    return B200ICP_OK;
}
void orc_icp_get_mean(const orc_icp* o, float mean[3]) {
    for (int d = 0; d < 3; ++d) mean[d] = o->mean[d];
}

/* LPM ICP.cpp ICPSequence::setMap -- ref: Map.cpp:111,178,528,581. */
int32_t orc_icp_set_map(orc_icp* o, const float* features, int32_t rows, const float* normals,
                        int64_t n) {
    if (!o || !features || rows != o->dim + 1 || n < 0) return B200ICP_ERR_INVALID_ARG;
    if (n == 0) return B200ICP_OK; /* "Ignoring attempt to create a map from an empty cloud" */
    const int dim = o->dim;
    free(o->map);
    free(o->normals);
    orc_kdtree_free(o->tree);
    o->normals = NULL;
    o->n_map = n;
    int64_t acc[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) acc[d] += llrint((double)features[i * rows + d] * 65536.0);
    for (int d = 0; d < 3; ++d) o->mean[d] = (d < dim) ? (float)(((double)acc[d] / 65536.0) / (double)n) : 0.f;
    o->map = (float*)calloc((size_t)n * 4, sizeof(float));
    for (int64_t i = 0; i < n; ++i) {
        for (int d = 0; d < dim; ++d) o->map[i * 4 + d] = features[i * rows + d] - o->mean[d];
        o->map[i * 4 + 3] = 1.f;
    }
    if (normals) {
        o->normals = (float*)calloc((size_t)n * 3, sizeof(float));
        for (int64_t i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d) o->normals[i * 3 + d] = normals[i * dim + d];
    }
    o->tree = orc_kdtree_build(o->map, 4, n, dim);
    return B200ICP_OK;
}

/* std::nth_element on floats (Hoare quickselect, median-of-three). */
static float nth_element_f32(float* v, int64_t n, int64_t nth) {
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const int64_t mid = lo + (hi - lo) / 2;
        const float a = v[lo], b = v[mid], e = v[hi - 1];
        const float pivot = (a < b) ? ((b < e) ? b : (a < e ? e : a)) : ((a < e) ? a : (b < e ? e : b));
        int64_t i = lo, j = hi - 1;
        while (i <= j) {
            while (v[i] < pivot) ++i;
            while (v[j] > pivot) --j;
            if (i <= j) {
                const float t = v[i];
                v[i] = v[j];
                v[j] = t;
                ++i;
                --j;
            }
        }
        if (nth <= j)
            hi = j + 1;
        else if (nth >= i)
            lo = i;
        else
            break;
    }
    return v[nth];
}

/* LPM Matches.cpp getDistsQuantile: over every finite dist, value at index
 * size_t(values.size() * quantile) (the product is evaluated in fp32, T = float). */
static int dists_quantile(const float* d2, int64_t m, float quantile, float* out, float* scratch) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < m; ++i)
        if (d2[i] != INFINITY) scratch[cnt++] = d2[i];
    if (cnt == 0) return -1;
    if (quantile == 1.0f) {
        float mx = scratch[0];
        for (int64_t i = 1; i < cnt; ++i) mx = fmaxf(mx, scratch[i]);
        *out = mx;
        return 0;
    }
    int64_t idx = (int64_t)((float)cnt * quantile);
    if (idx >= cnt) idx = cnt - 1;
    *out = nth_element_f32(scratch, cnt, idx);
    return 0;
}

static int cmp_f32(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* LPM Matches::getMedianAbsDeviation: median (element n / 2 after nth_element) of the finite squared distances, then the
 * median of their absolute deviations from it. */
static int median_abs_deviation(const float* d2, int64_t m, float* out, float* scratch) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < m; ++i)
        if (d2[i] != INFINITY) scratch[cnt++] = d2[i];
    if (cnt == 0) return -1;
    const float median = nth_element_f32(scratch, cnt, cnt / 2);
    for (int64_t i = 0; i < cnt; ++i) scratch[i] = fabsf(scratch[i] - median);
    *out = nth_element_f32(scratch, cnt, cnt / 2);
    return 0;
}

/* LPM Matches::getStandardDeviation: sqrt(sum (d - mean)^2 / (size - 1)) over ALL entries of the dists matrix (an
 * unmatched entry's +inf turns it into NaN, as upstream). */
static float dists_standard_deviation(const float* d2, int64_t m) {
    double sum = 0.0, sq = 0.0;
    for (int64_t i = 0; i < m; ++i) sum += d2[i];
    const double mean = sum / (double)m;
    for (int64_t i = 0; i < m; ++i) sq += ((double)d2[i] - mean) * ((double)d2[i] - mean);
    return (float)sqrt(sq / (double)(m - 1));
}

/* LPM OutlierFiltersImpl.cpp RobustOutlierFilter::robustFiltering, weight of one match: e2 = dist / scale^2, k = tuning */
static float robust_weight(int fct, float k, float approximation, float scale, float dist) {
    const float e2 = dist / (scale * scale);
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
        case B200ICP_ROBUST_STUDENT: w = powf(1.f + e2 / k, -(k + 3.f) / 2.f) * (k + 3.f) / (k + e2); break;
        default: break;
    }
    if (w <= 0.f) w = 0.f; /* upstream floors at 1e-50, which is 0 in T = float */
    if (approximation * approximation != INFINITY && e2 >= approximation * approximation) w = 0.f;
    return w;
}

/* LPM OutlierFiltersImpl.cpp VarTrimmedDistOutlierFilter::optimizeInlierRatio: the finite, positive squared
 * distances sorted ascending, their running sum (std::partial_sum in T = float), then over the indices
 * [floor(minRatio n), floor(maxRatio n)) -- n = rows * cols of the match matrix -- the minimum of
 * FRMS_i = (1 / (id_i / n)^lambda)^2 * cumsum_i / id_i, id_i = i + 1; returns (argmin + minEl) / n.
 * Upstream maps n entries of the running-sum vector even when fewer distances qualified (reading past its end);
 * here the range is clipped to the entries that exist. */
static int var_trimmed_ratio(const float* d2, int64_t m, float min_ratio, float max_ratio, float lambda, float* ratio, float* scratch) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < m; ++i)
        if (d2[i] != INFINITY && d2[i] > 0.f) scratch[cnt++] = d2[i];
    if (cnt == 0) return -1;
    qsort(scratch, (size_t)cnt, sizeof(float), cmp_f32);
    float run = 0.f;
    for (int64_t i = 0; i < cnt; ++i) {
        run += scratch[i];
        scratch[i] = run;
    }
    int64_t min_el = (int64_t)floorf(min_ratio * (float)m), max_el = (int64_t)floorf(max_ratio * (float)m);
    if (max_el > cnt) max_el = cnt;
    if (min_el >= max_el) min_el = max_el > 0 ? max_el - 1 : 0;
    float best = INFINITY;
    int64_t best_i = 0;
    for (int64_t i = min_el; i < max_el; ++i) {
        const float id = (float)(i + 1);
        const float deno = powf(id / (float)m, lambda);
        const float inv = 1.f / deno;
        const float frms = inv * inv * scratch[i] * (1.f / id);
        if (frms < best) {
            best = frms;
            best_i = i - min_el;
        }
    }
    *ratio = (float)(best_i + min_el) / (float)m;
    return 0;
}

typedef struct {
    float q[8][4]; /* ring of the last smoothLength+1 rotations (smoothLength <= 7) */
    float t[8][3];
    int count;
} diff_state;

int32_t orc_icp_match(orc_icp* o, const float* queries, int32_t rows, int64_t nq, int32_t* ids,
                      float* d2, int32_t nthreads) {
    if (!o || !queries || rows != o->dim + 1) return B200ICP_ERR_INVALID_ARG;
    if (!o->tree) return B200ICP_ERR_NO_MAP;
    float* c = (float*)calloc((size_t)nq * 4 + 4, sizeof(float));
    for (int64_t i = 0; i < nq; ++i)
        for (int d = 0; d < o->dim; ++d) c[i * 4 + d] = queries[i * rows + d] - o->mean[d];
    orc_kdtree_knn_ex(o->tree, c, 4, nq, o->cfg.knn, o->cfg.max_dist, (o->reading_max_dist && o->n_reading_max_dist == nq) ? o->reading_max_dist : NULL,
                      o->cfg.conventions & 1, ids, d2, nthreads);
    free(c);
    return B200ICP_OK;
}

#define FAIL(o, code, msg)                                \
    do {                                                  \
        snprintf((o)->err, sizeof((o)->err), "%s", msg);  \
        rc = (code);                                      \
        goto done;                                        \
    } while (0)

/* LPM ICP.cpp ICPSequence::operator()(cloudIn) -> compute -- ref: Mapper.cpp:213. */
int32_t orc_icp_register(orc_icp* o, const float* reading_in, int32_t rows, int64_t nq,
                         const float* T_init, float* T_out, b200icp_result* result, float* trace,
                         double* secs, int32_t nthreads) {
    if (!o || !reading_in || !T_out || rows != o->dim + 1 || nq < 0) return B200ICP_ERR_INVALID_ARG;
    const int dim = o->dim, n = dim + 1, k = o->cfg.knn;
    const b200icp_config* cfg = &o->cfg;
    int32_t rc = B200ICP_OK;
    double t_match = 0, t_outl = 0, t_min = 0;
    const double t_begin = now_s();
    float Tident[16];
    mat4_identity(Tident, n);
    if (!T_init) T_init = Tident;
    if (!o->tree) { /* LPM: no map -> identity */
        memcpy(T_out, Tident, sizeof(float) * (size_t)(n * n));
        if (result) memset(result, 0, sizeof(*result));
        return B200ICP_ERR_NO_MAP;
    }
    float *reading = NULL, *step = NULL, *d2 = NULL, *w = NULL, *scratch = NULL, *rnrm = NULL;
    int32_t* ids = NULL;
    reading = (float*)calloc((size_t)nq * 4 + 4, sizeof(float));
    step = (float*)calloc((size_t)nq * 4 + 4, sizeof(float));
    ids = (int32_t*)malloc(((size_t)nq * k + 1) * sizeof(int32_t));
    d2 = (float*)malloc(((size_t)nq * k + 1) * sizeof(float));
    w = (float*)malloc(((size_t)nq * k + 1) * sizeof(float));
    scratch = (float*)malloc(((size_t)nq * k + 1) * sizeof(float));

    /* T_refMean_dataIn = T_refIn_refMean^-1 * T_refIn_dataIn */
    float Tmean[16], Tmean_inv[16], Tpre[16];
    mat4_identity(Tmean, n);
    mat4_identity(Tmean_inv, n);
    for (int d = 0; d < dim; ++d) {
        Tmean[dim * n + d] = o->mean[d];
        Tmean_inv[dim * n + d] = -o->mean[d];
    }
    matn_mul(Tmean_inv, T_init, Tpre, n);
    if (fabsf(1.f - det_rot(Tpre, dim)) > 1e-3f) FAIL(o, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
    for (int64_t i = 0; i < nq; ++i) {
        apply_T_point(Tpre, dim, reading_in + i * rows, reading + i * 4);
        reading[i * 4 + 3] = 1.f;
    }

    /* descriptors named `normals` rotate with the cloud (LPM TransformationsImpl.cpp) */
    if (o->reading_normals && o->n_reading_normals == nq) {
        rnrm = (float*)malloc((size_t)nq * 3 * sizeof(float));
        for (int64_t i = 0; i < nq; ++i) {
            float out3[3] = {0, 0, 0};
            apply_R_vec(Tpre, dim, o->reading_normals + i * dim, out3);
            for (int d = 0; d < 3; ++d) rnrm[i * 3 + d] = d < dim ? out3[d] : 0.f;
        }
    }

    float T_iter[16];
    mat4_identity(T_iter, n);
    /* transformationCheckers.init(T_iter) */
    int counter = 0;
    diff_state ds;
    memset(&ds, 0, sizeof(ds));
    float bound_q0[4], bound_t0[3] = {0, 0, 0};
    quat_from_R(T_iter, dim, ds.q[0]);
    for (int d = 0; d < 3; ++d) ds.t[0][d] = 0.f;
    ds.count = 1;
    quat_from_R(T_iter, dim, bound_q0);

    int iterate = 1, iterations = 0, max_iter_reached = 0;
    float robust_scale[B200ICP_MAX_OUTLIER_FILTERS] = {0.f, 0.f, 0.f, 0.f}; /* RobustOutlierFilter::scale, per filter */
    float overlap = 0.f, used_ratio = 0.f;
    int64_t pairs = 0;
    const int smooth = cfg->smooth_length > 7 ? 7 : (cfg->smooth_length < 1 ? 1 : cfg->smooth_length);

    while (iterate) {
        /* stepReading = T_iter * reading */
        if (fabsf(1.f - det_rot(T_iter, dim)) > 1e-3f) FAIL(o, B200ICP_ERR_TRANSFORM, "RigidTransformation: Error, rotation matrix is not orthogonal.");
        for (int64_t i = 0; i < nq; ++i) apply_T_point(T_iter, dim, reading + i * 4, step + i * 4);

        /* LPM MatchersImpl.cpp KDTreeMatcher::findClosests */
        double t0 = now_s();
        orc_kdtree_knn_ex(o->tree, step, 4, nq, k, cfg->max_dist, (o->reading_max_dist && o->n_reading_max_dist == nq) ? o->reading_max_dist : NULL,
                          cfg->conventions & 1, ids, d2, nthreads);
        t_match += now_s() - t0;

        /* LPM OutlierFiltersImpl.cpp: product of all configured filters; none -> ones */
        t0 = now_s();
        const int64_t m = nq * k;
        for (int64_t i = 0; i < m; ++i) w[i] = 1.f;
        for (int f = 0; f < cfg->n_outlier; ++f) {
            const float prm = cfg->outlier_param[f];
            float limit = 0.f;
            switch (cfg->outlier_kind[f]) {
                case B200ICP_OUTLIER_TRIMMED_DIST:
                    if (dists_quantile(d2, m, prm, &limit, scratch)) FAIL(o, B200ICP_ERR_CONVERGENCE, "no outlier to filter");
                    for (int64_t i = 0; i < m; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
                    break;
                case B200ICP_OUTLIER_MEDIAN_DIST:
                    if (dists_quantile(d2, m, 0.5f, &limit, scratch)) FAIL(o, B200ICP_ERR_CONVERGENCE, "no outlier to filter");
                    /* conventions bit 1: the factor scales the distance, i.e. factor^2 on the squared distances compared here */
                    limit = ((cfg->conventions & 2) ? prm * prm : prm) * limit;
                    for (int64_t i = 0; i < m; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
                    break;
                case B200ICP_OUTLIER_VAR_TRIMMED_DIST: {
                    float ratio = 0.f;
                    if (var_trimmed_ratio(d2, m, prm, cfg->outlier_param2[f], cfg->outlier_param3[f], &ratio, scratch))
                        FAIL(o, B200ICP_ERR_CONVERGENCE, "Inlier ratio optimization failed: no finite distances");
                    if (dists_quantile(d2, m, ratio, &limit, scratch)) FAIL(o, B200ICP_ERR_CONVERGENCE, "no outlier to filter");
                    for (int64_t i = 0; i < m; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
                    o->last_var_ratio = ratio;
                    break;
                }
                case B200ICP_OUTLIER_SURFACE_NORMAL: {
                    /* LPM SurfaceNormalOutlierFilter{maxAngle}: eps = cos(maxAngle); w = 0 when the (normalised) reading normal,
                     * moved by T_iter, and the (normalised) reference normal of the match have dot < eps, or there is no match;
                     * all ones when either cloud lacks normals ("Skipping filtering") */
                    if (!rnrm || !o->normals) break;
                    const float eps = cosf(prm);
                    for (int64_t i = 0; i < nq; ++i) {
                        float nr[3] = {0, 0, 0};
                        apply_R_vec(T_iter, dim, rnrm + i * 3, nr);
                        float ln = 0.f;
                        for (int d = 0; d < dim; ++d) ln += nr[d] * nr[d];
                        ln = sqrtf(ln);
                        for (int j = 0; j < k; ++j) {
                            const int32_t id = ids[i * k + j];
                            if (id < 0) {
                                w[i * k + j] = 0.f;
                                continue;
                            }
                            const float* nf = o->normals + (int64_t)id * 3;
                            float lf = 0.f, dot = 0.f;
                            for (int d = 0; d < dim; ++d) lf += nf[d] * nf[d];
                            lf = sqrtf(lf);
                            for (int d = 0; d < dim; ++d) dot += (nr[d] / ln) * (nf[d] / lf);
                            w[i * k + j] *= (dot < eps) ? 0.f : 1.f;
                        }
                    }
                    break;
                }
                case B200ICP_OUTLIER_ROBUST: {
                    /* LPM RobustOutlierFilter{robustFct, tuning, scaleEstimator, nbIterationForScale, distanceType, approximation}.
                     * `iteration` counts this filter's calls from 1 (restarted per registration here); the scale is re-estimated
                     * while iteration <= nbIterationForScale, or always when that is 0. */
                    const int mode = cfg->outlier_mode[f];
                    const int fct = mode & 255, est = (mode >> 8) & 15, p2plane = (mode >> 12) & 1, nb = (mode >> 16) & 0x7fff;
                    const int iteration = iterations + 1;
                    float tuning = prm;
                    if (est == B200ICP_SCALE_BERG) { /* Bergstrom & Edlund 2014: tuning is the target scale, k comes from the paper */
                        if (fct == B200ICP_ROBUST_CAUCHY) tuning = 4.3040f;
                        else if (fct == B200ICP_ROBUST_TUKEY) tuning = 7.0589f;
                        else if (fct == B200ICP_ROBUST_HUBER) tuning = 2.0138f;
                    }
                    if (iteration <= nb || nb == 0) {
                        float v = 0.f;
                        if (est == B200ICP_SCALE_MAD) {
                            if (median_abs_deviation(d2, m, &v, scratch)) FAIL(o, B200ICP_ERR_CONVERGENCE, "no outlier to filter");
                            robust_scale[f] = sqrtf(v);
                        } else if (est == B200ICP_SCALE_STD) {
                            robust_scale[f] = sqrtf(dists_standard_deviation(d2, m));
                        } else if (est == B200ICP_SCALE_BERG) {
                            if (iteration == 1) {
                                if (dists_quantile(d2, m, 0.5f, &v, scratch)) FAIL(o, B200ICP_ERR_CONVERGENCE, "no outlier to filter");
                                robust_scale[f] = 1.9f * sqrtf(v);
                            } else {
                                robust_scale[f] = 0.85f * (robust_scale[f] - prm) + prm;
                            }
                        }
                    }
                    const float scale = est == B200ICP_SCALE_NONE ? 1.f : robust_scale[f];
                    o->last_robust_scale = scale;
                    if (p2plane && !o->normals) FAIL(o, B200ICP_ERR_INVALID_FIELD, "Cannot find descriptor normals in reference");
                    for (int64_t i = 0; i < nq; ++i)
                        for (int j = 0; j < k; ++j) {
                            const int32_t id = ids[i * k + j];
                            float dist = d2[i * k + j];
                            if (p2plane) { /* computePointToPlaneDistance: (n . (p - q))^2 with n normalised; 0 where unmatched */
                                dist = 0.f;
                                if (id >= 0) {
                                    const float* nf = o->normals + (int64_t)id * 3;
                                    const float* q = o->map + (int64_t)id * 4;
                                    const float* pp = step + i * 4;
                                    float ln = 0.f, dot = 0.f;
                                    for (int d = 0; d < dim; ++d) ln += nf[d] * nf[d];
                                    ln = sqrtf(ln);
                                    for (int d = 0; d < dim; ++d) dot += (nf[d] / ln) * (pp[d] - q[d]);
                                    dist = dot * dot;
                                }
                            }
                            w[i * k + j] *= robust_weight(fct, tuning, cfg->outlier_param2[f], scale, dist);
                        }
                    break;
                }
                case B200ICP_OUTLIER_MAX_DIST:
                    limit = prm * prm;
                    for (int64_t i = 0; i < m; ++i) w[i] *= (d2[i] <= limit) ? 1.f : 0.f;
                    break;
                case B200ICP_OUTLIER_MIN_DIST:
                    limit = prm * prm;
                    for (int64_t i = 0; i < m; ++i) w[i] *= (d2[i] >= limit) ? 1.f : 0.f;
                    break;
                default:
                    FAIL(o, B200ICP_ERR_INVALID_ARG, "unknown outlier filter");
            }
        }
        t_outl += now_s() - t0;

        /* LPM ErrorMinimizer.cpp ErrorElements: every (k, i) with finite dist and w != 0 is a
         * pair; loop order is k outer, i inner. */
        t0 = now_s();
        float dT[16];
        mat4_identity(dT, n);
        pairs = 0;
        float wsum = 0.f;
        if (cfg->minimizer == B200ICP_MIN_POINT_TO_PLANE) {
            if (!o->normals) FAIL(o, B200ICP_ERR_INVALID_FIELD, "Cannot find descriptor normals in reference");
            /* LPM ErrorMinimizers/PointToPlane.cpp compute_in_place */
            /* options force2D / force4DOF (3-D clouds only; mutually exclusive, checked at creation): force2D cuts both clouds
             * down to x, y (cross = p_x n_y - p_y n_x, F = [cross; n_x; n_y], residual without z), force4DOF keeps 3-D points but
             * only rotates about z (cross = ((Gamma p)^T n), Gamma = [0 -1 0; 1 0 0; 0 0 0], F = [cross; n_x; n_y; n_z]) */
            const int force2d = dim == 3 && (cfg->minimizer_flags & 1), force4dof = dim == 3 && (cfg->minimizer_flags & 2);
            const int ns = force2d ? 3 : (force4dof ? 4 : ((dim == 3) ? 6 : 3));
            float A[36], b[6], x[6];
            memset(A, 0, sizeof(A));
            memset(b, 0, sizeof(b));
            for (int kk = 0; kk < k; ++kk)
                for (int64_t i = 0; i < nq; ++i) {
                    const float dd = d2[i * k + kk], ww = w[i * k + kk];
                    if (dd == INFINITY || ww == 0.f) continue;
                    const float* p = step + i * 4;
                    const float* q = o->map + (int64_t)ids[i * k + kk] * 4;
                    const float* nr = o->normals + (int64_t)ids[i * k + kk] * 3;
                    float F[6];
                    float dot;
                    if (force2d) {
                        F[0] = p[0] * nr[1] - p[1] * nr[0];
                        F[1] = nr[0];
                        F[2] = nr[1];
                        dot = (p[0] - q[0]) * nr[0] + (p[1] - q[1]) * nr[1];
                    } else if (force4dof) {
                        F[0] = p[0] * nr[1] - p[1] * nr[0];
                        F[1] = nr[0];
                        F[2] = nr[1];
                        F[3] = nr[2];
                        dot = (p[0] - q[0]) * nr[0] + (p[1] - q[1]) * nr[1] + (p[2] - q[2]) * nr[2];
                    } else if (dim == 3) {
                        F[0] = p[1] * nr[2] - p[2] * nr[1];
                        F[1] = p[2] * nr[0] - p[0] * nr[2];
                        F[2] = p[0] * nr[1] - p[1] * nr[0];
                        F[3] = nr[0];
                        F[4] = nr[1];
                        F[5] = nr[2];
                        dot = (p[0] - q[0]) * nr[0] + (p[1] - q[1]) * nr[1] + (p[2] - q[2]) * nr[2];
                    } else {
                        F[0] = p[0] * nr[1] - p[1] * nr[0];
                        F[1] = nr[0];
                        F[2] = nr[1];
                        dot = (p[0] - q[0]) * nr[0] + (p[1] - q[1]) * nr[1];
                    }
                    for (int c = 0; c < ns; ++c) {
                        const float wf = ww * F[c];
                        for (int r = 0; r < ns; ++r) A[c * ns + r] += wf * F[r];
                        b[c] -= wf * dot;
                    }
                    wsum += ww;
                    ++pairs;
                }
            if (pairs == 0) FAIL(o, B200ICP_ERR_CONVERGENCE, "ErrorMnimizer: no point to minimize");
            solve_normal_eq(A, b, x, ns);
            if (force2d || force4dof) {
                /* Rotation2D(x0) in the top-left corner of an identity (force2D) / AngleAxis(x0, unitZ) (force4DOF) */
                const float sn = sinf(x[0]), cs = cosf(x[0]);
                dT[0] = cs;
                dT[1] = sn;
                dT[4] = -sn;
                dT[5] = cs;
                dT[12] = x[1];
                dT[13] = x[2];
                if (force4dof) dT[14] = x[3];
            } else if (dim == 3) {
                /* Eigen AngleAxis(|x|, x/|x|).toRotationMatrix(); NaN -> identity */
                const float nrm2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
                const float ang = sqrtf(nrm2);
                float ax[3] = {x[0], x[1], x[2]};
                if (nrm2 > 0.f)
                    for (int d = 0; d < 3; ++d) ax[d] = x[d] / ang;
                const float s = sinf(ang), c = cosf(ang);
                const float sx = s * ax[0], sy = s * ax[1], sz = s * ax[2];
                const float cx = (1.f - c) * ax[0], cy = (1.f - c) * ax[1], cz = (1.f - c) * ax[2];
                float R[3][3];
                float tmp = cx * ax[1];
                R[0][1] = tmp - sz;
                R[1][0] = tmp + sz;
                tmp = cx * ax[2];
                R[0][2] = tmp + sy;
                R[2][0] = tmp - sy;
                tmp = cy * ax[2];
                R[1][2] = tmp - sx;
                R[2][1] = tmp + sx;
                R[0][0] = cx * ax[0] + c;
                R[1][1] = cy * ax[1] + c;
                R[2][2] = cz * ax[2] + c;
                int bad = 0;
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 3; ++cc)
                        if (isnan(R[r][cc])) bad = 1;
                for (int r = 0; r < 3; ++r)
                    for (int cc = 0; cc < 3; ++cc) dT[cc * 4 + r] = bad ? (r == cc ? 1.f : 0.f) : R[r][cc];
                for (int d = 0; d < 3; ++d) dT[3 * 4 + d] = x[3 + d];
            } else {
                const float s = sinf(x[0]), c = cosf(x[0]);
                dT[0] = c;
                dT[1] = s;
                dT[3] = -s;
                dT[4] = c;
                dT[6] = x[1];
                dT[7] = x[2];
            }
        } else if (cfg->minimizer == B200ICP_MIN_POINT_TO_POINT) {
            /* LPM ErrorMinimizers/PointToPoint.cpp compute_in_place */
            float sw = 0.f, mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
            for (int kk = 0; kk < k; ++kk)
                for (int64_t i = 0; i < nq; ++i) {
                    const float dd = d2[i * k + kk], ww = w[i * k + kk];
                    if (dd == INFINITY || ww == 0.f) continue;
                    const float* p = step + i * 4;
                    const float* q = o->map + (int64_t)ids[i * k + kk] * 4;
                    sw += ww;
                    for (int d = 0; d < dim; ++d) {
                        mp[d] += ww * p[d];
                        mq[d] += ww * q[d];
                    }
                    ++pairs;
                }
            if (pairs == 0) FAIL(o, B200ICP_ERR_CONVERGENCE, "ErrorMnimizer: no point to minimize");
            wsum = sw;
            const float inv = 1.f / sw;
            for (int d = 0; d < dim; ++d) {
                mp[d] *= inv;
                mq[d] *= inv;
            }
            double M[9];
            memset(M, 0, sizeof(M));
            for (int kk = 0; kk < k; ++kk)
                for (int64_t i = 0; i < nq; ++i) {
                    const float dd = d2[i * k + kk], ww = w[i * k + kk];
                    if (dd == INFINITY || ww == 0.f) continue;
                    const float* p = step + i * 4;
                    const float* q = o->map + (int64_t)ids[i * k + kk] * 4;
                    for (int r = 0; r < dim; ++r)
                        for (int c = 0; c < dim; ++c)
                            M[c * dim + r] += (double)((q[r] - mq[r]) * ww * (p[c] - mp[c]));
                }
            /* SVD of M through the symmetric eigen-problems (fp64): M = U S V^T,
             * R = U V^T, flipping the last singular direction when det < 0. */
            double MtM[9], V[9], ev[3];
            for (int r = 0; r < dim; ++r)
                for (int c = 0; c < dim; ++c) {
                    double a = 0;
                    for (int j = 0; j < dim; ++j) a += M[r * dim + j] * M[c * dim + j];
                    MtM[c * dim + r] = a;
                }
            jacobi_eig_f64(MtM, dim, V, ev);
            /* sort eigenpairs descending */
            int order[3] = {0, 1, 2};
            for (int a = 0; a < dim; ++a)
                for (int bb = a + 1; bb < dim; ++bb)
                    if (ev[order[bb]] > ev[order[a]]) {
                        int tt = order[a];
                        order[a] = order[bb];
                        order[bb] = tt;
                    }
            double U[9], Vs[9];
            for (int e = 0; e < dim; ++e) {
                const int src = order[e];
                for (int r = 0; r < dim; ++r) Vs[e * dim + r] = V[src * dim + r];
            }
            /* make V a proper rotation-or-reflection basis; U_e = M v_e / sigma_e, last column by
             * orthogonal completion so that rank-deficient M still yields an orthonormal U */
            for (int e = 0; e < dim; ++e) {
                double u[3] = {0, 0, 0}, nn = 0;
                for (int r = 0; r < dim; ++r) {
                    for (int c = 0; c < dim; ++c) u[r] += M[c * dim + r] * Vs[e * dim + c];
                    nn += u[r] * u[r];
                }
                nn = sqrt(nn);
                for (int r = 0; r < dim; ++r) U[e * dim + r] = (nn > 0) ? u[r] / nn : 0.0;
            }
            if (dim == 3) { /* recompute the weakest direction as a cross product, sign fixed below */
                double* u0 = U;
                double* u1 = U + 3;
                double* u2 = U + 6;
                double cx[3] = {u0[1] * u1[2] - u0[2] * u1[1], u0[2] * u1[0] - u0[0] * u1[2], u0[0] * u1[1] - u0[1] * u1[0]};
                const double sgn = (cx[0] * u2[0] + cx[1] * u2[1] + cx[2] * u2[2]) < 0 ? -1.0 : 1.0;
                for (int r = 0; r < 3; ++r) u2[r] = sgn * cx[r];
            } else {
                double* u0 = U;
                double* u1 = U + 2;
                double px[2] = {-u0[1], u0[0]};
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
                double det = (dim == 2) ? R[0] * R[3] - R[2] * R[1]
                                        : R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) +
                                              R[6] * (R[1] * R[5] - R[4] * R[2]);
                if (det >= 0) break;
                for (int c = 0; c < dim; ++c) Vs[(dim - 1) * dim + c] = -Vs[(dim - 1) * dim + c];
            }
            for (int r = 0; r < dim; ++r)
                for (int c = 0; c < dim; ++c) dT[c * n + r] = (float)R[c * dim + r];
            for (int r = 0; r < dim; ++r) {
                float a = 0.f;
                for (int c = 0; c < dim; ++c) a += dT[c * n + r] * mp[c];
                dT[dim * n + r] = mq[r] - a;
            }
        } else { /* IdentityErrorMinimizer -- ref: examples/config.yaml:62-63 */
            for (int64_t i = 0; i < m; ++i)
                if (d2[i] != INFINITY && w[i] != 0.f) {
                    ++pairs;
                    wsum += w[i];
                }
            if (pairs == 0) FAIL(o, B200ICP_ERR_CONVERGENCE, "ErrorMnimizer: no point to minimize");
        }
        used_ratio = (float)pairs / (float)(k * nq);
        overlap = wsum / (float)(k * nq);
        matn_mul(dT, T_iter, T_iter, n); /* T_iter = dT * T_iter */
        t_min += now_s() - t0;
        if (trace) memcpy(trace + (size_t)iterations * n * n, T_iter, sizeof(float) * (size_t)(n * n));
        ++iterations;

        /* LPM TransformationCheckersImpl.cpp: checkers run in YAML order; Counter throws
         * MaxNumIterationsReached which the loop catches (iterate = false). */
        int all_ok = 1, counter_fired = 0;
        for (int phase = 0; phase < 2 && !counter_fired; ++phase) {
            /* phase 0: the checkers listed before the Counter (cfg->checker_order bits), then the Counter, phase 1: the rest */
            if (phase == 1 && cfg->max_iteration_count > 0) {
                ++counter;
                if (counter >= cfg->max_iteration_count) {
                    max_iter_reached = 1;
                    counter_fired = 1;
                    break;
                }
            }
            if (cfg->use_differential && (((cfg->checker_order & 1) != 0) == (phase == 0))) {
                const int slot = ds.count % 8;
                quat_from_R(T_iter, dim, ds.q[slot]);
                for (int d = 0; d < 3; ++d) ds.t[slot][d] = (d < dim) ? T_iter[dim * n + d] : 0.f;
                ds.count++;
                float vr = 0.f, vt = 0.f;
                if (ds.count > smooth) {
                    for (int j = ds.count - 1; j >= ds.count - smooth; --j) {
                        const int a = j % 8, bq = (j - 1) % 8;
                        vr += fabsf(quat_angular_distance(ds.q[a], ds.q[bq]));
                        const float dx = ds.t[a][0] - ds.t[bq][0], dy = ds.t[a][1] - ds.t[bq][1], dz = ds.t[a][2] - ds.t[bq][2];
                        vt += fabsf(sqrtf(dx * dx + dy * dy + dz * dz));
                    }
                    vr /= (float)smooth;
                    vt /= (float)smooth;
                    if (vr < cfg->min_diff_rot_err && vt < cfg->min_diff_trans_err) all_ok = 0;
                }
                if (isnan(vr)) FAIL(o, B200ICP_ERR_NAN, "abs rotation norm not a number");
                if (isnan(vt)) FAIL(o, B200ICP_ERR_NAN, "abs translation norm not a number");
            }
            if (cfg->use_bound && (((cfg->checker_order & 2) != 0) == (phase == 0))) {
                float qc[4];
                quat_from_R(T_iter, dim, qc);
                const float vr = quat_angular_distance(qc, bound_q0);
                float vt = 0.f;
                for (int d = 0; d < dim; ++d) vt += (T_iter[dim * n + d] - bound_t0[d]) * (T_iter[dim * n + d] - bound_t0[d]);
                vt = sqrtf(vt);
                if (isnan(vr)) FAIL(o, B200ICP_ERR_NAN, "abs rotation norm not a number");
                if (isnan(vt)) FAIL(o, B200ICP_ERR_NAN, "abs translation norm not a number");
                if (vr > cfg->max_rotation_norm || vt > cfg->max_translation_norm) FAIL(o, B200ICP_ERR_BOUND, "limit out of bounds");
            }
        }
        if (counter_fired) {
            iterate = 0;
            break;
        }
        iterate = all_ok;
        if (cfg->max_iteration_count <= 0 && !cfg->use_differential) iterate = 0; /* no checker: LPM would loop forever */
    }

    /* return T_refIn_refMean * T_iter * T_refMean_dataIn */
    {
        float tmp[16];
        matn_mul(T_iter, Tpre, tmp, n);
        matn_mul(Tmean, tmp, T_out, n);
    }
    if (result) {
        result->overlap = overlap;
        result->point_used_ratio = used_ratio;
        result->iterations = iterations;
        result->max_iter_reached = max_iter_reached;
        result->pairs_last_iter = pairs;
    }
done:
    if (secs) {
        secs[0] = t_match;
        secs[1] = t_outl;
        secs[2] = t_min;
        secs[3] = now_s() - t_begin;
    }
    free(reading);
    free(step);
    free(ids);
    free(d2);
    free(w);
    free(scratch);
    free(rnrm);
    return rc;
}

/* ============================================================================================
 * Map-update pieces
 * ============================================================================================ */

/* ref: MapperModules/PointDistanceMapperModule.cpp:28-50 -- fresh kd-tree on the map, 1-NN with
 * eps 0 and no radius, keep input points whose squared distance >= minDistNewPoint^2. */
int64_t orc_point_distance_keep(const float* map_feat, int32_t rows, int64_t n_map,
                                const float* input_feat, int64_t n_in, float min_dist_new_point,
                                uint8_t* keep, int32_t nthreads) {
    const int dim = rows - 1;
    orc_kdtree* t = orc_kdtree_build(map_feat, rows, n_map, dim);
    if (!t) return -1;
    int32_t* ids = (int32_t*)malloc((size_t)(n_in + 1) * sizeof(int32_t));
    float* d2 = (float*)malloc((size_t)(n_in + 1) * sizeof(float));
    orc_kdtree_knn(t, input_feat, rows, n_in, 1, INFINITY, ids, d2, nthreads);
    const float thr = powf(min_dist_new_point, 2.f);
    int64_t kept = 0;
    for (int64_t i = 0; i < n_in; ++i) {
        keep[i] = (d2[i] >= thr) ? 1 : 0;
        kept += keep[i];
    }
    free(ids);
    free(d2);
    orc_kdtree_free(t);
    return kept;
}

/* LPM DataPointsFilters/SurfaceNormal.cpp (ref: examples/config.yaml:26-27 via Map.cpp:524):
 * self k-NN (the point is its own first neighbour), mean-centred covariance of the finite
 * neighbours, normal = eigenvector of the smallest eigenvalue (unit, sign arbitrary). */
int32_t orc_surface_normals(const float* feat, int32_t rows, int64_t n, int32_t knn,
                            float* normals, int32_t nthreads) {
    const int dim = rows - 1;
    if (!feat || !normals || (dim != 2 && dim != 3) || knn < 1 || knn > ORC_MAXK) return B200ICP_ERR_INVALID_ARG;
    orc_kdtree* t = orc_kdtree_build(feat, rows, n, dim);
    if (!t) return B200ICP_ERR_INVALID_ARG;
    int32_t* ids = (int32_t*)malloc((size_t)(n * knn + 1) * sizeof(int32_t));
    float* d2 = (float*)malloc((size_t)(n * knn + 1) * sizeof(float));
    orc_kdtree_knn(t, feat, rows, n, knn, INFINITY, ids, d2, nthreads);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < n; ++i) {
        float mean[3] = {0, 0, 0};
        int cnt = 0;
        for (int j = 0; j < knn; ++j) {
            const int32_t id = ids[i * knn + j];
            if (id < 0) continue;
            for (int d = 0; d < dim; ++d) mean[d] += feat[(int64_t)id * rows + d];
            ++cnt;
        }
        for (int d = 0; d < dim; ++d) mean[d] /= (float)cnt;
        double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, V[9], ev[3];
        for (int j = 0; j < knn; ++j) {
            const int32_t id = ids[i * knn + j];
            if (id < 0) continue;
            float df[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d) df[d] = feat[(int64_t)id * rows + d] - mean[d];
            for (int r = 0; r < dim; ++r)
                for (int c = 0; c < dim; ++c) C[c * dim + r] += (double)(df[r] * df[c]);
        }
        jacobi_eig_f64(C, dim, V, ev);
        int best = 0;
        for (int e = 1; e < dim; ++e)
            if (ev[e] < ev[best]) best = e;
        for (int d = 0; d < dim; ++d) normals[i * dim + d] = (float)V[best * dim + d];
    }
    free(ids);
    free(d2);
    orc_kdtree_free(t);
    return B200ICP_OK;
}
