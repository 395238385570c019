/*
 * b200icp.h -- C-ABI of libb200icp.so, the B200 (sm_100a) ICP registration core that
 * drops in behind norlab_icp_mapper's Mapper::processInput / Map::updatePose path.
 *
 * The reference has no FFI: its seams are C++ objects from libpointmatcher (`PM::ICPSequence`,
 * `PM::Transformation`, `Nabo::NNS`) and its own `MapperModule` plugin interface.  Every entry
 * point below names the reference call site it replaces (file:line under the reference tree);
 * INTEGRATION.md shows the C++ shim a maintainer would put between those call sites and this
 * header.  Conventions shared by every function:
 *
 *   - All matrices are column-major fp32, exactly the memory of the reference's Eigen types:
 *     a cloud's `features` is (dim+1) x N with the homogeneous row last, `normals` is dim x N,
 *     a TransformationParameters is (dim+1) x (dim+1).
 *   - Plain pointers and sizes only.  "_device" variants take CUDA device pointers that live on
 *     the context's device; the unsuffixed variants take host pointers and do the copies.
 *   - Every function returns a b200icp_status; the message for the last failure on a context is
 *     returned by b200icp_last_error().  No C++ exception crosses this boundary.
 *   - A context is bound to one device and one stream.  Calls on one context must be serialised
 *     by the caller -- the reference already does this with `icpMapLock` (Mapper.cpp:212,
 *     Map.cpp:110,177,527,580).  Different contexts are independent.
 *   - There is no CPU fallback: if CUDA is unavailable b200icp_create fails with
 *     B200ICP_ERR_CUDA.
 */
#ifndef B200ICP_H
#define B200ICP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ICP_ABI_VERSION 6

typedef enum b200icp_status {
    B200ICP_OK = 0,
    B200ICP_ERR_INVALID_ARG = 1,   /* bad pointer / size / unknown enum                               */
    B200ICP_ERR_CUDA = 2,          /* CUDA runtime failure, message in last_error                     */
    B200ICP_ERR_NO_MAP = 3,        /* register/match before set_map                                   */
    B200ICP_ERR_CONVERGENCE = 4,   /* LPM ConvergenceError: "no point to minimize", empty quantile    */
    B200ICP_ERR_BOUND = 5,         /* LPM BoundTransformationChecker: "limit out of bounds"           */
    B200ICP_ERR_NAN = 6,           /* LPM checkers: "abs rotation/translation norm not a number"      */
    B200ICP_ERR_TRANSFORM = 7,     /* LPM TransformationError: rotation block is not orthonormal      */
    B200ICP_ERR_INVALID_FIELD = 8, /* LPM InvalidField: e.g. point-to-plane without map normals       */
    B200ICP_ERR_NOT_IMPLEMENTED = 9
} b200icp_status;

/* icp.outlierFilters entries (libpointmatcher OutlierFiltersImpl); weights multiply. */
typedef enum b200icp_outlier_kind {
    B200ICP_OUTLIER_TRIMMED_DIST = 1, /* param = ratio;   w = dist2 <= quantile(finite dist2, ratio)  */
    B200ICP_OUTLIER_MAX_DIST = 2,     /* param = maxDist; w = dist2 <= maxDist^2                      */
    B200ICP_OUTLIER_MIN_DIST = 3,     /* param = minDist; w = dist2 >= minDist^2                      */
    B200ICP_OUTLIER_MEDIAN_DIST = 4,  /* param = factor;  w = dist2 <= factor * median(finite dist2)  */
    B200ICP_OUTLIER_VAR_TRIMMED_DIST = 5, /* VarTrimmedDistOutlierFilter{minRatio = param, maxRatio = param2, lambda = param3}:
                                            ratio = argmin FRMS over the sorted finite positive dist2 (optimizeInlierRatio),
                                            then w = dist2 <= quantile(finite dist2, ratio)                               */
    B200ICP_OUTLIER_SURFACE_NORMAL = 6, /* SurfaceNormalOutlierFilter{maxAngle = param, rad}: w = 0 when the reading's normal (moved
                                           with the reading) and the matched map point's normal, both normalised, have a dot
                                           product below cos(maxAngle); all ones when either cloud has no normals (LPM: "Skipping
                                           filtering").  The reading's normals come in through b200icp_register_normals.        */
    B200ICP_OUTLIER_ROBUST = 7          /* RobustOutlierFilter{tuning = param, approximation = param2 (+inf: none), robustFct /
                                           scaleEstimator / distanceType / nbIterationForScale = outlier_mode, see
                                           B200ICP_ROBUST_MODE}: M-estimator weights w(e2), e2 = dist / scale^2 with dist the
                                           match's squared distance (point2point) or (n . (p - q))^2 with the map point's unit
                                           normal (point2plane); scale = 1 (none), sqrt(median |d2 - median d2|) (mad),
                                           sqrt(std d2) (std), or Bergstrom's schedule (berg: 1.9 sqrt(median d2) at the first
                                           iteration, then 0.85 (scale - tuning) + tuning; the weight function's constant becomes
                                           4.3040 / 7.0589 / 2.0138 for cauchy / tukey / huber).  The scale is re-estimated while
                                           iteration <= nbIterationForScale, or always when that is 0; the iteration count
                                           restarts at 1 with every registration.  w = 0 where e2 >= approximation^2.             */
} b200icp_outlier_kind;

/* RobustOutlierFilter: robustFct, scaleEstimator and distanceType names (libpointmatcher's parameter values) */
typedef enum b200icp_robust_fct {
    B200ICP_ROBUST_CAUCHY = 0, /* 1 / (1 + e2 / k^2)                          k = tuning */
    B200ICP_ROBUST_WELSCH = 1, /* exp(-e2 / k^2)                                          */
    B200ICP_ROBUST_SC = 2,     /* e2 >= k ? 4 k^2 / (k + e2)^2 : 1    (switchable constraint) */
    B200ICP_ROBUST_GM = 3,     /* k^2 / (k + e2)^2                       (Geman-McClure)  */
    B200ICP_ROBUST_TUKEY = 4,  /* e2 >= k^2 ? 0 : (1 - e2 / k^2)^2                       */
    B200ICP_ROBUST_HUBER = 5,  /* e2 >= k^2 ? k / sqrt(e2) : 1                           */
    B200ICP_ROBUST_L1 = 6,     /* 1 / sqrt(e2)                                            */
    B200ICP_ROBUST_STUDENT = 7 /* (k + 3) (1 + e2 / k)^(-(k + 3) / 2) / (k + e2)          */
} b200icp_robust_fct;
typedef enum b200icp_robust_scale { B200ICP_SCALE_NONE = 0, B200ICP_SCALE_MAD = 1, B200ICP_SCALE_BERG = 2, B200ICP_SCALE_STD = 3 } b200icp_robust_scale;
typedef enum b200icp_robust_dist { B200ICP_DIST_POINT2POINT = 0, B200ICP_DIST_POINT2PLANE = 1 } b200icp_robust_dist;
/* outlier_mode of a RobustOutlierFilter entry: bits 0..7 robustFct, 8..11 scaleEstimator, 12 distanceType, 16..30 nbIterationForScale */
#define B200ICP_ROBUST_MODE(fct, scale, dist, nb_iteration_for_scale) \
    ((int32_t)(fct) | ((int32_t)(scale) << 8) | ((int32_t)(dist) << 12) | ((int32_t)(nb_iteration_for_scale) << 16))

/* icp.errorMinimizer */
typedef enum b200icp_minimizer_kind {
    B200ICP_MIN_POINT_TO_PLANE = 0,
    B200ICP_MIN_POINT_TO_POINT = 1,
    B200ICP_MIN_IDENTITY = 2 /* examples/config.yaml:62-63 */
} b200icp_minimizer_kind;

#define B200ICP_MAX_OUTLIER_FILTERS 4

/*
 * The `icp:` YAML node of the reference (Mapper.cpp:70-78 -> PM::ICPSequence::loadFromYamlNode),
 * flattened.  Field names follow the libpointmatcher parameter names.
 */
typedef struct b200icp_config {
    int32_t dim; /* 2 or 3 (Mapper ctor `is3D`, Mapper.h:53)                                        */
    /* matcher: KDTreeMatcher (docs/MapperConfiguration.md:175-179) */
    int32_t knn;       /* 1..32                                                                     */
    float max_dist;    /* metres; +inf = unbounded (LPM default)                                    */
    float epsilon;     /* accepted for config compatibility; the search is always exact (eps = 0)   */
    /* outlierFilters (product of all) */
    int32_t n_outlier;
    int32_t outlier_kind[B200ICP_MAX_OUTLIER_FILTERS];
    float outlier_param[B200ICP_MAX_OUTLIER_FILTERS];
    float outlier_param2[B200ICP_MAX_OUTLIER_FILTERS]; /* second / third parameter of filters that have them (VarTrimmed) */
    float outlier_param3[B200ICP_MAX_OUTLIER_FILTERS];
    /* errorMinimizer */
    int32_t minimizer;
    /* transformationCheckers */
    int32_t max_iteration_count; /* CounterTransformationChecker; <=0 means absent                */
    int32_t use_differential;    /* DifferentialTransformationChecker                             */
    float min_diff_rot_err;
    float min_diff_trans_err;
    int32_t smooth_length;
    int32_t use_bound; /* BoundTransformationChecker                                              */
    float max_rotation_norm;
    float max_translation_norm;
    /* implementation knobs (no reference counterpart) */
    int32_t sort_reading; /* 1: Morton-sort the reading once per registration (default)          */
    int32_t use_graph;    /* reserved, ignored: the persistent loop kernel made graph replay pointless */
    int32_t nn_variant;   /* 0: default.  Test / development switches, results never change beyond summation order:
                             bit 1 (2)   kernel-per-step path: cold search every iteration (no warm bound)
                             bit 2 (4)   kernel-per-step path instead of the persistent loop kernel
                             bit 3 (8)   loop kernel: general three-barrier quantile with global radix passes
                             bit 4 (16)  loop kernel: no predicted quantile window (two barriers per iteration)
                             bit 5 (32)  loop kernel: no match-cache verification (every query searched every iteration)
                             bit 6 (64)  loop kernel: no histogram-derived window (with bit 4: three barriers)
                             bit 7 (128) loop kernel: work list in plain entry order (no cost classes); k > 1: shell walk instead of the ball pass
                             bit 8 (256) cold k = 1 search: shell-walk kernel even when maxDist is small
                             bit 9 (512) loop kernel: deal the reading to the CTAs in chunks of 8 points whatever its size
                             bits 10..11 loop kernel: chunk size 1, 2 or 16 points (values 1, 2, 3)
                             bits 12..14 reading sort key: value - 1 = block shift (1 = full cell id; default shift 3)
                             bit 17 (0x20000) loop kernel: no fine histogram in the two-barrier iteration (window = the whole level-0 bucket)
                             bit 20 (0x100000) self k-NN (SurfaceNormal filters): TMA-staged candidate tiles (csrc/selfknn.cu) instead of the shell walk
                             bit 21 (0x200000) SurfaceNormal's neighbour search: no speculative one-cell bound
                             bit 22 (0x400000) loop kernel: no one-query-per-warp search for tiny work lists
                             bit 23 (0x800000) loop kernel, k > 1: collect-then-rank selection instead of the sorted-list insertion (slower; A/B)
                             bit 25 (0x2000000) PointDistance insert: unbounded 1-NN instead of the search bounded by minDistNewPoint
                             bit 26 (0x4000000) RobustOutlierFilter chains on the kernel-per-step path (scale from two device sorts) instead of the loop kernel
                             bit 27 (0x8000000) loop kernel, RobustOutlierFilter: no one-barrier (predicted window) median select
                             bit 19 (0x80000) loop kernel: write the per-iteration development record (tools/gpu_loop_record.py) */
    int32_t outlier_mode[B200ICP_MAX_OUTLIER_FILTERS]; /* per filter: B200ICP_ROBUST_MODE(...) for RobustOutlierFilter, else 0 */
    int32_t checker_order; /* position of the Counter in the transformationCheckers list.  libpointmatcher runs the checkers in YAML
                              order and the Counter reports its limit by throwing MaxNumIterationsReached, which skips the checkers
                              listed after it for that iteration: bit 0 = Differential is listed BEFORE the Counter, bit 1 = Bound is.
                              0 = Counter first (the order of PM::ICPSequence::setDefault)                                         */
    int32_t minimizer_flags; /* PointToPlaneErrorMinimizer options (LPM ErrorMinimizers/PointToPlane.cpp): bit 0 = force2D (3-D clouds:
                                solve for yaw + x, y only, z / roll / pitch untouched), bit 1 = force4DOF (yaw + x, y, z)           */
    int32_t conventions;     /* choices between two readings of upstream behaviour that cannot be settled without its source here
                                (SURVEY App. A items marked "(?)"); 0 = the reading the oracle is written to:
                                bit 0  matcher accepts a neighbour when dist2 <  maxDist^2   (default: <=)
                                bit 1  MedianDistOutlierFilter: limit = factor^2 * median(dist2), i.e. the factor scales the
                                       distance, not its square                               (default: factor * median(dist2)) */
    int32_t reserved2[5];
} b200icp_config;

/* What `icp(input)` leaves behind for the caller (Mapper.cpp:213,219). */
typedef struct b200icp_result {
    float overlap;            /* errorMinimizer->getOverlap() == weightedPointUsedRatio (last iter) */
    float point_used_ratio;   /* ErrorElements::pointUsedRatio of the last iteration              */
    int32_t iterations;       /* iterations executed                                              */
    int32_t max_iter_reached; /* Counter checker fired                                            */
    int64_t pairs_last_iter;  /* number of pairs kept by the last iteration                       */
} b200icp_result;

/* Device timings of the last b200icp_register* call, measured with CUDA events on ctx's stream. */
typedef struct b200icp_timing {
    float total_ms;       /* whole registration, first kernel to pose ready                        */
    float nn_ms_sum;      /* sum over iterations of the correspondence-search kernel (profiling on) */
    int32_t nn_launches;  /* how many launches nn_ms_sum covers                                    */
    int32_t kernel_launches; /* kernels launched by the call (all of them ours + the sort)         */
    float setmap_ms;      /* last b200icp_set_map*: mean-centre + index build                      */
    float select_ms_sum;  /* sum over iterations of the quantile-select kernel (profiling on)      */
    float acc_ms_sum;     /* sum over iterations of the accumulate/solve kernel (profiling on)     */
    int32_t loop_iterations; /* persistent loop kernel: iterations whose search phase was timed    */
    float loop_search_ms_sum; /* ... sum of the in-kernel correspondence-search phase (%globaltimer, CTA 0) */
    float loop_total_ms;      /* ... first iteration start to last iteration end (%globaltimer)        */
    int32_t loop_fast_iterations; /* ... iterations that ran with ONE device-wide barrier (predicted quantile window) */
    int32_t loop_searched_queries; /* ... queries (summed over the iterations) that needed a search; the others were proven unchanged */
    int32_t loop_two_barrier_iterations; /* ... iterations that ran with TWO barriers (quantile bucket found by a histogram pass);
                                            RobustOutlierFilter chains add their scale selects that needed more than one barrier */
    float loop_kernel_ms;     /* ... duration of the loop kernel's launch, CUDA events on ctx's stream */
} b200icp_timing;

/* One entry of the YAML `input:` chain (libpointmatcher DataPointsFilters used on this path,
 * examples/config.yaml:1-23, Mapper.cpp:27-31). */
typedef enum b200icp_filter_kind {
    B200ICP_FILTER_BOUNDING_BOX = 1,   /* BoundingBoxDataPointsFilter{xMin..zMax, removeInside}             */
    B200ICP_FILTER_DISTANCE_LIMIT = 2, /* DistanceLimitDataPointsFilter{dim (-1 radial), dist, removeInside} */
    B200ICP_FILTER_RANDOM_SAMPLING = 3 /* RandomSamplingDataPointsFilter{prob}: `dist` = prob, `dim` = seed.  A point survives
                                          when u(seed, chain slot, point index) < prob, u from a counter-based generator:
                                          reproducible; upstream draws from rand(), so only the distribution can agree      */
} b200icp_filter_kind;
typedef struct b200icp_filter {
    int32_t kind;
    float lo[3], hi[3];    /* bounding box: a point is inside when every coordinate is in [lo, hi]         */
    int32_t dim;           /* distance limit: -1 = radial, else the axis                                   */
    float dist;
    int32_t remove_inside; /* 1: drop the points inside the box / closer than dist; 0: drop the others      */
} b200icp_filter;

typedef struct b200icp_ctx b200icp_ctx;

int32_t b200icp_abi_version(void);

/* Fill *cfg with: knn 1, maxDist inf, eps 0, Trimmed 0.85, PointToPlane, Counter 40 +
 * Differential(1e-3, 1e-3, 3) -- the matcher/outlier/minimiser/checker part of
 * PM::ICPSequence::setDefault() (Mapper.cpp:77). */
void b200icp_config_default(b200icp_config* cfg, int32_t dim);

/* Construct the ICP object (replaces the `PM::ICPSequence icp` member, Mapper.h:23, configured at
 * Mapper.cpp:72/77). */
int32_t b200icp_create(const b200icp_config* cfg, int32_t device, b200icp_ctx** out);
void b200icp_destroy(b200icp_ctx* ctx);
const char* b200icp_last_error(const b200icp_ctx* ctx); /* ctx may be NULL: creation errors */
void* b200icp_stream(b200icp_ctx* ctx);                 /* the cudaStream_t all work is issued on */
int32_t b200icp_set_profiling(b200icp_ctx* ctx, int32_t on); /* per-kernel event timing (adds syncs) */
int32_t b200icp_get_timing(const b200icp_ctx* ctx, b200icp_timing* out);
/* How many SMs the registration loop of this context may occupy (it is a persistent kernel with one CTA per SM); 0 = all
 * (default).  Contexts that work side by side on one GPU -- b200icp_register_batch does this by itself -- or a registration
 * that should leave room for a map update running on another context (Mapper isOnline, Mapper.cpp:280-283) take a share. */
int32_t b200icp_set_sm_share(b200icp_ctx* ctx, int32_t n_sms);

/* icp.setMap(cloud) -- Map.cpp:111,178,528,581.  Copies the cloud, mean-centres it
 * (T_refIn_refMean) and builds the spatial index that replaces libnabo's kd-tree.
 * normals may be NULL unless the minimiser is point-to-plane (then B200ICP_ERR_INVALID_FIELD is
 * reported by register, as LPM does).  n == 0 is ignored and returns B200ICP_ERR_NO_MAP state
 * unchanged (LPM: "Ignoring attempt to create a map from an empty cloud", returns false). */
int32_t b200icp_set_map(b200icp_ctx* ctx, const float* features, int32_t feature_rows,
                        const float* normals, int64_t n);
int32_t b200icp_set_map_device(b200icp_ctx* ctx, const float* d_features, int32_t feature_rows,
                               const float* d_normals, int64_t n);
int64_t b200icp_map_size(const b200icp_ctx* ctx);

/* correction = icp(input) + errorMinimizer->getOverlap() -- Mapper.cpp:213,219.
 * `reading` is the scan already moved by the estimated pose (Mapper.cpp:197).  T_init may be NULL
 * (identity: the overload the reference calls).  T_out receives the (dim+1)x(dim+1) correction. */
int32_t b200icp_register(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq,
                         const float* T_init, float* T_out, b200icp_result* result);
int32_t b200icp_register_device(b200icp_ctx* ctx, const float* d_reading, int32_t feature_rows,
                                int64_t nq, const float* T_init, float* T_out,
                                b200icp_result* result);

/* icp(input) for a reading that carries the `normals` descriptor (dim x nq, column-major; e.g. from a
 * SurfaceNormalDataPointsFilter in the `input:` chain): needed by SurfaceNormalOutlierFilter only.  reading_normals may be NULL
 * (then identical to b200icp_register).  Host pointers. */
int32_t b200icp_register_normals(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq,
                                 const float* reading_normals, const float* T_init, float* T_out, b200icp_result* result);

/* icp(input) for a reading that carries descriptors the ICP chain reads: `normals` (see above) and / or `maxSearchDist` (1 x nq):
 * libpointmatcher's KDTreeMatcher searches every reading point within its own radius when the reading has that descriptor, and
 * the matcher's maxDist is then not used (libnabo's knn overload with a vector of radii).  Either pointer may be NULL. */
int32_t b200icp_register_descriptors(b200icp_ctx* ctx, const float* reading, int32_t feature_rows, int64_t nq, const float* reading_normals,
                                     const float* reading_max_search_dist, const float* T_init, float* T_out, b200icp_result* result);

/* ---- batched registration (BASELINE.json config 5; no reference counterpart: the reference aligns one scan at a time) ----
 * Independent scan <-> submap alignments.  One pair = `icp.setMap(submap)` (Map.cpp:111) followed by `icp(reading)`
 * (Mapper.cpp:213) on the same ICP object; pairs do not interact. */
typedef struct b200icp_pair {
    const float* map_features; /* (dim+1) x n_map, column-major, host.  NULL: keep the map the context already holds
                                  (the previous pair's submap, or one installed by b200icp_set_map) -- scans against one map */
    const float* map_normals;  /* dim x n_map or NULL (then point-to-plane fails with B200ICP_ERR_INVALID_FIELD)          */
    int64_t n_map;
    const float* reading;      /* (dim+1) x n_reading, column-major, host                                                   */
    int64_t n_reading;
    const float* T_init;       /* (dim+1) x (dim+1) or NULL = identity                                                      */
} b200icp_pair;

typedef struct b200icp_pair_result {
    float T[16];              /* the (dim+1) x (dim+1) correction, column-major, in the first (dim+1)^2 entries           */
    b200icp_result result;
    int32_t status;           /* b200icp_status of this pair (a failed pair does not stop the batch)                      */
    float setmap_ms, register_ms; /* device time of the two steps (CUDA events on the context's stream)                   */
} b200icp_pair_result;

/* Registers pairs[0 .. n_pairs): pair j runs on ctxs[j % n_ctx], each context working through its share in order on its
 * own host thread and CUDA stream.  Contexts may live on different devices (one process driving several GPUs) and / or
 * share a device: two contexts on one GPU overlap pair j+1's submap upload (copy engine) with pair j's ICP loop (SMs).
 * All contexts must have the same `dim`.  Returns B200ICP_OK when every pair succeeded, else the status of the first
 * failed pair (per-pair statuses in out[]).  Blocks until the whole batch is done.  The contexts must not be used by
 * other threads meanwhile. */
int32_t b200icp_register_batch(b200icp_ctx* const* ctxs, int32_t n_ctx, const b200icp_pair* pairs, int64_t n_pairs,
                               b200icp_pair_result* out);

/* matcher->findClosests(cloud) against the current map -- the KDTreeMatcher step of the loop and
 * the direct Nabo::NNS::knn call sites (PointDistanceMapperModule.cpp:36).  `queries` are in the
 * map frame; ids index the cloud given to set_map; dists are squared; missing = -1 / +inf.
 * Output is knn x nq, column-major (ids[i*knn + j] = j-th neighbour of query i). */
int32_t b200icp_match(b200icp_ctx* ctx, const float* queries, int32_t feature_rows, int64_t nq,
                      int32_t* ids, float* dists2);

/* Nabo::NNS::create(ref) + knn(query, k, eps=0, ALLOW_SELF_MATCH, maxRadius) on arbitrary clouds
 * (PointDistanceMapperModule.cpp:33-36, DynamicPointsMapperModule.cpp:75-78, and the self-kNN of
 * SurfaceNormalDataPointsFilter).  No centring.  Host pointers. */
int32_t b200icp_knn(b200icp_ctx* ctx, const float* ref, int32_t ref_rows, int64_t nref,
                    const float* queries, int32_t query_rows, int64_t nq, int32_t dim, int32_t k,
                    float max_radius, int32_t* ids, float* dists2);

/* PM::Transformation("RigidTransformation")->compute(cloud, T) -- Mapper.cpp:197,221;
 * Map.cpp:523,525.  In place on host buffers; normals may be NULL.  Fails with
 * B200ICP_ERR_TRANSFORM when |1 - det R| > 1e-3. */
int32_t b200icp_transform(b200icp_ctx* ctx, float* features, int32_t feature_rows, float* normals,
                          int64_t n, const float* T);

/* ---- the device-resident local map (Map::localPointCloud, Map.h:38-46) ---------------------------
 * b200icp_set_map* also (re)initialises this store: every point of the cloud, map frame, all
 * "loaded".  The calls below are the steps of Map::updateLocalPointCloud (Map.cpp:502-534) and of
 * the cell window of Map::updatePose (Map.cpp:246-460) on that store; the spatial index is rebuilt
 * from the loaded points by b200icp_map_commit (== icp.setMap(localPointCloud), Map.cpp:111,178,528). */

/* PointDistanceMapperModule::inPlaceUpdateMap (MapperModules/PointDistanceMapperModule.cpp:28-50):
 * 1-NN (eps 0, no radius) of every input point against the local map -- on the live index, no
 * second kd-tree -- keep those with dist2 >= minDistNewPoint^2, append them in input order
 * (map.concatenate: the map's normals survive only if input_normals is given).  On an empty local
 * map this is createMap (:9-19): the whole input is taken.  `input` is in the map frame (host
 * pointers).  keep_out (optional): n_in bytes, 1 = appended. */
int32_t b200icp_map_insert_point_distance(b200icp_ctx* ctx, const float* input, int32_t feature_rows,
                                          int64_t n_in, const float* input_normals,
                                          float min_dist_new_point, int64_t* n_added, uint8_t* keep_out);

/* SurfaceNormalDataPointsFilter{knn} over the whole local map (the `post:` chain of
 * examples/config.yaml:26-27 applied at Map.cpp:523-525): self k-NN, centred covariance,
 * eigenvector of the smallest eigenvalue -> `normals` (unit, sign arbitrary).  Computed in the map
 * frame: k-NN sets and normals are invariant/covariant under the rigid sensor<->map transform the
 * reference wraps around the filter, so the two whole-map transforms are not needed. */
int32_t b200icp_map_surface_normals(b200icp_ctx* ctx, int32_t knn);

/* The same filter on an arbitrary host cloud -- a SurfaceNormalDataPointsFilter in the `input:` chain (Mapper.cpp:187-191), which
 * configurations that use SurfaceNormalOutlierFilter need on the reading.  normals_out: dim x n, column-major. */
int32_t b200icp_cloud_surface_normals(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, int32_t knn,
                                      float* normals_out);

/* Map::loadCells (load = 1, Map.cpp:71-128) / Map::unloadCells (load = 0, Map.cpp:140-230) for the
 * slab {startRow, endRow, startColumn, endColumn, startAisle, endAisle} of 20 m cells.  Points never
 * leave HBM: unloading clears their `loaded` flag (the reference moves them to the CellManager),
 * loading sets it again.  Does not rebuild the index (call b200icp_map_commit). */
int32_t b200icp_map_window(b200icp_ctx* ctx, int32_t load, const int32_t* slab6, int64_t* n_changed);

/* Pre-size the device buffers for a map of n_points: the point store and its compaction scratch, the index (points,
 * normals, cell table, sort scratch), the auxiliary index the update steps build over the changed points and, if
 * normals_knn > 0, the self-k-NN scratch and bookkeeping of b200icp_map_surface_normals.  An online map then grows
 * without cudaMalloc / cudaFree on the update path (a cudaFree next to gigabytes of live buffers was measured at
 * 70-870 ms; B200ICP_TRACE_ALLOC=1 reports every allocation and release on stderr).  Optional. */
int32_t b200icp_map_reserve(b200icp_ctx* ctx, int64_t n_points, int32_t normals_knn);

/* icp.setMap(localPointCloud): rebuild the index over the loaded points.  An empty local cloud is
 * ignored like LPM does (the previous index stays). */
int32_t b200icp_map_commit(b200icp_ctx* ctx);

/* ---- online mapping (Mapper isOnline: Mapper.cpp:225-228,248-255,280-283; Map.cpp:29-57,482-494) ----
 * The reference hands Map::updateLocalPointCloud to a std::async worker and keeps registering scans against the OLD map until
 * the worker's final icp.setMap(localPointCloud) (Map.cpp:527-529).  Here: every entry point of a context is atomic with
 * respect to the others (any host thread may call; a registration waits for at most one update STEP, never for the update),
 * and between begin_update and end_update the update steps (module inserts, b200icp_map_commit, the post filters) build a second
 * index from the store while b200icp_register* / b200icp_match keep using the live one; end_update publishes it.
 *   b200icp_scan_snapshot   caller thread, before dispatching the worker: the scan in the slot becomes the worker's input
 *                           (the reference copies `currentInput` into the async call); the caller's slot is empty afterwards
 *   b200icp_map_begin_update   worker thread: from here on this thread's b200icp_scan_* calls see the snapshot
 *   b200icp_map_end_update     worker thread, under the caller's icpMapLock: swap the indexes
 * Without begin_update nothing changes: b200icp_map_commit rebuilds the live index in place. */
int32_t b200icp_scan_snapshot(b200icp_ctx* ctx);
int32_t b200icp_map_begin_update(b200icp_ctx* ctx);
int32_t b200icp_map_end_update(b200icp_ctx* ctx);
int32_t b200icp_map_update_in_progress(const b200icp_ctx* ctx);

/* Point counts: local (Map::localPointCloud) and global (local + parked cells, Map.cpp:538-562). */
int32_t b200icp_map_counts(const b200icp_ctx* ctx, int64_t* n_local, int64_t* n_global);
int32_t b200icp_map_has_normals(const b200icp_ctx* ctx);

/* Map::getLocalPointCloud (global = 0) / getGlobalPointCloud (global = 1): copy to host buffers of
 * `capacity` points ((dim+1) x N features, dim x N normals or NULL).  With features == NULL only
 * *n_out is written. */
int32_t b200icp_map_download(b200icp_ctx* ctx, int32_t global, float* features, float* normals,
                             int64_t capacity, int64_t* n_out);

/* `probabilityDynamic` descriptor of the map (AddDescriptorDataPointsFilter on a map given by set_map):
 * per-point values (prob != NULL, one per map point in insertion order) or a constant. */
int32_t b200icp_map_set_prob(b200icp_ctx* ctx, const float* prob, float constant);
int32_t b200icp_map_has_prob(const b200icp_ctx* ctx);
int32_t b200icp_map_download_prob(b200icp_ctx* ctx, int32_t global, float* prob, int64_t capacity);

/* map.concatenate(input) / `DataPoints outputMap(input)` of the modules' createMap: append every input
 * point; a descriptor survives only if both clouds carry it (an empty map takes the input's). */
int32_t b200icp_map_append(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                           const float* input_normals, const float* input_prob, int64_t* n_added);

/* OctreeMapperModule::inPlaceUpdateMap (MapperModules/OctreeMapperModule.cpp:35-39): map.concatenate(input)
 * then libpointmatcher's OctreeGridDataPointsFilter{maxPointByNode, maxSizeByNode, samplingMethod}
 * over the whole local map.  The octree descent is emulated bit for bit (child = p > centre per axis,
 * centre +- r/2, until 2r <= maxSizeByNode); one survivor per occupied leaf: samplingMethod 0 = the
 * first point (lowest index), 1 = a uniformly random member (counter-based generator keyed by the leaf:
 * reproducible, base seed from B200ICP_OCTREE_SEED; upstream draws from its own unseeded generator, so only the
 * distribution can agree), 2 = centroid of features and descriptors, 3 = medoid (the member closest to the
 * leaf's centroid).  maxPointByNode must be 1 (the value OctreeMapperModule.cpp hard-codes).
 * input_normals / input_prob may be NULL (concatenate then drops that descriptor from the map). */
int32_t b200icp_map_octree(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                           const float* input_normals, const float* input_prob, float max_size_by_node,
                           int32_t max_point_by_node, int32_t sampling_method, int64_t* n_after);

/* CutAtDescriptorThresholdDataPointsFilter{descName probabilityDynamic, useLargerThan, threshold}
 * (examples/config.yaml:29-32) over the local map. */
int32_t b200icp_map_cut_at_threshold(b200icp_ctx* ctx, float threshold, int32_t use_larger_than, int64_t* n_removed);

/* DynamicPointsMapperModule parameters (MapperModules/DynamicPointsMapperModule.h:33-44). */
typedef struct b200icp_dynamic_params {
    float threshold_dynamic, alpha, beta, beam_half_angle, epsilon_a, epsilon_d, sensor_max_range;
} b200icp_dynamic_params;

/* DynamicPointsMapperModule::inPlaceUpdateMap (MapperModules/DynamicPointsMapperModule.cpp:34-151):
 * scan and map to the sensor frame (pose^-1), spherical coordinates, 1-NN of every map point within
 * sensorMaxRange against the scan in (elevation, azimuth) space with radius 2*beamHalfAngle -- on a
 * 2-D grid index over the scan angles, no kd-tree -- then the Bayesian update of the map's
 * probabilityDynamic.  Needs map normals and probabilityDynamic and an input probabilityDynamic
 * (B200ICP_ERR_INVALID_FIELD otherwise, like the reference's InvalidField). */
int32_t b200icp_map_dynamic_points(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                                   const float* input_prob, const float* pose, const b200icp_dynamic_params* prm);

/* DataPointsFilters::apply(cloud) for the `input:` chain (Mapper::applyInputFilters, Mapper.cpp:187-191):
 * one predicate kernel evaluating the whole chain per point + ordered stream compaction on the device.
 * In place on the host buffer; *n is updated to the surviving count. */
int32_t b200icp_filter_cloud(b200icp_ctx* ctx, float* features, int32_t feature_rows, int64_t* n,
                             const b200icp_filter* chain, int32_t n_filters);

/* The map's other descriptors: everything a libpointmatcher cloud carries besides `normals` and `probabilityDynamic` (the bundled
 * scans have intensity, t, ring ...; PointDistanceMapperModule.cpp:49 / OctreeMapperModule.cpp:37 concatenate them into the map).
 * The device keeps them as one point-major block of extra_rows floats per point == the reference's column-major extra_rows x N
 * matrix; their names and row ranges live with the caller (host/DataPoints.h), which also applies DataPoints::concatenate's
 * rule -- only descriptors present in BOTH clouds survive -- by selecting the surviving rows on both sides before an insert
 * (b200icp_map_select_extra / b200icp_scan_select_extra).  set_extra
 * attaches the block to the map given by the last b200icp_set_map (insertion order, host pointer). */
#define B200ICP_MAX_EXTRA_ROWS 32
int32_t b200icp_map_set_extra(b200icp_ctx* ctx, const float* extra, int32_t extra_rows);
int32_t b200icp_map_extra_rows(const b200icp_ctx* ctx);
int32_t b200icp_map_select_extra(b200icp_ctx* ctx, const int32_t* rows, int32_t n_rows);
int32_t b200icp_map_download_extra(b200icp_ctx* ctx, int32_t global, float* extra, int64_t capacity);

/* `localPointCloud = cloud`: what a MapperModule written against the reference's host signature (MapperModules/MapperModule.h:20-29,
 * inPlaceUpdateMap(const DataPoints& input, DataPoints& map, pose)) leaves in `map` replaces the loaded points of the device map;
 * parked cells stay.  Host pointers; normals / prob / extra may be NULL.  Does not rebuild the index (b200icp_map_commit). */
int32_t b200icp_map_replace_local(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                                  const float* prob, const float* extra, int32_t extra_rows);

/* ---- spill tier under the device grid: the CellManager seam (CellManager.h:15-18; RAMCellManager.cpp, HardDriveCellManager.cpp) ----
 * By default the cells the window leaves stay in HBM with their `loaded` flag cleared.  A Map configured with a CellManager
 * (host RAM or disk, host/CellManager.h) moves them out instead: after b200icp_map_window(load = 0) it calls
 *   b200icp_map_evict_parked   copy the parked points (every point with loaded = 0, insertion order) to host buffers and remove
 *                              them from the device store; with features == NULL only *n_out (their count) is written
 * groups them by 20 m cell id "row_col_aisle" and hands them to CellManager::saveCell; when the window comes back it
 * concatenates CellManager::retrieveCell(id) into the local map with
 *   b200icp_map_append_cloud   map.concatenate(cloud) for a host cloud with all its descriptors (loaded = 1). */
int32_t b200icp_map_evict_parked(b200icp_ctx* ctx, float* features, float* normals, float* prob, float* extra, int64_t capacity,
                                 int64_t* n_out);
int32_t b200icp_map_append_cloud(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                                 const float* prob, const float* extra, int32_t extra_rows, int64_t* n_added);

/* b200icp_map_insert_point_distance for an input that carries `probabilityDynamic` (the chain DynamicPointsMapperModule +
 * PointDistanceMapperModule: map.concatenate keeps the descriptor because both clouds have it). */
int32_t b200icp_map_insert_point_distance_prob(b200icp_ctx* ctx, const float* input, int32_t feature_rows, int64_t n_in,
                                               const float* input_normals, const float* input_prob, float min_dist_new_point,
                                               int64_t* n_added, uint8_t* keep_out);

/* ---- the device-resident scan slot (no reference counterpart: it removes the host round trips between the steps of
 * Mapper::applyInputFilters + Mapper::processInput, Mapper.cpp:187-238).  The raw scan is uploaded ONCE, with its descriptors;
 * the `input:` filter chain (Mapper.cpp:187-191), RigidTransformation (Mapper.cpp:197,221), icp(input) (:213) and the
 * MapperModules' inPlaceUpdateMap (via Map.cpp:502-534) then run on that copy, stream-ordered.
 *   upload            features only; clears the descriptors of the previous scan
 *   set_descriptors   normals (dim x n), probabilityDynamic (1 x n), the other descriptors (extra_rows x n); any may be NULL.
 *                     rotating_rows lists the first rows of the dim-row descriptors inside `extra` that rotate with the cloud
 *                     (libpointmatcher rotates `normals` and `observationDirections`)
 *   filter            DataPointsFilters::apply of the BoundingBox / DistanceLimit / RandomSampling chain, descriptors follow
 *   add_prob          AddDescriptorDataPointsFilter{probabilityDynamic, 1, [constant]} (examples/config.yaml:19-23)
 *   surface_normals   SurfaceNormalDataPointsFilter{knn} on the scan
 *   select_extra      keep the listed rows of `extra`, in that order
 *   transform         RigidTransformation::compute in place (features; normals and rotating descriptors by R)
 *   register          icp(scan); the scan's normals, if any, go to SurfaceNormalOutlierFilter
 *   insert_point_distance / append / octree / dynamic_points
 *                     PointDistanceMapperModule / `map.concatenate(scan)` / OctreeMapperModule / DynamicPointsMapperModule
 *                     ::inPlaceUpdateMap with the scan as `input` (map frame)
 *   download*         host copies (tests, modules without a device entry point) */
int32_t b200icp_scan_upload(b200icp_ctx* ctx, const float* features, int32_t feature_rows, int64_t n);
int32_t b200icp_scan_set_descriptors(b200icp_ctx* ctx, const float* normals, const float* prob, const float* extra, int32_t extra_rows,
                                     const int32_t* rotating_rows, int32_t n_rotating);
int64_t b200icp_scan_size(const b200icp_ctx* ctx);
int32_t b200icp_scan_info(const b200icp_ctx* ctx, int32_t* has_normals, int32_t* has_prob, int32_t* extra_rows);
int32_t b200icp_scan_filter(b200icp_ctx* ctx, const b200icp_filter* chain, int32_t n_filters, int64_t* n_out);
int32_t b200icp_scan_add_prob(b200icp_ctx* ctx, float constant);
int32_t b200icp_scan_surface_normals(b200icp_ctx* ctx, int32_t knn);
int32_t b200icp_scan_select_extra(b200icp_ctx* ctx, const int32_t* rows, int32_t n_rows);
int32_t b200icp_scan_transform(b200icp_ctx* ctx, const float* T);
int32_t b200icp_scan_register(b200icp_ctx* ctx, const float* T_init, float* T_out, b200icp_result* result);
int32_t b200icp_scan_insert_point_distance(b200icp_ctx* ctx, float min_dist_new_point, int64_t* n_added);
int32_t b200icp_scan_append(b200icp_ctx* ctx, int64_t* n_added);
int32_t b200icp_scan_octree(b200icp_ctx* ctx, float max_size_by_node, int32_t max_point_by_node, int32_t sampling_method, int64_t* n_after);
int32_t b200icp_scan_dynamic_points(b200icp_ctx* ctx, const float* pose, const b200icp_dynamic_params* prm);
int32_t b200icp_scan_download(b200icp_ctx* ctx, float* features, int64_t capacity, int64_t* n_out);
int32_t b200icp_scan_download_descriptors(b200icp_ctx* ctx, float* normals, float* prob, float* extra, int64_t capacity);

/* Device-pointer variant of b200icp_transform (features/normals live on ctx's device). */
int32_t b200icp_transform_device(b200icp_ctx* ctx, float* d_features, int32_t feature_rows,
                                 float* d_normals, int64_t n, const float* T);

/* Introspection (no reference counterpart; used by tests and the bench):
 *  - the translation of T_refIn_refMean chosen by the last set_map (ICPSequence::setMap's mean),
 *  - the cell edge and grid dimensions of the spatial index,
 *  - per-iteration T_iter of the last registration (refMean frame, (dim+1)^2 column-major each;
 *    enable with b200icp_set_trace before registering; get_trace returns the iteration count). */
int32_t b200icp_get_map_mean(const b200icp_ctx* ctx, float* mean3);
int32_t b200icp_get_grid_info(const b200icp_ctx* ctx, float* cell_edge, int32_t* dims3);
int32_t b200icp_set_trace(b200icp_ctx* ctx, int32_t on);
int32_t b200icp_get_trace(const b200icp_ctx* ctx, float* out, int32_t max_iterations);

#ifdef __cplusplus
}
#endif
#endif /* B200ICP_H */
