/*
 * icp_oracle.h -- CPU oracle for the ICP registration hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (libb200icp.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the arithmetic on this path lives in libpointmatcher (pinned >= 1.4.3 by
 * /root/reference/CMakeLists.txt:33) and libnabo (package.xml:14), neither of which is vendored
 * under /root/reference nor installed here, and the reference ships no tests or golden vectors
 * (SURVEY.md section 4 and 8c).  Every function below restates the published upstream algorithm and
 * names the reference call site that reaches it; the restatement is cross-checked against
 * scipy.spatial.cKDTree, numpy.linalg and closed-form transforms in tests/, not against a reference
 * binary.
 */
#ifndef ICP_ORACLE_H
#define ICP_ORACLE_H

#include <stdint.h>
#include "b200icp.h" /* shares the flattened `icp:` config struct and result struct with the product */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- libnabo NearestNeighbourSearch (KDTREE_LINEAR_HEAP, bucket 8) ------------------------- */
typedef struct orc_kdtree orc_kdtree;

/* Nabo::NNS::create(cloud, dim, KDTREE_LINEAR_HEAP) -- PointDistanceMapperModule.cpp:33-34,
 * DynamicPointsMapperModule.cpp:75-76, and KDTreeMatcher::init under icp.setMap (Map.cpp:111).
 * pts: n points, `stride` floats apart, first `dim` floats of each are the coordinates. */
orc_kdtree* orc_kdtree_build(const float* pts, int32_t stride, int64_t n, int32_t dim);
void orc_kdtree_free(orc_kdtree* t);
/* NNS::knn(query, ids, dists2, k, epsilon=0, ALLOW_SELF_MATCH, maxRadius) --
 * PointDistanceMapperModule.cpp:36.  ids/d2 are k x nq column-major; missing = -1 / +inf.
 * nthreads: OpenMP threads over queries (libnabo's own parallelisation); <=0 = all. */
void orc_kdtree_knn(const orc_kdtree* t, const float* q, int32_t qstride, int64_t nq, int32_t k,
                    float max_radius, int32_t* ids, float* d2, int32_t nthreads);

/* ---- PM::ICPSequence ----------------------------------------------------------------------- */
typedef struct orc_icp orc_icp;

void orc_kdtree_knn_ex(const orc_kdtree* t, const float* q, int32_t qstride, int64_t nq, int32_t k, float max_radius,
                       const float* max_radii, int32_t strict, int32_t* ids, float* d2, int32_t nthreads);

orc_icp* orc_icp_create(const b200icp_config* cfg);
void orc_icp_destroy(orc_icp* o);
const char* orc_icp_last_error(const orc_icp* o);
float orc_icp_last_var_ratio(const orc_icp* o); /* VarTrimmed: tuned ratio of the last iteration */
float orc_icp_last_robust_scale(const orc_icp* o); /* RobustOutlierFilter: scale used by the last iteration */
/* `normals` descriptor of the reading for the next orc_icp_register (dim x n column-major; NULL clears): SurfaceNormalOutlierFilter */
int32_t orc_icp_set_reading_normals(orc_icp* o, const float* normals, int64_t n);
int32_t orc_icp_set_reading_max_search_dist(orc_icp* o, const float* radii, int64_t n);
/* icp.setMap(cloud) -- Map.cpp:111,178,528,581. Returns b200icp_status. */
int32_t orc_icp_set_map(orc_icp* o, const float* features, int32_t rows, const float* normals,
                        int64_t n);
/* Mean used by setMap (T_refIn_refMean translation), for tests. */
void orc_icp_get_mean(const orc_icp* o, float mean[3]);
/* correction = icp(input) -- Mapper.cpp:213; result->overlap = getOverlap() -- Mapper.cpp:219.
 * trace (optional): room for max_iteration_count (dim+1)^2 matrices, receives T_iter after each
 * iteration (refMean frame).  secs (optional, 4 doubles): matching, outlier, minimise, total. */
int32_t orc_icp_register(orc_icp* o, const float* reading, int32_t rows, int64_t nq,
                         const float* T_init, float* T_out, b200icp_result* result, float* trace,
                         double* secs, int32_t nthreads);
/* matcher->findClosests(cloud) against the centred map; cloud given in the map (refIn) frame. */
int32_t orc_icp_match(orc_icp* o, const float* queries, int32_t rows, int64_t nq, int32_t* ids,
                      float* d2, int32_t nthreads);

/* PM::Transformation("RigidTransformation")->compute -- Mapper.cpp:197,221; Map.cpp:523,525. */
int32_t orc_transform(float* features, int32_t rows, float* normals, int64_t n, const float* T);

/* PointDistanceMapperModule::inPlaceUpdateMap -- MapperModules/PointDistanceMapperModule.cpp:28-50.
 * keep[i] = 1 iff input point i is appended to the map.  Returns the number kept. */
int64_t orc_point_distance_keep(const float* map_feat, int32_t rows, int64_t n_map,
                                const float* input_feat, int64_t n_in, float min_dist_new_point,
                                uint8_t* keep, int32_t nthreads);

/* SurfaceNormalDataPointsFilter{knn} -- examples/config.yaml:26-27 via Map.cpp:524.
 * normals: dim x n column-major (sign arbitrary, as upstream). */
int32_t orc_surface_normals(const float* feat, int32_t rows, int64_t n, int32_t knn,
                            float* normals, int32_t nthreads);

int32_t orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
