/*
 * b200mapper.h -- C face of the host-side mirror of norlab_icp_mapper::Mapper
 * (norlab_icp_mapper_b200/host/, libb200mapper.so).  It exists so that non-C++ callers (the ctypes
 * tests, a cgo/JNI binding) can drive the same Mapper::processInput sequence the reference exposes
 * through pybind11 (python/src/mapper.cpp:11-25).  Matrices are column-major fp32 like b200icp.h.
 */
#ifndef B200MAPPER_H
#define B200MAPPER_H

#include <stdint.h>
#include "b200icp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200mapper b200mapper;

typedef struct b200mapper_config {
    b200icp_config icp;        /* YAML `icp:`                                                        */
    int32_t update_condition;  /* mapper.updateCondition.type: 0 distance, 1 delay, 2 overlap        */
    float update_value;        /* mapper.updateCondition.value                                       */
    float sensor_max_range;    /* mapper.sensorMaxRange (default 200)                                */
    float min_dist_new_point;  /* PointDistanceMapperModule{minDistNewPoint}; < 0: the default 0.15  */
    int32_t surface_normal_knn;/* post: SurfaceNormalDataPointsFilter{knn}; 0 = absent               */
    int32_t is_3d, is_online, is_mapping;
    /* mapper.mapperModule list, applied in this order when enabled (examples/config.yaml:38-52):
     * DynamicPointsMapperModule, then OctreeMapperModule or PointDistanceMapperModule */
    int32_t use_dynamic_points;
    b200icp_dynamic_params dynamic_points;
    int32_t use_octree;            /* replaces the PointDistance module */
    float octree_max_size_by_node;
    int32_t octree_sampling_method;
    /* post: CutAtDescriptorThresholdDataPointsFilter{probabilityDynamic} after SurfaceNormal */
    int32_t use_cut_at_threshold;
    float cut_threshold;
    /* input: chain (after Mapper's own radius filter) + AddDescriptor{probabilityDynamic} */
    int32_t n_input_filters;
    b200icp_filter input_filters[6];
    int32_t add_probability_dynamic;
    float probability_dynamic_value;
    int32_t reserve_points;    /* > 0: pre-size the device map for this many points (no reference counterpart) */
    int32_t input_surface_normal_knn; /* input: SurfaceNormalDataPointsFilter{knn} on the reading (for SurfaceNormalOutlierFilter); 0 = absent */
    int32_t cell_spill;        /* where the cells the window leaves go: 0 stay in device memory (flag only; default), 1 RAMCellManager
                                  (host memory), 2 HardDriveCellManager (the constructor's saveMapCellsOnHardDrive = true)        */
    int32_t reserved[1];
    char cell_folder[128];     /* HardDriveCellManager: folder of cell_<id>.vtk; empty = "/tmp/" like the reference                */
} b200mapper_config;

typedef struct b200mapper_stats {
    float overlap;             /* icp.errorMinimizer->getOverlap() of the last processInput          */
    int32_t iterations;
    int32_t map_updated;       /* the last processInput called updateMap                             */
    int32_t n_window_updates;  /* slabs loaded/unloaded by the last Map::updatePose                  */
    int64_t n_local, n_global; /* points in Map::localPointCloud / in the whole map                  */
} b200mapper_stats;

/* Mapper::Mapper(config, is3D, isOnline, isMapping, saveMapCellsOnHardDrive) -- Mapper.cpp:15-33 */
int32_t b200mapper_create(const b200mapper_config* cfg, int32_t device, b200mapper** out);
/* The reference's constructor signature: Mapper(configFilePath, is3D, isOnline, isMapping, saveMapCellsOnHardDrive) with the YAML
 * file of Mapper::loadYamlConfig (Mapper.cpp:59-185; /root/reference/examples/config.yaml loads unmodified).  Unknown filter / module
 * / parameter names fail with B200ICP_ERR_INVALID_ARG and the message libpointmatcher's registrar would give. */
int32_t b200mapper_create_from_yaml(const char* config_file_path, int32_t is_3d, int32_t is_online, int32_t is_mapping,
                                    int32_t save_map_cells_on_hard_drive, int32_t device, int32_t reserve_points, b200mapper** out);
/* What host/YamlConfig.h reads from a configuration file, as `key=value` lines (tests; needs no GPU). */
int32_t b200mapper_yaml_summary(const char* config_file_path, int32_t is_3d, char* out, int32_t capacity);
void b200mapper_destroy(b200mapper* m);
const char* b200mapper_last_error(const b200mapper* m);
/* Mapper::applyInputFilters -- Mapper.cpp:187-191; in place, returns the new point count in *n */
int32_t b200mapper_apply_input_filters(b200mapper* m, float* features, int32_t feature_rows, int64_t* n);
/* Mapper::processInput -- Mapper.cpp:194-238; status codes of b200icp.h (exceptions of the C++ class) */
int32_t b200mapper_process_input(b200mapper* m, const float* features_sensor_frame, int32_t feature_rows, int64_t n,
                                 const float* estimated_pose, double time_stamp_seconds);
/* applyInputFilters + processInput on a RAW scan with one host-to-device copy: the filter chain, the descriptors it attaches, the
 * transforms, icp(input) and the map update all work on the device-resident copy (Mapper::setDeviceResidentInput).  *n_filtered
 * (optional) receives the point count after the `input:` chain. */
int32_t b200mapper_process_raw_input(b200mapper* m, const float* features_sensor_frame, int32_t feature_rows, int64_t n,
                                     const float* estimated_pose, double time_stamp_seconds, int64_t* n_filtered);
/* isOnline (Mapper.cpp:280-283): processInput hands the map update to a worker and returns; these two expose the future the
 * reference keeps in `mapUpdateFuture`: is an update still running / block until it (and the queued cell-window updates) are done,
 * reporting the update's error if it failed. */
int32_t b200mapper_map_update_in_flight(b200mapper* m);
int32_t b200mapper_wait_for_map_update(b200mapper* m);
int32_t b200mapper_get_pose(b200mapper* m, float* pose);                       /* Mapper::getPose        */
int32_t b200mapper_get_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n); /* getMap */
int32_t b200mapper_get_local_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n); /* Map::getLocalPointCloud */
int32_t b200mapper_get_new_local_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n,
                                     int32_t* available);                      /* Mapper::getNewLocalMap */
int32_t b200mapper_set_map(b200mapper* m, const float* features, int32_t feature_rows, const float* normals, int64_t n);
/* Mapper::setMap for a map that carries probabilityDynamic (a map saved by a DynamicPointsMapperModule run): prob may be NULL */
int32_t b200mapper_set_map_descriptors(b200mapper* m, const float* features, int32_t feature_rows, const float* normals, const float* prob, int64_t n);
/* the probabilityDynamic descriptor of Mapper::getMap(); *n = 0 when the map does not carry it */
int32_t b200mapper_get_map_prob(b200mapper* m, float* prob, int64_t capacity, int64_t* n);
int32_t b200mapper_get_is_mapping(const b200mapper* m);
int32_t b200mapper_set_is_mapping(b200mapper* m, int32_t is_mapping);
int64_t b200mapper_trajectory_size(b200mapper* m);                             /* Mapper::getTrajectory  */
int32_t b200mapper_get_trajectory(b200mapper* m, float* poses, double* stamps, int64_t capacity);
int32_t b200mapper_get_stats(b200mapper* m, b200mapper_stats* out);
/* slabs applied by the last Map::updatePose, 7 ints each (see Map::lastUpdates) */
int32_t b200mapper_get_window_updates(b200mapper* m, int32_t* out7, int32_t capacity);

/* DataPoints::save / DataPoints::load for legacy VTK POLYDATA (host/IO.h) -- the disk format of maps, scans and map cells
 * (examples/build_map_from_scans_and_trajectory.cpp:50,94; HardDriveCellManager.cpp:14-27).  No GPU involved; errors go to
 * b200mapper_last_error(NULL).  save: normals / prob may be NULL.  load: call with features == NULL for *n and the flags, then
 * with buffers of that capacity (normals, prob: NULL to skip; rows of NaN are never written -- check the flags). */
int32_t b200mapper_vtk_save(const char* path, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                            const float* prob, int32_t binary);
int32_t b200mapper_vtk_load(const char* path, int32_t dim, float* features, float* normals, float* prob, int64_t capacity, int64_t* n,
                            int32_t* has_normals, int32_t* has_prob);

#ifdef __cplusplus
}
#endif
#endif
