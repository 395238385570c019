// capi.cpp -- extern "C" face of the host mirror (include/b200mapper.h).
#include <cstring>
#include <string>

#include <sstream>

#include "IO.h"
#include "Mapper.h"
#include "YamlConfig.h"
#include "b200mapper.h"

using namespace norlab_icp_mapper_b200;

struct b200mapper {
    std::unique_ptr<Mapper> mapper;
    int dim = 3;
    std::string err;
};

namespace {
std::string g_err;

template <typename F>
int32_t guarded(b200mapper* m, F&& f) {
    try {
        f();
        return B200ICP_OK;
    } catch (const ConvergenceError& e) {
        (m ? m->err : g_err) = e.what();
        return B200ICP_ERR_CONVERGENCE;
    } catch (const TransformationError& e) {
        (m ? m->err : g_err) = e.what();
        return B200ICP_ERR_TRANSFORM;
    } catch (const InvalidField& e) {
        (m ? m->err : g_err) = e.what();
        return B200ICP_ERR_INVALID_FIELD;
    } catch (const InvalidParameter& e) {
        (m ? m->err : g_err) = e.what();
        return B200ICP_ERR_INVALID_ARG;
    } catch (const std::exception& e) {
        (m ? m->err : g_err) = e.what();
        return B200ICP_ERR_CUDA;
    }
}

DataPoints wrap(const float* f, int rows, int64_t n, const float* nrm) {
    DataPoints d;
    d.dim = rows - 1;
    d.features.assign(f, f + (size_t)n * rows);
    if (nrm) d.normals.assign(nrm, nrm + (size_t)n * d.dim);
    return d;
}
TransformationParameters wrapT(const float* T, int n) {
    TransformationParameters P = TransformationParameters::Identity(n);
    if (T) std::memcpy(P.m, T, sizeof(float) * (size_t)(n * n));
    return P;
}
int32_t copy_out(const DataPoints& d, float* features, float* normals, int64_t capacity, int64_t* n) {
    *n = d.getNbPoints();
    if (!features) return B200ICP_OK;
    if (capacity < *n) return B200ICP_ERR_INVALID_ARG;
    std::memcpy(features, d.features.data(), d.features.size() * sizeof(float));
    if (normals && !d.normals.empty()) std::memcpy(normals, d.normals.data(), d.normals.size() * sizeof(float));
    return B200ICP_OK;
}
}  // namespace

extern "C" {

int32_t b200mapper_create(const b200mapper_config* cfg, int32_t device, b200mapper** out) {
    if (!cfg || !out) return B200ICP_ERR_INVALID_ARG;
    *out = nullptr;
    b200mapper* m = new b200mapper();
    const int32_t rc = guarded(nullptr, [&] {
        MapperConfig mc;
        mc.icp = cfg->icp;
        mc.post.surfaceNormalKnn = cfg->surface_normal_knn;
        mc.mapUpdateCondition = cfg->update_condition == 0 ? "distance" : (cfg->update_condition == 1 ? "delay" : (cfg->update_condition == 2 ? "overlap" : "?"));
        mc.mapUpdateValue = cfg->update_value;
        mc.sensorMaxRange = cfg->sensor_max_range;
        if (cfg->use_dynamic_points) {
            const b200icp_dynamic_params& d = cfg->dynamic_points;
            mc.mapperModules.push_back({"DynamicPointsMapperModule",
                                        Parameters{{"thresholdDynamic", std::to_string(d.threshold_dynamic)}, {"alpha", std::to_string(d.alpha)},
                                                   {"beta", std::to_string(d.beta)}, {"beamHalfAngle", std::to_string(d.beam_half_angle)},
                                                   {"epsilonA", std::to_string(d.epsilon_a)}, {"epsilonD", std::to_string(d.epsilon_d)},
                                                   {"sensorMaxRange", std::to_string(d.sensor_max_range)}}});
        }
        if (cfg->use_octree)
            mc.mapperModules.push_back({"OctreeMapperModule", Parameters{{"buildParallel", "1"}, {"maxSizeByNode", std::to_string(cfg->octree_max_size_by_node)},
                                                                         {"samplingMethod", std::to_string(cfg->octree_sampling_method)}}});
        else if (cfg->min_dist_new_point >= 0.f)
            mc.mapperModules.push_back({"PointDistanceMapperModule", Parameters{{"minDistNewPoint", std::to_string(cfg->min_dist_new_point)}}});
        mc.post.cutAtThreshold = cfg->use_cut_at_threshold != 0;
        mc.post.cutThreshold = cfg->cut_threshold;
        for (int i = 0; i < cfg->n_input_filters && i < 6; ++i) mc.inputFilters.push_back(cfg->input_filters[i]);
        mc.addProbabilityDynamic = cfg->add_probability_dynamic != 0;
        mc.probabilityDynamicValue = cfg->probability_dynamic_value;
        mc.inputSurfaceNormalKnn = cfg->input_surface_normal_knn;
        m->dim = cfg->is_3d ? 3 : 2;
        mc.spillCellsToHostRam = cfg->cell_spill == 1;
        if (cfg->cell_folder[0]) mc.cellFolder = cfg->cell_folder;
        m->mapper.reset(new Mapper(mc, cfg->is_3d != 0, cfg->is_online != 0, cfg->is_mapping != 0, /*saveMapCellsOnHardDrive=*/cfg->cell_spill == 2, device));
        if (cfg->reserve_points > 0)
            ICPSequence::check(m->mapper->getICP().context(), b200icp_map_reserve(m->mapper->getICP().context(), cfg->reserve_points, cfg->surface_normal_knn));
    });
    if (rc != B200ICP_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return B200ICP_OK;
}

int32_t b200mapper_create_from_yaml(const char* config_file_path, int32_t is_3d, int32_t is_online, int32_t is_mapping,
                                    int32_t save_map_cells_on_hard_drive, int32_t device, int32_t reserve_points, b200mapper** out) {
    if (!config_file_path || !out) return B200ICP_ERR_INVALID_ARG;
    *out = nullptr;
    b200mapper* m = new b200mapper();
    const int32_t rc = guarded(nullptr, [&] {
        m->dim = is_3d ? 3 : 2;
        m->mapper.reset(new Mapper(std::string(config_file_path), is_3d != 0, is_online != 0, is_mapping != 0, save_map_cells_on_hard_drive != 0, device));
        if (reserve_points > 0)
            ICPSequence::check(m->mapper->getICP().context(), b200icp_map_reserve(m->mapper->getICP().context(), reserve_points, 10));
    });
    if (rc != B200ICP_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return B200ICP_OK;
}

int32_t b200mapper_yaml_summary(const char* config_file_path, int32_t is_3d, char* out, int32_t capacity) {
    if (!config_file_path || !out || capacity < 1) return B200ICP_ERR_INVALID_ARG;
    out[0] = 0;
    return guarded(nullptr, [&] {
        const MapperConfig c = loadYamlConfig(std::string(config_file_path), is_3d != 0);
        std::ostringstream o;
        o << "icp.knn=" << c.icp.knn << "\nicp.maxDist=" << c.icp.max_dist << "\nicp.epsilon=" << c.icp.epsilon << "\nicp.minimizer=" << c.icp.minimizer
          << "\nicp.minimizerFlags=" << c.icp.minimizer_flags << "\nicp.nOutlier=" << c.icp.n_outlier;
        for (int i = 0; i < c.icp.n_outlier; ++i)
            o << "\nicp.outlier" << i << "=" << c.icp.outlier_kind[i] << "," << c.icp.outlier_param[i] << "," << c.icp.outlier_param2[i] << ","
              << c.icp.outlier_param3[i] << "," << c.icp.outlier_mode[i];
        o << "\nicp.counter=" << c.icp.max_iteration_count << "\nicp.differential=" << c.icp.use_differential << "," << c.icp.min_diff_rot_err << ","
          << c.icp.min_diff_trans_err << "," << c.icp.smooth_length << "\nicp.bound=" << c.icp.use_bound << "," << c.icp.max_rotation_norm << ","
          << c.icp.max_translation_norm << "\nicp.checkerOrder=" << c.icp.checker_order;
        o << "\ninput.n=" << c.inputFilters.size();
        for (size_t i = 0; i < c.inputFilters.size(); ++i) {
            const b200icp_filter& f = c.inputFilters[i];
            o << "\ninput" << i << "=" << f.kind << "," << f.lo[0] << "," << f.hi[0] << "," << f.lo[1] << "," << f.hi[1] << "," << f.lo[2] << "," << f.hi[2] << ","
              << f.dim << "," << f.dist << "," << f.remove_inside;
        }
        o << "\ninput.addProbabilityDynamic=" << (c.addProbabilityDynamic ? 1 : 0) << "," << c.probabilityDynamicValue
          << "\ninput.surfaceNormalKnn=" << c.inputSurfaceNormalKnn << "\npost.surfaceNormalKnn=" << c.post.surfaceNormalKnn
          << "\npost.cut=" << (c.post.cutAtThreshold ? 1 : 0) << "," << (c.post.cutUseLargerThan ? 1 : 0) << "," << c.post.cutThreshold
          << "\nmapper.updateCondition=" << c.mapUpdateCondition << "," << c.mapUpdateValue << "\nmapper.sensorMaxRange=" << c.sensorMaxRange
          << "\nmapper.nModules=" << c.mapperModules.size();
        for (size_t i = 0; i < c.mapperModules.size(); ++i) {
            o << "\nmodule" << i << "=" << c.mapperModules[i].first;
            for (const auto& kv : c.mapperModules[i].second) o << ";" << kv.first << "=" << kv.second;
        }
        const std::string t = o.str();
        if ((int)t.size() + 1 > capacity) throw InvalidParameter("summary buffer too small");
        std::memcpy(out, t.c_str(), t.size() + 1);
    });
}

void b200mapper_destroy(b200mapper* m) { delete m; }
const char* b200mapper_last_error(const b200mapper* m) { return m ? m->err.c_str() : g_err.c_str(); }

int32_t b200mapper_apply_input_filters(b200mapper* m, float* features, int32_t feature_rows, int64_t* n) {
    if (!m || !features || !n || feature_rows != m->dim + 1) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] {
        DataPoints d = wrap(features, feature_rows, *n, nullptr);
        m->mapper->applyInputFilters(d);
        *n = d.getNbPoints();
        std::memcpy(features, d.features.data(), d.features.size() * sizeof(float));
    });
}

int32_t b200mapper_process_input(b200mapper* m, const float* features, int32_t feature_rows, int64_t n, const float* estimated_pose,
                                 double time_stamp_seconds) {
    if (!m || (n > 0 && !features) || feature_rows != m->dim + 1) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] {
        DataPoints in = wrap(features, feature_rows, n, nullptr);
        m->mapper->attachInputDescriptors(in);  // what AddDescriptor in applyInputFilters attached
        m->mapper->processInput(in, wrapT(estimated_pose, m->dim + 1), time_stamp_seconds);
    });
}

int32_t b200mapper_process_raw_input(b200mapper* m, const float* features, int32_t feature_rows, int64_t n, const float* estimated_pose,
                                     double time_stamp_seconds, int64_t* n_filtered) {
    if (!m || (n > 0 && !features) || feature_rows != m->dim + 1) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] {
        DataPoints in = wrap(features, feature_rows, n, nullptr);
        m->mapper->setDeviceResidentInput(true);
        try {
            m->mapper->applyInputFilters(in);  // upload + filter chain + descriptors, all in the device slot
            if (n_filtered) *n_filtered = in.getNbPoints();
            m->mapper->processInput(in, wrapT(estimated_pose, m->dim + 1), time_stamp_seconds);
        } catch (...) {
            m->mapper->setDeviceResidentInput(false);
            throw;
        }
        m->mapper->setDeviceResidentInput(false);
    });
}

int32_t b200mapper_wait_for_map_update(b200mapper* m) {
    if (!m) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] { m->mapper->waitForMapUpdate(); });
}

int32_t b200mapper_map_update_in_flight(b200mapper* m) { return (m && m->mapper->mapUpdateInFlight()) ? 1 : 0; }

int32_t b200mapper_get_pose(b200mapper* m, float* pose) {
    if (!m || !pose) return B200ICP_ERR_INVALID_ARG;
    const TransformationParameters T = m->mapper->getPose();
    std::memcpy(pose, T.m, sizeof(float) * (size_t)(T.n * T.n));
    return B200ICP_OK;
}

int32_t b200mapper_get_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n) {
    if (!m || !n) return B200ICP_ERR_INVALID_ARG;
    int32_t rc2 = B200ICP_OK;
    const int32_t rc = guarded(m, [&] { rc2 = copy_out(m->mapper->getMap(), features, normals, capacity, n); });
    return rc != B200ICP_OK ? rc : rc2;
}

int32_t b200mapper_get_local_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n) {
    if (!m || !n) return B200ICP_ERR_INVALID_ARG;
    int32_t rc2 = B200ICP_OK;
    const int32_t rc = guarded(m, [&] { rc2 = copy_out(m->mapper->getMapObject().getLocalPointCloud(), features, normals, capacity, n); });
    return rc != B200ICP_OK ? rc : rc2;
}

int32_t b200mapper_get_new_local_map(b200mapper* m, float* features, float* normals, int64_t capacity, int64_t* n, int32_t* available) {
    if (!m || !n || !available) return B200ICP_ERR_INVALID_ARG;
    int32_t rc2 = B200ICP_OK;
    const int32_t rc = guarded(m, [&] {
        DataPoints d;
        d.dim = m->dim;
        *available = m->mapper->getNewLocalMap(d) ? 1 : 0;
        *n = 0;
        if (*available) rc2 = copy_out(d, features, normals, capacity, n);
    });
    return rc != B200ICP_OK ? rc : rc2;
}

int32_t b200mapper_vtk_save(const char* path, const float* features, int32_t feature_rows, int64_t n, const float* normals,
                            const float* prob, int32_t binary) {
    if (!path || n < 0 || (n > 0 && !features) || (feature_rows != 3 && feature_rows != 4)) return B200ICP_ERR_INVALID_ARG;
    return guarded(nullptr, [&] {
        DataPoints d = wrap(features, feature_rows, n, normals);
        if (prob) d.probabilityDynamic.assign(prob, prob + n);
        io::saveVTK(d, std::string(path), binary != 0);
    });
}

int32_t b200mapper_vtk_load(const char* path, int32_t dim, float* features, float* normals, float* prob, int64_t capacity, int64_t* n,
                            int32_t* has_normals, int32_t* has_prob) {
    if (!path || !n || (dim != 2 && dim != 3)) return B200ICP_ERR_INVALID_ARG;
    int32_t rc2 = B200ICP_OK;
    const int32_t rc = guarded(nullptr, [&] {
        const DataPoints d = io::loadVTK(std::string(path), dim);
        if (has_normals) *has_normals = d.normals.empty() ? 0 : 1;
        if (has_prob) *has_prob = d.probabilityDynamic.empty() ? 0 : 1;
        rc2 = copy_out(d, features, normals, capacity, n);
        if (rc2 == B200ICP_OK && features && prob && !d.probabilityDynamic.empty())
            std::memcpy(prob, d.probabilityDynamic.data(), d.probabilityDynamic.size() * sizeof(float));
    });
    return rc != B200ICP_OK ? rc : rc2;
}

int32_t b200mapper_set_map(b200mapper* m, const float* features, int32_t feature_rows, const float* normals, int64_t n) {
    return b200mapper_set_map_descriptors(m, features, feature_rows, normals, nullptr, n);
}

int32_t b200mapper_set_map_descriptors(b200mapper* m, const float* features, int32_t feature_rows, const float* normals, const float* prob, int64_t n) {
    if (!m || (n > 0 && !features) || feature_rows != m->dim + 1) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] {
        DataPoints d = wrap(features, feature_rows, n, normals);
        if (prob) d.probabilityDynamic.assign(prob, prob + n);
        m->mapper->setMap(d);
    });
}

int32_t b200mapper_get_map_prob(b200mapper* m, float* prob, int64_t capacity, int64_t* n) {
    if (!m || !n) return B200ICP_ERR_INVALID_ARG;
    return guarded(m, [&] {
        const DataPoints d = m->mapper->getMap();
        *n = (int64_t)d.probabilityDynamic.size();
        if (prob && *n > 0) {
            if (capacity < *n) throw InvalidParameter("capacity too small");
            std::memcpy(prob, d.probabilityDynamic.data(), d.probabilityDynamic.size() * sizeof(float));
        }
    });
}

int32_t b200mapper_get_is_mapping(const b200mapper* m) { return (m && m->mapper->getIsMapping()) ? 1 : 0; }
int32_t b200mapper_set_is_mapping(b200mapper* m, int32_t v) {
    if (!m) return B200ICP_ERR_INVALID_ARG;
    m->mapper->setIsMapping(v != 0);
    return B200ICP_OK;
}

int64_t b200mapper_trajectory_size(b200mapper* m) { return m ? (int64_t)m->mapper->getTrajectory().size() : 0; }
int32_t b200mapper_get_trajectory(b200mapper* m, float* poses, double* stamps, int64_t capacity) {
    if (!m || !poses) return B200ICP_ERR_INVALID_ARG;
    const auto tr = m->mapper->getTrajectory();
    const int nn = (m->dim + 1) * (m->dim + 1);
    for (int64_t i = 0; i < (int64_t)tr.size() && i < capacity; ++i) {
        std::memcpy(poses + i * nn, tr[i].first.m, sizeof(float) * (size_t)nn);
        if (stamps) stamps[i] = tr[i].second;
    }
    return B200ICP_OK;
}

int32_t b200mapper_get_stats(b200mapper* m, b200mapper_stats* out) {
    if (!m || !out) return B200ICP_ERR_INVALID_ARG;
    const b200icp_result& r = m->mapper->getICP().lastResult();
    out->overlap = r.overlap;
    out->iterations = r.iterations;
    out->map_updated = m->mapper->lastInputTriggeredMapUpdate() ? 1 : 0;
    out->n_window_updates = (int32_t)m->mapper->getMapObject().lastUpdates().size();
    out->n_local = m->mapper->getMapObject().localSize();
    out->n_global = m->mapper->getMapObject().globalSize();
    return B200ICP_OK;
}

int32_t b200mapper_get_window_updates(b200mapper* m, int32_t* out7, int32_t capacity) {
    if (!m || !out7) return B200ICP_ERR_INVALID_ARG;
    const auto u = m->mapper->getMapObject().lastUpdates();
    for (int i = 0; i < (int)u.size() && i < capacity; ++i)
        for (int j = 0; j < 7; ++j) out7[i * 7 + j] = u[i][j];
    return (int32_t)u.size();
}

}  // extern "C"
