// MapperModule.h -- the plugin surface of the reference (MapperModules/MapperModule.h:12-33) with the
// map living on the device: the four operations keep their names and argument meaning; `DataPoints&
// map` becomes `DeviceMap& map` (a handle on the context's device-resident local map) so a module
// never drags the map through host memory.  Registration by class name + string parameters mirrors
// ADD_TO_REGISTRAR / createFromYAML (Mapper.cpp:9-13,167-171); unknown parameters throw
// InvalidParameter like PM::Parametrizable does (OctreeMapperModule.cpp:6-11).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>

#include "DataPoints.h"
#include "ICPSequence.h"

namespace norlab_icp_mapper_b200 {

typedef std::map<std::string, std::string> Parameters;

class DeviceMap {
    ICPSequence& icp;

   public:
    explicit DeviceMap(ICPSequence& icp_) : icp(icp_) {}
    b200icp_ctx* context() { return icp.context(); }
    //! host copy of an input that lives in the device scan slot (no-op for host inputs)
    DataPoints host(const DataPoints& in) { return icp.materialize(in); }
    int64_t getNbPoints() {
        int64_t nl = 0, ng = 0;
        b200icp_map_counts(icp.context(), &nl, &ng);
        return nl;
    }
    int64_t getNbPointsGlobal() {
        int64_t nl = 0, ng = 0;
        b200icp_map_counts(icp.context(), &nl, &ng);
        return ng;
    }
    DataPoints download(bool global, int dim) {
        DataPoints out;
        out.dim = dim;
        int64_t n = 0;
        ICPSequence::check(icp.context(), b200icp_map_download(icp.context(), global, nullptr, nullptr, 0, &n));
        out.features.resize((size_t)n * (dim + 1));
        const bool has_n = b200icp_map_has_normals(icp.context()) != 0;
        if (has_n) out.normals.resize((size_t)n * dim);
        if (n > 0)
            ICPSequence::check(icp.context(), b200icp_map_download(icp.context(), global, out.features.data(),
                                                                   has_n ? out.normals.data() : nullptr, n, &n));
        if (n > 0 && b200icp_map_has_prob(icp.context())) {
            out.probabilityDynamic.resize((size_t)n);
            ICPSequence::check(icp.context(), b200icp_map_download_prob(icp.context(), global, out.probabilityDynamic.data(), n));
        }
        return out;
    }
};

class MapperModule {
   public:
    virtual ~MapperModule() = default;
    //! Create a map from the input cloud (the device map is empty on entry).
    virtual void inPlaceCreateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) = 0;
    //! Update the map with the input cloud; the input is in the map frame.
    virtual void inPlaceUpdateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) = 0;
    //! Non-destructive flavours of the reference return the resulting cloud: here, a download after the in-place call.
    DataPoints createMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) {
        inPlaceCreateMap(input, map, pose);
        return map.download(false, input.dim);
    }
    DataPoints updateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) {
        inPlaceUpdateMap(input, map, pose);
        return map.download(false, input.dim);
    }
};

class MapperModuleRegistrar {
    std::map<std::string, std::function<std::shared_ptr<MapperModule>(const Parameters&)>> factories;

   public:
    void add(const std::string& name, std::function<std::shared_ptr<MapperModule>(const Parameters&)> f) { factories[name] = std::move(f); }
    std::shared_ptr<MapperModule> create(const std::string& name, const Parameters& params) const {
        auto it = factories.find(name);
        if (it == factories.end()) throw InvalidParameter("Trying to instantiate unknown MapperModule " + name);
        return it->second(params);
    }
};

// PointDistanceMapperModule (MapperModules/PointDistanceMapperModule.{h,cpp}): parameter
// minDistNewPoint (default 0.15, Mapper.cpp:330-336), insert input points farther than that from the map.
class PointDistanceMapperModule : public MapperModule {
    float minDistNewPoint;

   public:
    explicit PointDistanceMapperModule(const Parameters& params) : minDistNewPoint(0.15f) {
        for (const auto& kv : params) {
            if (kv.first != "minDistNewPoint") throw InvalidParameter("PointDistanceMapperModule: unknown parameter " + kv.first);
            minDistNewPoint = std::stof(kv.second);
            if (minDistNewPoint < 0.f) throw InvalidParameter("PointDistanceMapperModule: minDistNewPoint must be >= 0");
        }
    }
    void inPlaceCreateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters&) override {
        // "keep the input untouched" (PointDistanceMapperModule.cpp:16-19): the map is the input
        insert(input, map);
    }
    void inPlaceUpdateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters&) override { insert(input, map); }

   private:
    void insert(const DataPoints& input, DeviceMap& map) {
        int64_t added = 0;
        if (input.onDevice) {  // the scan is already on the device (Mapper::processInput uploaded it once)
            ICPSequence::check(map.context(), b200icp_scan_insert_point_distance(map.context(), minDistNewPoint, &added));
            return;
        }
        ICPSequence::check(map.context(), b200icp_map_insert_point_distance(map.context(), input.features.data(), input.dim + 1,
                                                                            input.getNbPoints(), input.normals.empty() ? nullptr : input.normals.data(),
                                                                            minDistNewPoint, &added, nullptr));
    }
};

// OctreeMapperModule (MapperModules/OctreeMapperModule.{h,cpp}): map.concatenate(input) then
// libpointmatcher's OctreeGridDataPointsFilter over the whole map, parameters passed through.
class OctreeMapperModule : public MapperModule {
    bool buildParallel = true;
    int maxPointByNode = 1;
    float maxSizeByNode = 0.f;
    int samplingMethod = 0;

   public:
    explicit OctreeMapperModule(const Parameters& params) {
        for (const auto& kv : params) {
            if (kv.first == "buildParallel") buildParallel = std::stoi(kv.second) != 0;
            else if (kv.first == "maxPointByNode") maxPointByNode = std::stoi(kv.second);
            else if (kv.first == "maxSizeByNode") maxSizeByNode = std::stof(kv.second);
            else if (kv.first == "samplingMethod") samplingMethod = std::stoi(kv.second);
            else throw InvalidParameter("OctreeMapperModule: unknown parameter " + kv.first);
        }
        if (maxPointByNode < 1) throw InvalidParameter("OctreeMapperModule: maxPointByNode must be >= 1");
        if (!(maxSizeByNode > 0.f)) throw InvalidParameter("OctreeMapperModule: maxSizeByNode must be > 0 on this implementation");
        (void)buildParallel;  // the device build is always parallel
    }
    void inPlaceCreateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) override {
        inPlaceUpdateMap(input, map, pose);  // inPlaceUpdateMap(emptyMap, input): concatenate into an empty map, then filter
    }
    void inPlaceUpdateMap(const DataPoints& input_, DeviceMap& map, const TransformationParameters&) override {
        const DataPoints hostCopy = input_.onDevice ? map.host(input_) : DataPoints();  // (no device entry point for this module yet)
        const DataPoints& input = input_.onDevice ? hostCopy : input_;
        int64_t n_after = 0;
        ICPSequence::check(map.context(),
                           b200icp_map_octree(map.context(), input.features.data(), input.dim + 1, input.getNbPoints(),
                                              input.normals.empty() ? nullptr : input.normals.data(),
                                              input.probabilityDynamic.empty() ? nullptr : input.probabilityDynamic.data(), maxSizeByNode,
                                              maxPointByNode, samplingMethod, &n_after));
    }
};

// DynamicPointsMapperModule (MapperModules/DynamicPointsMapperModule.{h,cpp}): Bayesian update of the
// map's probabilityDynamic from the new scan; parameters and defaults of the reference (:33-44).
class DynamicPointsMapperModule : public MapperModule {
    b200icp_dynamic_params prm{0.6f, 0.8f, 0.99f, 0.01f, 0.01f, 0.01f, 200.f};

   public:
    explicit DynamicPointsMapperModule(const Parameters& params) {
        for (const auto& kv : params) {
            const float v = std::stof(kv.second);
            if (kv.first == "thresholdDynamic") prm.threshold_dynamic = v;
            else if (kv.first == "alpha") prm.alpha = v;
            else if (kv.first == "beta") prm.beta = v;
            else if (kv.first == "beamHalfAngle") prm.beam_half_angle = v;
            else if (kv.first == "epsilonA") prm.epsilon_a = v;
            else if (kv.first == "epsilonD") prm.epsilon_d = v;
            else if (kv.first == "sensorMaxRange") prm.sensor_max_range = v;
            else throw InvalidParameter("DynamicPointsMapperModule: unknown parameter " + kv.first);
        }
    }
    void inPlaceCreateMap(const DataPoints& input_, DeviceMap& map, const TransformationParameters&) override {
        const DataPoints hostCopy = input_.onDevice ? map.host(input_) : DataPoints();
        const DataPoints& input = input_.onDevice ? hostCopy : input_;
        // createMap copies the input (DynamicPointsMapperModule.cpp:16-25): an unfiltered insert
        int64_t added = 0;
        if (map.getNbPoints() != 0) throw std::runtime_error("DynamicPointsMapperModule::createMap on a non-empty map");
        ICPSequence::check(map.context(), b200icp_map_append(map.context(), input.features.data(), input.dim + 1, input.getNbPoints(),
                                                             input.normals.empty() ? nullptr : input.normals.data(),
                                                             input.probabilityDynamic.empty() ? nullptr : input.probabilityDynamic.data(), &added));
    }
    void inPlaceUpdateMap(const DataPoints& input_, DeviceMap& map, const TransformationParameters& pose) override {
        const DataPoints hostCopy = input_.onDevice ? map.host(input_) : DataPoints();
        const DataPoints& input = input_.onDevice ? hostCopy : input_;
        ICPSequence::check(map.context(), b200icp_map_dynamic_points(map.context(), input.features.data(), input.dim + 1, input.getNbPoints(),
                                                                     input.probabilityDynamic.empty() ? nullptr : input.probabilityDynamic.data(),
                                                                     pose.m, &prm));
    }
};

}  // namespace norlab_icp_mapper_b200
