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
    ICPSequence& sequence() { return icp; }
    //! host copy of an input that lives in the device scan slot (no-op for host inputs)
    DataPoints host(const DataPoints& in) { return icp.materialize(in); }
    //! the input in the device scan slot (uploaded now if it is a host cloud), descriptors reconciled with the map's
    //! (DataPoints::concatenate keeps the common ones): what every `map.concatenate(input)` of a module starts with
    DataPoints stage(const DataPoints& in) {
        DataPoints dev = icp.toDevice(in);
        icp.reconcileScanWithMap(getNbPointsGlobal() == 0);
        return dev;
    }
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
        const int xr = b200icp_map_extra_rows(icp.context());
        if (n > 0 && xr > 0) {
            out.descriptors.resize((size_t)n * xr);
            out.descriptorLabels = icp.mapLabels;
            ICPSequence::check(icp.context(), b200icp_map_download_extra(icp.context(), global, out.descriptors.data(), n));
        }
        return out;
    }
    //! the local map becomes `cloud` (all its descriptors); parked cells stay where they are
    void assignLocal(const DataPoints& cloud) {
        if (cloud.onDevice) throw std::runtime_error("assignLocal: host cloud expected");
        ICPSequence::check(icp.context(),
                           b200icp_map_replace_local(icp.context(), cloud.features.data(), cloud.dim + 1, cloud.getNbPoints(),
                                                     cloud.normals.empty() ? nullptr : cloud.normals.data(),
                                                     cloud.probabilityDynamic.empty() ? nullptr : cloud.probabilityDynamic.data(),
                                                     cloud.descriptors.empty() ? nullptr : cloud.descriptors.data(), cloud.getDescriptorRows()));
        icp.mapLabels = cloud.descriptors.empty() ? Labels() : cloud.descriptorLabels;
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
        // map.concatenate(inputPointsToKeep) (PointDistanceMapperModule.cpp:49) on the device: the scan is in the slot already
        // when Mapper::processInput uploaded it, else it goes there now -- with every descriptor it carries
        int64_t added = 0;
        map.stage(input);
        ICPSequence::check(map.context(), b200icp_scan_insert_point_distance(map.context(), minDistNewPoint, &added));
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
    void inPlaceUpdateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters&) override {
        int64_t n_after = 0;
        map.stage(input);  // map.concatenate(input) (OctreeMapperModule.cpp:37) takes the scan from the slot
        ICPSequence::check(map.context(), b200icp_scan_octree(map.context(), maxSizeByNode, maxPointByNode, samplingMethod, &n_after));
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
    void inPlaceCreateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters&) override {
        // createMap copies the input (DynamicPointsMapperModule.cpp:16-25): an unfiltered insert
        int64_t added = 0;
        if (map.getNbPoints() != 0) throw std::runtime_error("DynamicPointsMapperModule::createMap on a non-empty map");
        map.stage(input);
        ICPSequence::check(map.context(), b200icp_scan_append(map.context(), &added));
    }
    void inPlaceUpdateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) override {
        // (no concatenation here: the scan's descriptors stay as they are; only its probabilityDynamic is read)
        map.sequence().toDevice(input);
        ICPSequence::check(map.context(), b200icp_scan_dynamic_points(map.context(), pose.m, &prm));
    }
};

// ---- modules written against the REFERENCE's signature (MapperModules/MapperModule.h:20-29) ----------------------
// A third-party module that takes and returns host DataPoints drops in through this adapter: the local map is downloaded,
// handed to the module with the input (map frame) and the pose, and what the module leaves in `map` becomes the new local
// map on the device (all descriptors included).  It costs two PCIe trips of the local map per update -- the price of running
// host code on a device-resident map -- and nothing when no such module is configured.
class HostMapperModule {
   public:
    virtual ~HostMapperModule() = default;
    virtual DataPoints createMap(const DataPoints& input, const TransformationParameters& pose) = 0;
    virtual void inPlaceCreateMap(DataPoints& input, const TransformationParameters& pose) = 0;
    virtual DataPoints updateMap(const DataPoints& input, const DataPoints& map, const TransformationParameters& pose) = 0;
    virtual void inPlaceUpdateMap(const DataPoints& input, DataPoints& map, const TransformationParameters& pose) = 0;
};

class HostMapperModuleAdapter : public MapperModule {
    std::shared_ptr<HostMapperModule> inner;

   public:
    explicit HostMapperModuleAdapter(std::shared_ptr<HostMapperModule> m) : inner(std::move(m)) {}
    void inPlaceCreateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) override {
        map.assignLocal(inner->createMap(map.host(input), pose));
    }
    void inPlaceUpdateMap(const DataPoints& input, DeviceMap& map, const TransformationParameters& pose) override {
        DataPoints local = map.download(false, input.dim);
        inner->inPlaceUpdateMap(map.host(input), local, pose);
        map.assignLocal(local);
    }
};

}  // namespace norlab_icp_mapper_b200
