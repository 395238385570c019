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
        ICPSequence::check(map.context(), b200icp_map_insert_point_distance(map.context(), input.features.data(), input.dim + 1,
                                                                            input.getNbPoints(), input.normals.empty() ? nullptr : input.normals.data(),
                                                                            minDistNewPoint, &added, nullptr));
    }
};

}  // namespace norlab_icp_mapper_b200
