// Mapper.h -- host mirror of norlab_icp_mapper::Mapper (reference Mapper.{h,cpp}): same public
// methods, same processInput / shouldUpdateMap / updateMap sequence, the heavy steps delegated to
// libb200icp.so.  The configuration arrives as the struct the reference's loadYamlConfig would produce, or as the path of
// the YAML file itself (host/YamlConfig.h reads the subset of YAML the reference's configuration files use).
#pragma once
#include <atomic>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <chrono>
#include <utility>
#include <vector>

#include "Map.h"

namespace norlab_icp_mapper_b200 {

struct MapperConfig {
    b200icp_config icp;                                      // YAML `icp:` (Mapper.cpp:70-78)
    PostFilters post;                                        // YAML `post:` (Mapper.cpp:90-98)
    std::vector<b200icp_filter> inputFilters;                // YAML `input:` (Mapper.cpp:80-88): BoundingBox / DistanceLimit entries
    bool addProbabilityDynamic = false;                      // ... AddDescriptorDataPointsFilter{probabilityDynamic, 1, [value]}
    float probabilityDynamicValue = 0.6f;
    int inputSurfaceNormalKnn = 0;                           // ... SurfaceNormalDataPointsFilter{knn} on the reading; 0 = absent
    std::string mapUpdateCondition = "distance";            // mapper.updateCondition.type (Mapper.cpp:117-146)
    float mapUpdateValue = 1.0f;                             // ... .value (DEFAULT_MAP_UPDATE_DISTANCE, Mapper.h:20)
    float sensorMaxRange = 200.0f;                           // mapper.sensorMaxRange (Mapper.cpp:152-160)
    std::vector<std::pair<std::string, Parameters>> mapperModules;  // mapper.mapperModule (Mapper.cpp:162-172); empty -> default
    bool spillCellsToHostRam = false;  // cells the window leaves move to a RAMCellManager (host memory) instead of staying in HBM;
                                       // the constructor's saveMapCellsOnHardDrive = true selects the HardDriveCellManager
    std::string cellFolder = "/tmp/";  // ... and this is where it writes cell_<id>.vtk (HardDriveCellManager.h)
    std::vector<std::shared_ptr<MapperModule>> extraModules;        // ready-made module objects appended after the named ones (third-party
                                                                     // modules, e.g. a HostMapperModuleAdapter around a reference-signature one)
};

class Mapper {
    ICPSequence icp;
    PostFilters mapPostFilters;
    std::vector<b200icp_filter> inputFilters;
    bool addProbabilityDynamic;
    float probabilityDynamicValue;
    int inputSurfaceNormalKnn;
    std::string mapUpdateCondition;
    float mapUpdateOverlap = 0.f, mapUpdateDelay = 0.f, mapUpdateDistance = 1.0f;
    bool is3D, isOnline;
    std::atomic_bool isMapping;
    std::mutex icpMapLock;
    Map map;
    TransformationParameters pose;
    std::vector<std::pair<TransformationParameters, double>> trajectory;
    double lastTimeMapWasUpdated = 0.0;
    TransformationParameters lastPoseWhereMapWasUpdated;
    std::mutex poseLock, trajectoryLock;
    MapperModuleRegistrar registrar;
    bool lastInputUpdatedMap = false;
    bool deviceResidentInput = false;
    mutable std::future<void> mapUpdateFuture;  // isOnline: the asynchronous Map::updateLocalPointCloud in flight (Mapper.h, Mapper.cpp:282)

    void fillRegistrar();
    void updateMap(const DataPoints& currentInput, const TransformationParameters& currentPose, double currentTimeStamp);
    bool shouldUpdateMap(double currentTime, const TransformationParameters& currentPose, float currentOverlap) const;

   public:
    Mapper(const MapperConfig& config, bool is3D, bool isOnline, bool isMapping, bool saveMapCellsOnHardDrive, int device = 0);
    // the reference's signature (Mapper.cpp:15-33): configFilePath is the YAML file of Mapper::loadYamlConfig, read by host/YamlConfig.h
    Mapper(const std::string& configFilePath, bool is3D, bool isOnline, bool isMapping, bool saveMapCellsOnHardDrive, int device = 0);
    void applyInputFilters(DataPoints& inputInSensorFrame);
    // true: applyInputFilters leaves the filtered scan in the device slot (`onDevice` cloud) and processInput continues on it --
    // one host-to-device copy per scan.  false (default, the reference's contract): the caller gets the filtered cloud back.
    void setDeviceResidentInput(bool on) { deviceResidentInput = on; }
    DataPoints materialize(const DataPoints& cloud) { return icp.materialize(cloud); }
    // the descriptor-producing entries of the `input:` chain: AddDescriptorDataPointsFilter{probabilityDynamic} and
    // SurfaceNormalDataPointsFilter{knn} (normals on the reading, for SurfaceNormalOutlierFilter)
    void attachInputDescriptors(DataPoints& input) {
        if (addProbabilityDynamic && input.probabilityDynamic.empty()) input.probabilityDynamic.assign((size_t)input.getNbPoints(), probabilityDynamicValue);
        if (inputSurfaceNormalKnn > 0 && input.normals.empty() && input.getNbPoints() > 0) {
            input.normals.resize((size_t)input.getNbPoints() * input.dim);
            ICPSequence::check(icp.context(), b200icp_cloud_surface_normals(icp.context(), input.features.data(), input.dim + 1, input.getNbPoints(),
                                                                            inputSurfaceNormalKnn, input.normals.data()));
        }
    }
    void processInput(const DataPoints& filteredInputInSensorFrame, const TransformationParameters& estimatedPose, double timeStamp);
    DataPoints getMap();
    void setMap(const DataPoints& newMap);
    bool getNewLocalMap(DataPoints& mapOut);
    TransformationParameters getPose();
    bool getIsMapping() const;
    void setIsMapping(bool newIsMapping);
    std::vector<std::pair<TransformationParameters, double>> getTrajectory();
    // introspection for tests / benches
    Map& getMapObject() { return map; }
    ICPSequence& getICP() { return icp; }
    bool lastInputTriggeredMapUpdate() const { return lastInputUpdatedMap; }
    //! isOnline: block until the map update in flight (if any) and the queued cell-window updates are done; rethrows the update's exception
    void waitForMapUpdate() {
        if (mapUpdateFuture.valid()) mapUpdateFuture.get();
        map.waitForWindowUpdates();
    }
    bool mapUpdateInFlight() const {
        return mapUpdateFuture.valid() && mapUpdateFuture.wait_for(std::chrono::milliseconds(0)) != std::future_status::ready;
    }
    ~Mapper() {
        if (mapUpdateFuture.valid()) mapUpdateFuture.wait();
    }
};

}  // namespace norlab_icp_mapper_b200
