// Mapper.cpp -- see Mapper.h.  Reference: /root/reference/norlab_icp_mapper/Mapper.cpp.
#include "Mapper.h"

#include "YamlConfig.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace norlab_icp_mapper_b200 {

void Mapper::fillRegistrar() {  // Mapper.cpp:9-13
    registrar.add("PointDistanceMapperModule", [](const Parameters& p) { return std::make_shared<PointDistanceMapperModule>(p); });
    registrar.add("OctreeMapperModule", [](const Parameters& p) { return std::make_shared<OctreeMapperModule>(p); });
    registrar.add("DynamicPointsMapperModule", [](const Parameters& p) { return std::make_shared<DynamicPointsMapperModule>(p); });
}

Mapper::Mapper(const MapperConfig& config, bool is3D_, bool isOnline_, bool isMapping_, bool saveMapCellsOnHardDrive, int device)
    : icp(config.icp, device),
      mapPostFilters(config.post),
      inputFilters(config.inputFilters),
      addProbabilityDynamic(config.addProbabilityDynamic),
      probabilityDynamicValue(config.probabilityDynamicValue),
      inputSurfaceNormalKnn(config.inputSurfaceNormalKnn),
      mapUpdateCondition(config.mapUpdateCondition),
      is3D(is3D_),
      isOnline(isOnline_),
      isMapping(isMapping_),
      map(is3D_, isOnline_, icp, icpMapLock,
          saveMapCellsOnHardDrive ? std::unique_ptr<CellManager>(new HardDriveCellManager(is3D_ ? 3 : 2, config.cellFolder))
                                  : (config.spillCellsToHostRam ? std::unique_ptr<CellManager>(new RAMCellManager()) : nullptr)),
      pose(TransformationParameters::Identity(is3D_ ? 4 : 3)),
      lastPoseWhereMapWasUpdated(TransformationParameters::Identity(is3D_ ? 4 : 3)) {
    if ((config.icp.dim == 3) != is3D_) throw InvalidParameter("icp.dim does not match is3D");
    fillRegistrar();
    // the validation of loadYamlConfig (Mapper.cpp:117-160)
    if (mapUpdateCondition == "distance") {
        mapUpdateDistance = config.mapUpdateValue;
        if (mapUpdateDistance < 0) throw InvalidParameter("Invalid map update distance: " + std::to_string(mapUpdateDistance));
    } else if (mapUpdateCondition == "overlap") {
        mapUpdateOverlap = config.mapUpdateValue;
        if (mapUpdateOverlap < 0 || mapUpdateOverlap > 1) throw InvalidParameter("Invalid map update overlap: " + std::to_string(mapUpdateOverlap));
    } else if (mapUpdateCondition == "delay") {
        mapUpdateDelay = config.mapUpdateValue;
        if (mapUpdateDelay < 0) throw InvalidParameter("Invalid map update delay: " + std::to_string(mapUpdateDelay));
    } else {
        throw InvalidParameter("Invalid map update condition: " + mapUpdateCondition);
    }
    if (config.sensorMaxRange < 0) throw InvalidParameter("Invalid sensor max range: " + std::to_string(config.sensorMaxRange));
    map.setSensorMaxRange(config.sensorMaxRange);
    if (config.mapperModules.empty() && config.extraModules.empty()) {  // setDefaultMapperModule (Mapper.cpp:330-336)
        map.addMapperModule(registrar.create("PointDistanceMapperModule", Parameters{{"minDistNewPoint", "0.15"}}));
    } else {
        for (const auto& m : config.mapperModules) map.addMapperModule(registrar.create(m.first, m.second));
        for (const auto& m : config.extraModules) map.addMapperModule(m);
    }
}

Mapper::Mapper(const std::string& configFilePath, bool is3D_, bool isOnline_, bool isMapping_, bool saveMapCellsOnHardDrive, int device)
    : Mapper(loadYamlConfig(configFilePath, is3D_), is3D_, isOnline_, isMapping_, saveMapCellsOnHardDrive, device) {}

// Mapper.cpp:187-191: radiusFilter = DistanceLimitDataPointsFilter{dim -1, dist sensorMaxRange,
// removeInside 0} (built at Mapper.cpp:27-31), then the YAML `input:` chain -- one predicate kernel +
// ordered compaction on the device (b200icp_filter_cloud); AddDescriptor attaches a constant.
void Mapper::applyInputFilters(DataPoints& in) {
    std::vector<b200icp_filter> chain;
    b200icp_filter radius{};
    radius.kind = B200ICP_FILTER_DISTANCE_LIMIT;
    radius.dim = -1;
    radius.dist = map.getSensorMaxRange();
    radius.remove_inside = 0;
    chain.push_back(radius);
    chain.insert(chain.end(), inputFilters.begin(), inputFilters.end());
    // The raw scan goes to the device ONCE, with every descriptor it carries; the chain runs on that copy (descriptors follow
    // the surviving points) and, with setDeviceResidentInput(true), stays there for processInput.
    DataPoints dev = icp.toDevice(in);
    int64_t n = dev.deviceCount;
    ICPSequence::check(icp.context(), b200icp_scan_filter(icp.context(), chain.data(), (int32_t)chain.size(), &n));
    dev.deviceCount = n;
    if (addProbabilityDynamic && !icp.scanHas("probabilityDynamic")) ICPSequence::check(icp.context(), b200icp_scan_add_prob(icp.context(), probabilityDynamicValue));
    if (inputSurfaceNormalKnn > 0 && !icp.scanHas("normals") && n > 0)
        ICPSequence::check(icp.context(), b200icp_scan_surface_normals(icp.context(), inputSurfaceNormalKnn));
    in = deviceResidentInput ? dev : icp.materialize(dev);
}

// Mapper.cpp:194-238
namespace {
struct StepTimer {  // B200MAPPER_TIMING=1 prints the wall time of each step of processInput (development aid)
    bool on = std::getenv("B200MAPPER_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "  [mapper] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

void Mapper::processInput(const DataPoints& filteredInputInSensorFrame, const TransformationParameters& estimatedPose, double timeStamp) {
    StepTimer timer;
    // One upload: the scan goes to the context's device slot (it is there already when applyInputFilters left it there) and the
    // steps below (two rigid transforms, icp(input), the map update) work on that copy.  B200MAPPER_HOST_SCAN=1 takes the host path.
    static const bool hostScan = std::getenv("B200MAPPER_HOST_SCAN") != nullptr;
    const bool deviceScan = filteredInputInSensorFrame.onDevice || (!hostScan && filteredInputInSensorFrame.getNbPoints() > 0);
    DataPoints input = deviceScan ? rigidTransform(icp, icp.toDevice(filteredInputInSensorFrame), estimatedPose)
                                  : rigidTransform(icp, filteredInputInSensorFrame, estimatedPose);
    timer.lap("transform(input, T_est)");
    lastInputUpdatedMap = false;

    TransformationParameters correctedPose;
    if (map.isLocalPointCloudEmpty()) {
        correctedPose = estimatedPose;
        map.updatePose(correctedPose);
        updateMap(input, correctedPose, timeStamp);
    } else {
        TransformationParameters correction;
        {
            std::lock_guard<std::mutex> icpMapLockGuard(icpMapLock);
            correction = icp(input);
        }
        timer.lap("icp(input)");
        correctedPose = correction * estimatedPose;
        map.updatePose(correctedPose);
        timer.lap("map.updatePose");
        if (shouldUpdateMap(timeStamp, correctedPose, icp.getOverlap())) {
            DataPoints corrected = rigidTransform(icp, input, correction);
            timer.lap("transform(input, correction)");
            updateMap(corrected, correctedPose, timeStamp);
            timer.lap("updateMap");
        }
    }
    // Mapper.cpp:225-228: collect a finished asynchronous update (its exceptions surface here).  The reference waits up to 1 ms for
    // it -- nothing next to a CPU registration, but as long as a whole registration here: a poll instead.
    if (mapUpdateFuture.valid() && mapUpdateFuture.wait_for(std::chrono::milliseconds(0)) == std::future_status::ready) mapUpdateFuture.get();
    {
        std::lock_guard<std::mutex> lock(poseLock);
        pose = correctedPose;
    }
    {
        std::lock_guard<std::mutex> lock(trajectoryLock);
        trajectory.emplace_back(correctedPose, timeStamp);
    }
}

// Mapper.cpp:240-272
bool Mapper::shouldUpdateMap(double currentTime, const TransformationParameters& currentPose, float currentOverlap) const {
    if (!isMapping.load()) return false;
    // Mapper.cpp:248-255: if previous update is not over
    if (isOnline && mapUpdateInFlight()) return false;
    if (mapUpdateCondition == "overlap") return currentOverlap < mapUpdateOverlap;
    if (mapUpdateCondition == "delay") return (currentTime - lastTimeMapWasUpdated) > (double)mapUpdateDelay;
    const int euclideanDim = is3D ? 3 : 2;
    float d2 = 0.f;
    for (int c = 0; c < euclideanDim; ++c) {
        const float d = currentPose(c, euclideanDim) - lastPoseWhereMapWasUpdated(c, euclideanDim);
        d2 += d * d;
    }
    return std::fabs(std::sqrt(d2)) > mapUpdateDistance;
}

// Mapper.cpp:274-288
void Mapper::updateMap(const DataPoints& currentInput, const TransformationParameters& currentPose, double currentTimeStamp) {
    lastTimeMapWasUpdated = currentTimeStamp;
    lastPoseWhereMapWasUpdated = currentPose;
    if (isOnline && !map.isLocalPointCloudEmpty()) {  // Mapper.cpp:280-283
        if (mapUpdateFuture.valid()) mapUpdateFuture.get();  // (ready: shouldUpdateMap saw to that)
        // the reference copies currentInput into the async call: the device-resident scan is handed over as a snapshot, so the next
        // processInput can upload into the slot while the worker still reads this one
        if (currentInput.onDevice) icp.snapshotScan();
        mapUpdateFuture = std::async(std::launch::async, &Map::updateLocalPointCloud, &map, currentInput, currentPose, mapPostFilters, true);
    } else {
        map.updateLocalPointCloud(currentInput, currentPose, mapPostFilters);
    }
    lastInputUpdatedMap = true;
}

DataPoints Mapper::getMap() { return map.getGlobalPointCloud(); }

void Mapper::setMap(const DataPoints& newMap) {
    map.setGlobalPointCloud(newMap);
    std::lock_guard<std::mutex> lock(trajectoryLock);
    trajectory.clear();
}

bool Mapper::getNewLocalMap(DataPoints& mapOut) { return map.getNewLocalPointCloud(mapOut); }

TransformationParameters Mapper::getPose() {
    std::lock_guard<std::mutex> lock(poseLock);
    return pose;
}

bool Mapper::getIsMapping() const { return isMapping.load(); }
void Mapper::setIsMapping(bool newIsMapping) { isMapping.store(newIsMapping); }

std::vector<std::pair<TransformationParameters, double>> Mapper::getTrajectory() {
    std::lock_guard<std::mutex> lock(trajectoryLock);
    return trajectory;
}

}  // namespace norlab_icp_mapper_b200
