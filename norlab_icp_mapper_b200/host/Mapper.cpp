// Mapper.cpp -- see Mapper.h.  Reference: /root/reference/norlab_icp_mapper/Mapper.cpp.
#include "Mapper.h"

#include <cmath>

namespace norlab_icp_mapper_b200 {

void Mapper::fillRegistrar() {  // Mapper.cpp:9-13 (Octree / DynamicPoints modules: DESIGN.md section 7, next)
    registrar.add("PointDistanceMapperModule", [](const Parameters& p) { return std::make_shared<PointDistanceMapperModule>(p); });
}

Mapper::Mapper(const MapperConfig& config, bool is3D_, bool isOnline_, bool isMapping_, bool /*saveMapCellsOnHardDrive*/, int device)
    : icp(config.icp, device),
      mapPostFilters(config.post),
      mapUpdateCondition(config.mapUpdateCondition),
      is3D(is3D_),
      isOnline(isOnline_),
      isMapping(isMapping_),
      map(is3D_, isOnline_, icp, icpMapLock),
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
    if (config.mapperModules.empty()) {  // setDefaultMapperModule (Mapper.cpp:330-336)
        map.addMapperModule(registrar.create("PointDistanceMapperModule", Parameters{{"minDistNewPoint", "0.15"}}));
    } else {
        for (const auto& m : config.mapperModules) map.addMapperModule(registrar.create(m.first, m.second));
    }
}

// Mapper.cpp:187-191: radiusFilter = DistanceLimitDataPointsFilter{dim -1, dist sensorMaxRange,
// removeInside 0} (built at Mapper.cpp:27-31) keeps ||p|| < dist; then the YAML `input:` chain.
// O(N) predicate + compaction on the caller's host buffer before the single upload; the device-side
// chain is the first "next" row of SURVEY 8f.
void Mapper::applyInputFilters(DataPoints& in) {
    const int rows = in.dim + 1;
    const float r = std::fabs(map.getSensorMaxRange());
    const int64_t n = in.getNbPoints();
    int64_t o = 0;
    const bool has_n = !in.normals.empty();
    for (int64_t i = 0; i < n; ++i) {
        float d2 = 0.f;
        for (int c = 0; c < in.dim; ++c) d2 += in.features[i * rows + c] * in.features[i * rows + c];
        if (std::sqrt(d2) < r) {
            if (o != i) {
                for (int c = 0; c < rows; ++c) in.features[o * rows + c] = in.features[i * rows + c];
                if (has_n)
                    for (int c = 0; c < in.dim; ++c) in.normals[o * in.dim + c] = in.normals[i * in.dim + c];
            }
            ++o;
        }
    }
    in.features.resize((size_t)o * rows);
    if (has_n) in.normals.resize((size_t)o * in.dim);
}

// Mapper.cpp:194-238
void Mapper::processInput(const DataPoints& filteredInputInSensorFrame, const TransformationParameters& estimatedPose, double timeStamp) {
    DataPoints input = rigidTransform(icp, filteredInputInSensorFrame, estimatedPose);
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
        correctedPose = correction * estimatedPose;
        map.updatePose(correctedPose);
        if (shouldUpdateMap(timeStamp, correctedPose, icp.getOverlap())) {
            updateMap(rigidTransform(icp, input, correction), correctedPose, timeStamp);
        }
    }
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
    // isOnline: "previous update is not over" never holds -- updates complete inside updateMap
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
    // isOnline && map non-empty: the reference hands this to std::async; the device update takes
    // milliseconds, so it is done in line (async overlap on a second stream: SURVEY 8f rank 3)
    map.updateLocalPointCloud(currentInput, currentPose, mapPostFilters);
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
