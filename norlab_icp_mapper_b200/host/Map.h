// Map.h -- host mirror of norlab_icp_mapper::Map (reference Map.{h,cpp}) over the device-resident map.
// The local cloud, the parked cells and the index all live in HBM (libb200icp.so); this class keeps
// the reference's bookkeeping: the 20 m cell window around the pose and the order of the update steps.
#pragma once
#include <array>
#include <atomic>
#include <list>
#include <thread>
#include <memory>
#include <mutex>
#include <vector>

#include <unordered_set>

#include "CellManager.h"
#include "MapperModule.h"

namespace norlab_icp_mapper_b200 {

struct PostFilters {           // the `post:` chain entries this path implements (examples/config.yaml:25-32)
    int surfaceNormalKnn = 0;  // SurfaceNormalDataPointsFilter{knn}; 0 = absent
    bool cutAtThreshold = false;       // CutAtDescriptorThresholdDataPointsFilter{descName probabilityDynamic, ...}
    bool cutUseLargerThan = true;
    float cutThreshold = 0.65f;
};

class Map {
    struct Update {
        int start[3], end[3];  // rows, columns, aisles
        bool load;
    };
    static constexpr int BUFFER_SIZE = 2;                        // Map.h:30
    static constexpr float CELL_SIZE = 20.0f;                    // Map.h:31
    static constexpr float DEFAULT_SENSOR_MAX_RANGE = 200.0f;    // Map.h:33

    float sensorMaxRange = DEFAULT_SENSOR_MAX_RANGE;
    bool is3D;
    bool isOnline;
    ICPSequence& icp;
    std::mutex& icpMapLock;
    DeviceMap localPointCloud;
    std::mutex localPointCloudLock;
    int inferiorLastUpdateIndex[3] = {0, 0, 0};  // row, column, aisle
    int superiorLastUpdateIndex[3] = {0, 0, 0};
    bool newLocalPointCloudAvailable = false;
    std::atomic_bool localPointCloudEmpty{true};
    std::atomic_bool firstPoseUpdate{true};
    std::vector<std::shared_ptr<MapperModule>> mapperModuleVec;
    // Spill tier (Map.cpp:20-27): null = the cells the window leaves stay in device memory (flag only); else they move to this
    // manager when unloaded and come back through it when loaded.  loadedCellIds as in the reference (Map.h:44, Map.cpp:79-99).
    std::unique_ptr<CellManager> cellManager;
    std::mutex cellManagerLock;
    std::unordered_set<std::string> loadedCellIds;
    void spillUnloaded();
    void loadSpilledCells(const Update& update);
    std::vector<Update> appliedUpdates;  // log of the slabs of the last updatePose (tests)
    // isOnline: cell-window updates are queued and applied by `updateThread` (Map.cpp:29-57,482-494)
    std::list<Update> updateList;
    std::mutex updateListLock;
    std::atomic_bool updateThreadLooping{true};
    std::atomic_int updatesInFlight{0};
    std::thread updateThread;
    void updateThreadFunction();

    void applyUpdate(const Update& update);
    int toInferiorGridCoordinate(float worldCoordinate, float range) const;
    int toSuperiorGridCoordinate(float worldCoordinate, float range) const;
    void scheduleUpdate(const Update& update);

   public:
    Map(bool is3D, bool isOnline, ICPSequence& icp, std::mutex& icpMapLock, std::unique_ptr<CellManager> cellManager = nullptr);
    CellManager* getCellManager() { return cellManager.get(); }
    ~Map();
    void updatePose(const TransformationParameters& pose);
    DataPoints getLocalPointCloud();
    // asynchronous = true: called on a worker thread while registrations go on (Mapper::updateMap, isOnline): the update steps build a
    // second index that replaces the live one in the final icp.setMap (b200icp_map_begin_update / b200icp_map_end_update)
    void updateLocalPointCloud(DataPoints input, TransformationParameters pose, PostFilters postFilters, bool asynchronous = false);
    //! isOnline: block until the queued cell-window updates have been applied (tests, orderly shutdown)
    void waitForWindowUpdates();
    bool getNewLocalPointCloud(DataPoints& localPointCloudOut);
    DataPoints getGlobalPointCloud();
    void setGlobalPointCloud(const DataPoints& newLocalPointCloud);
    bool isLocalPointCloudEmpty() const { return localPointCloudEmpty.load(); }
    void addMapperModule(std::shared_ptr<MapperModule> mapperModule) { mapperModuleVec.push_back(std::move(mapperModule)); }
    void setSensorMaxRange(float r) { sensorMaxRange = r; }
    float getSensorMaxRange() const { return sensorMaxRange; }
    int64_t localSize() { return localPointCloud.getNbPoints(); }
    int64_t globalSize() { return localPointCloud.getNbPointsGlobal(); }
    // slabs applied by the last updatePose: {start row, end row, start col, end col, start aisle, end aisle, load}
    std::vector<std::array<int, 7>> lastUpdates() const;
};

}  // namespace norlab_icp_mapper_b200
