// Map.cpp -- see Map.h.  Reference: /root/reference/norlab_icp_mapper/Map.cpp.
#include "Map.h"

#include <chrono>
#include <cmath>
#include <limits>

namespace norlab_icp_mapper_b200 {

Map::Map(bool is3D_, bool isOnline_, ICPSequence& icp_, std::mutex& icpMapLock_)
    : is3D(is3D_), isOnline(isOnline_), icp(icp_), icpMapLock(icpMapLock_), localPointCloud(icp_) {
    if (isOnline) updateThread = std::thread(&Map::updateThreadFunction, this);  // Map.cpp:29-32
}

Map::~Map() {  // Map.cpp:34-41
    if (isOnline) {
        updateThreadLooping.store(false);
        if (updateThread.joinable()) updateThread.join();
    }
}

// Map.cpp:43-57.  loadCells / unloadCells end with icp.setMap(localPointCloud) (Map.cpp:110-112,177-179): here one index rebuild once
// the queue has drained, under the same two locks.
void Map::updateThreadFunction() {
    while (updateThreadLooping.load()) {
        Update update{};
        bool have = false, last = false;
        {
            std::lock_guard<std::mutex> lock(updateListLock);
            if (!updateList.empty()) {
                update = updateList.front();
                updateList.pop_front();
                have = true;
                last = updateList.empty();
            }
        }
        if (have) {
            try {
                applyUpdate(update);
                if (last) {
                    std::lock_guard<std::mutex> l1(localPointCloudLock);
                    std::lock_guard<std::mutex> l2(icpMapLock);
                    ICPSequence::check(icp.context(), b200icp_map_commit(icp.context()));
                    localPointCloudEmpty.store(localPointCloud.getNbPoints() == 0);
                }
            } catch (const std::exception&) {
                // (a failed window update leaves the previous window in place; the next registration reports device errors)
            }
            updatesInFlight.fetch_sub(1);
        } else {
            std::this_thread::sleep_for(std::chrono::duration<float>(0.01f));
        }
    }
}

void Map::waitForWindowUpdates() {
    while (updatesInFlight.load() > 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
}

// Map.cpp:472-480
int Map::toInferiorGridCoordinate(float worldCoordinate, float range) const { return (int)std::ceil(((worldCoordinate - range) / CELL_SIZE) - 1.0); }
int Map::toSuperiorGridCoordinate(float worldCoordinate, float range) const { return (int)std::floor((worldCoordinate + range) / CELL_SIZE); }

// Map.cpp:59-69 applyUpdate -> loadCells (:71-128) / unloadCells (:140-230): on the device a slab
// update flips the `loaded` flag of the points concerned; icp.setMap follows once per updatePose.
void Map::applyUpdate(const Update& u) {
    const int32_t slab[6] = {u.start[0], u.end[0], u.start[1], u.end[1], u.start[2], u.end[2]};
    int64_t changed = 0;
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    ICPSequence::check(icp.context(), b200icp_map_window(icp.context(), u.load ? 1 : 0, slab, &changed));
    if (changed > 0) newLocalPointCloudAvailable = true;
    appliedUpdates.push_back(u);
}

void Map::scheduleUpdate(const Update& update) {  // Map.cpp:482-494
    if (isOnline) {
        updatesInFlight.fetch_add(1);
        std::lock_guard<std::mutex> lock(updateListLock);
        updateList.push_back(update);
    } else {
        applyUpdate(update);
    }
}

// Map.cpp:246-460.  Per axis a (row, column, aisle) the window is [inferior - BUFFER, superior + BUFFER]
// in 20 m cells; an edge that moved by >= 2 cells loads / unloads the slab it swept, spanning the
// current extent of the other axes.
void Map::updatePose(const TransformationParameters& pose) {
    const int positionColumn = is3D ? 3 : 2;
    const int axes = is3D ? 3 : 2;
    {
        std::lock_guard<std::mutex> lock(localPointCloudLock);
        appliedUpdates.clear();
    }
    if (firstPoseUpdate.load()) {
        for (int a = 0; a < axes; ++a) {
            inferiorLastUpdateIndex[a] = toInferiorGridCoordinate(pose(a, positionColumn), sensorMaxRange);
            superiorLastUpdateIndex[a] = toSuperiorGridCoordinate(pose(a, positionColumn), sensorMaxRange);
        }
        // cellManager->clearAllCells(); unload everything; load the window (Map.cpp:260-271)
        const int lo = std::numeric_limits<int>::lowest() / 2, hi = std::numeric_limits<int>::max() / 2;
        Update all{{lo, lo, lo}, {hi, hi, hi}, false};
        applyUpdate(all);
        Update win{};
        win.load = true;
        for (int a = 0; a < 3; ++a) {
            win.start[a] = inferiorLastUpdateIndex[a] - BUFFER_SIZE;
            win.end[a] = superiorLastUpdateIndex[a] + BUFFER_SIZE;
        }
        applyUpdate(win);
        firstPoseUpdate.store(false);
    } else {
        for (int a = 0; a < axes; ++a) {
            auto slab = [&](int s, int e, bool load) {
                Update u{};
                u.load = load;
                for (int o = 0; o < 3; ++o) {
                    u.start[o] = inferiorLastUpdateIndex[o] - BUFFER_SIZE;
                    u.end[o] = superiorLastUpdateIndex[o] + BUFFER_SIZE;
                }
                u.start[a] = s;
                u.end[a] = e;
                scheduleUpdate(u);
            };
            const int newInf = toInferiorGridCoordinate(pose(a, positionColumn), sensorMaxRange);
            if (std::abs(newInf - inferiorLastUpdateIndex[a]) >= 2) {
                if (newInf < inferiorLastUpdateIndex[a]) slab(newInf - BUFFER_SIZE, inferiorLastUpdateIndex[a] - BUFFER_SIZE - 1, true);
                if (newInf > inferiorLastUpdateIndex[a]) slab(inferiorLastUpdateIndex[a] - BUFFER_SIZE, newInf - BUFFER_SIZE - 1, false);
                inferiorLastUpdateIndex[a] = newInf;
            }
            const int newSup = toSuperiorGridCoordinate(pose(a, positionColumn), sensorMaxRange);
            if (std::abs(newSup - superiorLastUpdateIndex[a]) >= 2) {
                if (newSup < superiorLastUpdateIndex[a]) slab(newSup + BUFFER_SIZE + 1, superiorLastUpdateIndex[a] + BUFFER_SIZE, false);
                if (newSup > superiorLastUpdateIndex[a]) slab(superiorLastUpdateIndex[a] + BUFFER_SIZE + 1, newSup + BUFFER_SIZE, true);
                superiorLastUpdateIndex[a] = newSup;
            }
        }
    }
    if (!isOnline && !appliedUpdates.empty()) {  // (isOnline: the update thread rebuilds once its queue has drained)
        // icp.setMap(localPointCloud) of loadCells / unloadCells (Map.cpp:110-112,177-179), once for all slabs
        std::lock_guard<std::mutex> l1(localPointCloudLock);
        std::lock_guard<std::mutex> l2(icpMapLock);
        ICPSequence::check(icp.context(), b200icp_map_commit(icp.context()));
        localPointCloudEmpty.store(localPointCloud.getNbPoints() == 0);
    }
}

DataPoints Map::getLocalPointCloud() {
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    return localPointCloud.download(false, is3D ? 3 : 2);
}

// Map.cpp:502-534
void Map::updateLocalPointCloud(DataPoints input, TransformationParameters pose, PostFilters postFilters, bool asynchronous) {
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    // asynchronous: this thread works on the scan snapshot and on a second index until the final setMap
    std::unique_ptr<ICPSequence::UpdateThreadScope> scope;
    struct EndUpdate {  // (also on the way out of an exception: the context must not stay in update mode)
        b200icp_ctx* ctx = nullptr;
        ~EndUpdate() {
            if (ctx && b200icp_map_update_in_progress(ctx)) b200icp_map_end_update(ctx);
        }
    } endUpdate;
    if (asynchronous) {
        scope.reset(new ICPSequence::UpdateThreadScope());
        ICPSequence::check(icp.context(), b200icp_map_begin_update(icp.context()));
        endUpdate.ctx = icp.context();
    }
    if (isLocalPointCloudEmpty()) {
        auto iter = mapperModuleVec.begin();
        (*iter)->inPlaceCreateMap(input, localPointCloud, pose);
        ++iter;
        for (; iter != mapperModuleVec.end(); ++iter) (*iter)->inPlaceUpdateMap(input, localPointCloud, pose);
    } else {
        for (const auto& module : mapperModuleVec) module->inPlaceUpdateMap(input, localPointCloud, pose);
    }
    {
        // post filters, then icp.setMap(localPointCloud) (Map.cpp:523-529).  The reference moves the
        // whole map to the sensor frame and back around the filters; SurfaceNormal is covariant under
        // that rigid motion, so it runs in the map frame directly.
        std::lock_guard<std::mutex> l2(icpMapLock);
        ICPSequence::check(icp.context(), b200icp_map_commit(icp.context()));
        if (postFilters.surfaceNormalKnn > 0)
            ICPSequence::check(icp.context(), b200icp_map_surface_normals(icp.context(), postFilters.surfaceNormalKnn));
        if (postFilters.cutAtThreshold) {
            int64_t removed = 0;
            ICPSequence::check(icp.context(), b200icp_map_cut_at_threshold(icp.context(), postFilters.cutThreshold,
                                                                           postFilters.cutUseLargerThan ? 1 : 0, &removed));
            if (removed > 0) ICPSequence::check(icp.context(), b200icp_map_commit(icp.context()));
        }
        if (asynchronous) {  // icp.setMap(localPointCloud): the new index goes live
            ICPSequence::check(icp.context(), b200icp_map_end_update(icp.context()));
            endUpdate.ctx = nullptr;
        }
    }
    localPointCloudEmpty.store(localPointCloud.getNbPoints() == 0);
    newLocalPointCloudAvailable = true;
}

bool Map::getNewLocalPointCloud(DataPoints& out) {  // Map.cpp:536-551
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    if (!newLocalPointCloudAvailable) return false;
    out = localPointCloud.download(false, is3D ? 3 : 2);
    newLocalPointCloudAvailable = false;
    return true;
}

DataPoints Map::getGlobalPointCloud() {  // Map.cpp:553-573
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    return localPointCloud.download(true, is3D ? 3 : 2);
}

void Map::setGlobalPointCloud(const DataPoints& newLocalPointCloud) {  // Map.cpp:575-588
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    {
        std::lock_guard<std::mutex> l2(icpMapLock);
        icp.setMap(newLocalPointCloud);
    }
    localPointCloudEmpty.store(newLocalPointCloud.getNbPoints() == 0);
    firstPoseUpdate.store(true);
}

std::vector<std::array<int, 7>> Map::lastUpdates() const {
    std::vector<std::array<int, 7>> out;
    for (const auto& u : appliedUpdates) out.push_back({u.start[0], u.end[0], u.start[1], u.end[1], u.start[2], u.end[2], u.load ? 1 : 0});
    return out;
}

}  // namespace norlab_icp_mapper_b200
