// Map.cpp -- see Map.h.  Reference: /root/reference/norlab_icp_mapper/Map.cpp.
#include "Map.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <limits>
#include <unordered_map>

namespace norlab_icp_mapper_b200 {

Map::Map(bool is3D_, bool isOnline_, ICPSequence& icp_, std::mutex& icpMapLock_, std::unique_ptr<CellManager> cellManager_)
    : is3D(is3D_), isOnline(isOnline_), icp(icp_), icpMapLock(icpMapLock_), localPointCloud(icp_), cellManager(std::move(cellManager_)) {
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
    if (cellManager) {
        if (u.load) loadSpilledCells(u);
        else spillUnloaded();
    }
    if (changed > 0) newLocalPointCloudAvailable = true;
    appliedUpdates.push_back(u);
}

// unloadCells, second half (Map.cpp:196-230): the points that just left the local cloud are grouped by 20 m cell and handed to
// the CellManager; they leave device memory.
void Map::spillUnloaded() {
    b200icp_ctx* ctx = icp.context();
    const int dim = is3D ? 3 : 2, rows = dim + 1;
    int64_t n = 0;
    ICPSequence::check(ctx, b200icp_map_evict_parked(ctx, nullptr, nullptr, nullptr, nullptr, 0, &n));
    if (n == 0) return;
    DataPoints chunk;
    chunk.dim = dim;
    chunk.features.resize((size_t)n * rows);
    const bool hn = b200icp_map_has_normals(ctx) != 0, hp = b200icp_map_has_prob(ctx) != 0;
    const int xr = b200icp_map_extra_rows(ctx);
    if (hn) chunk.normals.resize((size_t)n * dim);
    if (hp) chunk.probabilityDynamic.resize((size_t)n);
    if (xr) {
        chunk.descriptors.resize((size_t)n * xr);
        chunk.descriptorLabels = icp.mapLabels;
    }
    ICPSequence::check(ctx, b200icp_map_evict_parked(ctx, chunk.features.data(), hn ? chunk.normals.data() : nullptr,
                                                     hp ? chunk.probabilityDynamic.data() : nullptr, xr ? chunk.descriptors.data() : nullptr, n, &n));
    std::unordered_map<std::string, DataPoints> cells;
    for (int64_t i = 0; i < n; ++i) {
        const int row = (int)std::floor(chunk.features[i * rows + 0] / CELL_SIZE), column = (int)std::floor(chunk.features[i * rows + 1] / CELL_SIZE);
        const int aisle = is3D ? (int)std::floor(chunk.features[i * rows + 2] / CELL_SIZE) : 0;
        const std::string id = std::to_string(row) + "_" + std::to_string(column) + "_" + std::to_string(aisle);
        DataPoints& c = cells[id];
        if (c.features.empty()) {
            c.dim = dim;
            c.descriptorLabels = chunk.descriptorLabels;
        }
        c.features.insert(c.features.end(), chunk.features.begin() + i * rows, chunk.features.begin() + (i + 1) * rows);
        if (hn) c.normals.insert(c.normals.end(), chunk.normals.begin() + i * dim, chunk.normals.begin() + (i + 1) * dim);
        if (hp) c.probabilityDynamic.push_back(chunk.probabilityDynamic[i]);
        if (xr) c.descriptors.insert(c.descriptors.end(), chunk.descriptors.begin() + i * xr, chunk.descriptors.begin() + (i + 1) * xr);
        loadedCellIds.erase(id);
    }
    std::lock_guard<std::mutex> lock(cellManagerLock);
    for (auto& kv : cells) {
        // a cell can be cut by the window's metric edge (Map.cpp:161-174 tests coordinates, the ids are floors): what an earlier
        // unload already stored under this id stays
        DataPoints merged = cellManager->retrieveCell(kv.first);
        if (merged.getNbPoints() > 0) {
            merged.concatenate(kv.second);
            cellManager->saveCell(kv.first, merged);
        } else {
            cellManager->saveCell(kv.first, kv.second);
        }
    }
}

// loadCells, first half (Map.cpp:79-99): cells of the slab that are stored and not loaded yet come back into the local cloud
void Map::loadSpilledCells(const Update& u) {
    b200icp_ctx* ctx = icp.context();
    const int dim = is3D ? 3 : 2;
    std::vector<std::string> ids;
    {
        std::lock_guard<std::mutex> lock(cellManagerLock);
        ids = cellManager->getAllCellIds();
    }
    for (const std::string& id : ids) {
        int r = 0, c = 0, a = 0;
        if (std::sscanf(id.c_str(), "%d_%d_%d", &r, &c, &a) != 3) continue;
        const int a0 = is3D ? u.start[2] : 0, a1 = is3D ? u.end[2] : 0;
        if (r < u.start[0] || r > u.end[0] || c < u.start[1] || c > u.end[1] || a < a0 || a > a1) continue;
        if (loadedCellIds.count(id)) continue;
        DataPoints cell;
        {
            std::lock_guard<std::mutex> lock(cellManagerLock);
            cell = cellManager->retrieveCell(id);
            cellManager->saveCell(id, DataPoints());  // (it lives on the device again; an empty cloud keeps the id known, as unloading rewrites it)
        }
        loadedCellIds.insert(id);
        if (cell.getNbPoints() == 0) continue;
        cell.dim = dim;
        // DataPoints::concatenate's rule for the named descriptors: both sides keep what they have in common
        const bool mapEmpty = localPointCloud.getNbPointsGlobal() == 0;
        if (mapEmpty) {
            icp.mapLabels = cell.descriptorLabels;
        } else {
            const Labels common = commonLabels(icp.mapLabels, cell.descriptorLabels);
            if (!(common == icp.mapLabels)) {
                const std::vector<int32_t> sel = rowsOf(icp.mapLabels, common);
                ICPSequence::check(ctx, b200icp_map_select_extra(ctx, sel.data(), (int32_t)sel.size()));
                icp.mapLabels = common;
            }
            if (!(common == cell.descriptorLabels)) cell.selectDescriptors(common);
        }
        int64_t added = 0;
        ICPSequence::check(ctx, b200icp_map_append_cloud(ctx, cell.features.data(), dim + 1, cell.getNbPoints(), cell.normals.empty() ? nullptr : cell.normals.data(),
                                                         cell.probabilityDynamic.empty() ? nullptr : cell.probabilityDynamic.data(),
                                                         cell.descriptors.empty() ? nullptr : cell.descriptors.data(), cell.getDescriptorRows(), &added));
        newLocalPointCloudAvailable = true;
    }
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
        if (cellManager) {
            std::lock_guard<std::mutex> l3(cellManagerLock);
            cellManager->clearAllCells();
            loadedCellIds.clear();
        }
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
    DataPoints global = localPointCloud.download(true, is3D ? 3 : 2);
    if (cellManager) {  // ... + every stored cell that is not loaded
        std::lock_guard<std::mutex> l2(cellManagerLock);
        for (const std::string& id : cellManager->getAllCellIds())
            if (!loadedCellIds.count(id)) {
                const DataPoints cell = cellManager->retrieveCell(id);
                if (cell.getNbPoints() > 0) global.concatenate(cell);
            }
    }
    return global;
}

void Map::setGlobalPointCloud(const DataPoints& newLocalPointCloud) {  // Map.cpp:575-588
    std::lock_guard<std::mutex> lock(localPointCloudLock);
    {
        std::lock_guard<std::mutex> l2(icpMapLock);
        icp.setMap(newLocalPointCloud);
    }
    if (cellManager) {
        std::lock_guard<std::mutex> l3(cellManagerLock);
        cellManager->clearAllCells();
        loadedCellIds.clear();
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
