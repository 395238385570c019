// CellManager.h -- the reference's cell-storage seam (CellManager.h:15-18) with its two implementations (RAMCellManager.cpp,
// HardDriveCellManager.cpp) over this path's DataPoints.  On the device-resident map the cells a moving window leaves normally STAY in
// HBM (a `loaded` flag, nothing copied); a Map given one of these managers uses it as a spill tier instead: the unloaded points are
// moved out of device memory into per-cell clouds keyed "row_col_aisle" and concatenated back when the window returns
// (host/Map.cpp, b200icp_map_evict_parked / b200icp_map_append_cloud).  180 GB of HBM hold ~5 billion map points, so this matters only
// for maps beyond that, or to exchange cells with code written against the reference's interface.
#pragma once
#include <cstdio>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "DataPoints.h"
#include "IO.h"

namespace norlab_icp_mapper_b200 {

class CellManager {
   public:
    virtual ~CellManager() = default;
    virtual std::vector<std::string> getAllCellIds() const = 0;
    virtual void saveCell(const std::string& cellId, const DataPoints& cell) = 0;
    virtual DataPoints retrieveCell(const std::string& cellId) const = 0;  // an empty cloud when the cell is unknown
    virtual void clearAllCells() = 0;
};

// RAMCellManager.cpp:3-31
class RAMCellManager : public CellManager {
    std::unordered_map<std::string, DataPoints> cells;

   public:
    std::vector<std::string> getAllCellIds() const override {
        std::vector<std::string> ids;
        for (const auto& kv : cells) ids.push_back(kv.first);
        return ids;
    }
    void saveCell(const std::string& cellId, const DataPoints& cell) override { cells[cellId] = cell; }
    DataPoints retrieveCell(const std::string& cellId) const override {
        const auto it = cells.find(cellId);
        return it == cells.end() ? DataPoints() : it->second;
    }
    void clearAllCells() override { cells.clear(); }
};

// HardDriveCellManager.cpp:1-37: one legacy-VTK file per cell, "<folder>cell_<id>.vtk" (the reference's folder is /tmp/)
class HardDriveCellManager : public CellManager {
    std::string folder, prefix = "cell_", suffix = ".vtk";
    std::unordered_set<std::string> cellIds;
    int dim;
    std::string path(const std::string& id) const { return folder + prefix + id + suffix; }

   public:
    explicit HardDriveCellManager(int dim_ = 3, const std::string& cellFolder = "/tmp/") : folder(cellFolder), dim(dim_) {
        if (!folder.empty() && folder.back() != '/') folder += '/';
    }
    ~HardDriveCellManager() override { clearAllCells(); }
    std::vector<std::string> getAllCellIds() const override { return std::vector<std::string>(cellIds.begin(), cellIds.end()); }
    void saveCell(const std::string& cellId, const DataPoints& cell) override {
        io::saveVTK(cell, path(cellId), /*binary=*/true);
        cellIds.insert(cellId);
    }
    DataPoints retrieveCell(const std::string& cellId) const override {
        if (cellIds.find(cellId) == cellIds.end()) return DataPoints();
        return io::loadVTK(path(cellId), dim);
    }
    void clearAllCells() override {
        for (const auto& id : cellIds) std::remove(path(id).c_str());
        cellIds.clear();
    }
};

}  // namespace norlab_icp_mapper_b200
