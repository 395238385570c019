// build_map_from_scans_and_trajectory -- the reference's example pipeline
// (/root/reference/examples/build_map_from_scans_and_trajectory.cpp:174-239) on the B200 path: read <dataPath>/trajectory.csv
// and <dataPath>/scans/*.vtk, run every scan through Mapper::applyInputFilters + Mapper::processInput, write
// <dataPath>/map.vtk.  The YAML file of the reference is out of scope: the configuration below IS the shipped
// examples/config.yaml, spelled out on the MapperConfig struct; `--icp point_to_plane` swaps the example's Identity error
// minimiser (the shipped config does no registration at all) for the chain of docs/MapperConfiguration.md:172-189.
//
//   build_map_from_scans_and_trajectory <dataPath> [config.yaml] [--icp identity|point_to_plane] [--io-only] [--binary] [--host-module]
//
// With a YAML file as second argument -- the reference's own command line -- the configuration is read from it
// (host/YamlConfig.h; /root/reference/examples/config.yaml loads unmodified); without one the same configuration is built in code.
//
// --host-module swaps the device OctreeMapperModule for a module written against the REFERENCE's plugin signature
// (MapperModules/MapperModule.h:20-29: host DataPoints in, host DataPoints out), run through HostMapperModuleAdapter: a
// voxel-grid thinning that keeps the first point of every 15 cm voxel -- a stand-in for a third-party module.
//
// --io-only concatenates the scans at their trajectory poses on the host and writes the result: a check of the readers /
// writer that needs no GPU.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <iostream>

#include "../norlab_icp_mapper_b200/host/IO.h"
#include "../norlab_icp_mapper_b200/host/Mapper.h"
#include "../norlab_icp_mapper_b200/host/YamlConfig.h"

namespace fs = std::filesystem;
using namespace norlab_icp_mapper_b200;

static b200icp_filter box(float x0, float x1, float y0, float y1, float z0, float z1) {
    b200icp_filter f{};
    f.kind = B200ICP_FILTER_BOUNDING_BOX;
    f.lo[0] = x0; f.hi[0] = x1; f.lo[1] = y0; f.hi[1] = y1; f.lo[2] = z0; f.hi[2] = z1;
    f.remove_inside = 1;
    return f;
}

// A third-party style module: reference signature, host clouds only, knows nothing about the device.
class VoxelFirstHostModule : public HostMapperModule {
    float voxel;

   public:
    explicit VoxelFirstHostModule(float v) : voxel(v) {}
    DataPoints createMap(const DataPoints& input, const TransformationParameters& pose) override {
        DataPoints out;
        out.dim = input.dim;
        inPlaceUpdateMap(input, out, pose);
        return out;
    }
    void inPlaceCreateMap(DataPoints& input, const TransformationParameters& pose) override { input = createMap(input, pose); }
    DataPoints updateMap(const DataPoints& input, const DataPoints& map, const TransformationParameters& pose) override {
        DataPoints out = map;
        inPlaceUpdateMap(input, out, pose);
        return out;
    }
    void inPlaceUpdateMap(const DataPoints& input, DataPoints& map, const TransformationParameters&) override {
        map.concatenate(input);  // (descriptors both clouds carry survive, like PM::DataPoints::concatenate)
        const int rows = map.dim + 1, xr = map.getDescriptorRows();
        const int64_t n = map.getNbPoints();
        std::vector<std::array<int64_t, 3>> seen;
        std::vector<int64_t> keep;
        {
            std::vector<std::pair<std::array<int64_t, 3>, int64_t>> keyed((size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                std::array<int64_t, 3> k{0, 0, 0};
                for (int d = 0; d < map.dim; ++d) k[d] = (int64_t)std::floor(map.features[i * rows + d] / voxel);
                keyed[i] = {k, i};
            }
            std::sort(keyed.begin(), keyed.end());
            for (size_t j = 0; j < keyed.size(); ++j)
                if (j == 0 || keyed[j].first != keyed[j - 1].first) keep.push_back(keyed[j].second);
            std::sort(keep.begin(), keep.end());
        }
        DataPoints out;
        out.dim = map.dim;
        out.descriptorLabels = map.descriptorLabels;
        for (int64_t i : keep) {
            out.features.insert(out.features.end(), map.features.begin() + i * rows, map.features.begin() + (i + 1) * rows);
            if (!map.normals.empty()) out.normals.insert(out.normals.end(), map.normals.begin() + i * map.dim, map.normals.begin() + (i + 1) * map.dim);
            if (!map.probabilityDynamic.empty()) out.probabilityDynamic.push_back(map.probabilityDynamic[i]);
            if (xr > 0) out.descriptors.insert(out.descriptors.end(), map.descriptors.begin() + i * xr, map.descriptors.begin() + (i + 1) * xr);
        }
        map = out;
    }
};

int main(int argc, char* argv[]) {
    if (argc < 2) {
        std::cerr << "Please provide a dataPath as an argument." << std::endl;
        return -1;
    }
    const fs::path dataPath = argv[1];
    bool ioOnly = false, binary = false, pointToPlane = false, hostModule = false;
    std::string configYaml;
    for (int i = 2; i < argc; ++i) {
        if (argv[i][0] != '-') {
            configYaml = argv[i];
            continue;
        }
        if (!std::strcmp(argv[i], "--io-only")) ioOnly = true;
        else if (!std::strcmp(argv[i], "--binary")) binary = true;
        else if (!std::strcmp(argv[i], "--host-module")) hostModule = true;
        else if (!std::strcmp(argv[i], "--icp") && i + 1 < argc) pointToPlane = !std::strcmp(argv[++i], "point_to_plane");
    }
    try {
        const auto poses = io::loadTrajectoryCSV((dataPath / "trajectory.csv").string());
        std::vector<std::string> scans;
        for (const auto& e : fs::directory_iterator(dataPath / "scans"))
            if (e.path().extension() == ".vtk") scans.push_back(e.path().string());
        std::sort(scans.begin(), scans.end());
        if (poses.size() != scans.size()) throw std::runtime_error("trajectory.csv and scans/ disagree on the number of scans");
        const fs::path outputPath = dataPath / "map.vtk";

        if (ioOnly) {
            DataPoints all;
            for (size_t i = 0; i < scans.size(); ++i) {
                const DataPoints c = io::loadVTK(scans[i]);
                for (int64_t p = 0; p < c.getNbPoints(); ++p) {
                    const float* q = &c.features[(size_t)p * 4];
                    for (int r = 0; r < 3; ++r)
                        all.features.push_back(poses[i].T(r, 0) * q[0] + poses[i].T(r, 1) * q[1] + poses[i].T(r, 2) * q[2] + poses[i].T(r, 3));
                    all.features.push_back(1.f);
                }
            }
            io::saveVTK(all, outputPath.string(), binary);
            std::cout << "Output saved to " << outputPath << " (" << all.getNbPoints() << " points, I/O only)" << std::endl;
            return 0;
        }

        // examples/config.yaml
        MapperConfig cfg;
        b200icp_config_default(&cfg.icp, 3);
        cfg.icp.knn = 6;                                      // icp.matcher KDTreeMatcher{knn 6, maxDist 2.0, epsilon 1}
        cfg.icp.max_dist = 2.0f;
        cfg.icp.epsilon = 1.0f;                               // (accepted, the search is exact)
        cfg.icp.n_outlier = 0;                                // no outlierFilters in the YAML
        cfg.icp.minimizer = pointToPlane ? B200ICP_MIN_POINT_TO_PLANE : B200ICP_MIN_IDENTITY;
        cfg.icp.max_iteration_count = 10;                     // CounterTransformationChecker
        cfg.icp.use_differential = 0;
        cfg.inputFilters = {box(-1.5f, 0.5f, -1.f, 1.f, -1.f, 0.5f), box(-6.f, -1.5f, -2.5f, 2.5f, -1.f, 1.f)};
        cfg.addProbabilityDynamic = true;                     // AddDescriptor{probabilityDynamic, 1, [0.6]}
        cfg.probabilityDynamicValue = 0.6f;
        cfg.post.surfaceNormalKnn = 10;
        cfg.post.cutAtThreshold = true;
        cfg.post.cutUseLargerThan = true;
        cfg.post.cutThreshold = 0.65f;
        cfg.mapUpdateCondition = "delay";
        cfg.mapUpdateValue = 0.05f;
        cfg.sensorMaxRange = 200.f;
        cfg.mapperModules = {
            {"DynamicPointsMapperModule", {{"thresholdDynamic", "0.9"}, {"alpha", "0.8"}, {"beta", "0.99"}, {"beamHalfAngle", "0.01"},
                                           {"epsilonA", "0.01"}, {"epsilonD", "0.01"}}},
            {"OctreeMapperModule", {{"buildParallel", "1"}, {"maxSizeByNode", "0.15"}, {"samplingMethod", "1"}}}};

        if (!configYaml.empty()) {
            cfg = loadYamlConfig(configYaml, /*is3D=*/true);
            if (pointToPlane) cfg.icp.minimizer = B200ICP_MIN_POINT_TO_PLANE;
        }
        if (hostModule) {
            cfg.mapperModules.pop_back();
            cfg.extraModules.push_back(std::make_shared<HostMapperModuleAdapter>(std::make_shared<VoxelFirstHostModule>(0.15f)));
        }

        Mapper mapper(cfg, /*is3D=*/true, /*isOnline=*/false, /*isMapping=*/true, /*saveMapCellsOnHardDrive=*/false);
        mapper.setDeviceResidentInput(true);  // the filtered scan stays in the device slot between applyInputFilters and processInput
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < scans.size(); ++i) {
            DataPoints inputCloud = io::loadVTK(scans[i]);
            mapper.applyInputFilters(inputCloud);
            mapper.processInput(inputCloud, poses[i].T, 1e-9 * (double)poses[i].stamp_ns);
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const DataPoints map = mapper.getMap();
        io::saveVTK(map, outputPath.string(), binary);
        std::cout << "Output saved to " << outputPath << " (" << map.getNbPoints() << " points, " << scans.size() << " scans, " << ms
                  << " ms incl. reading the scans)" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
