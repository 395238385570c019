// build_map_from_scans_and_trajectory -- the reference's example pipeline
// (/root/reference/examples/build_map_from_scans_and_trajectory.cpp:174-239) on the B200 path: read <dataPath>/trajectory.csv
// and <dataPath>/scans/*.vtk, run every scan through Mapper::applyInputFilters + Mapper::processInput, write
// <dataPath>/map.vtk.  The YAML file of the reference is out of scope: the configuration below IS the shipped
// examples/config.yaml, spelled out on the MapperConfig struct; `--icp point_to_plane` swaps the example's Identity error
// minimiser (the shipped config does no registration at all) for the chain of docs/MapperConfiguration.md:172-189.
//
//   build_map_from_scans_and_trajectory <dataPath> [--icp identity|point_to_plane] [--io-only] [--binary]
//
// --io-only concatenates the scans at their trajectory poses on the host and writes the result: a check of the readers /
// writer that needs no GPU.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <iostream>

#include "../norlab_icp_mapper_b200/host/IO.h"
#include "../norlab_icp_mapper_b200/host/Mapper.h"

namespace fs = std::filesystem;
using namespace norlab_icp_mapper_b200;

static b200icp_filter box(float x0, float x1, float y0, float y1, float z0, float z1) {
    b200icp_filter f{};
    f.kind = B200ICP_FILTER_BOUNDING_BOX;
    f.lo[0] = x0; f.hi[0] = x1; f.lo[1] = y0; f.hi[1] = y1; f.lo[2] = z0; f.hi[2] = z1;
    f.remove_inside = 1;
    return f;
}

int main(int argc, char* argv[]) {
    if (argc < 2) {
        std::cerr << "Please provide a dataPath as an argument." << std::endl;
        return -1;
    }
    const fs::path dataPath = argv[1];
    bool ioOnly = false, binary = false, pointToPlane = false;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--io-only")) ioOnly = true;
        else if (!std::strcmp(argv[i], "--binary")) binary = true;
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

        Mapper mapper(cfg, /*is3D=*/true, /*isOnline=*/false, /*isMapping=*/true, /*saveMapCellsOnHardDrive=*/false);
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
