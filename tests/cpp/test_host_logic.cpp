// CPU-only checks of the host mirror's data structures (no GPU, no libb200icp calls): DataPoints descriptors and
// DataPoints::concatenate's common-descriptor rule, the CellManager implementations behind the reference's interface
// (CellManager.h:15-18, RAMCellManager.cpp, HardDriveCellManager.cpp) including the VTK round trip of named descriptors, and
// the YAML reader's tree.  Built and run by tests/test_host_logic.py; prints "ok" lines, exits non-zero on the first failure.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "../../norlab_icp_mapper_b200/host/CellManager.h"
#include "../../norlab_icp_mapper_b200/host/YamlConfig.h"

using namespace norlab_icp_mapper_b200;

#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            std::exit(1);                                                           \
        }                                                                           \
    } while (0)

static DataPoints cloud(int n, float x0, bool withNormals, bool withProb, const Labels& labels) {
    DataPoints c;
    c.dim = 3;
    for (int i = 0; i < n; ++i) {
        c.features.insert(c.features.end(), {x0 + (float)i, 0.5f * (float)i, -1.f, 1.f});
        if (withNormals) c.normals.insert(c.normals.end(), {0.f, 0.f, 1.f});
        if (withProb) c.probabilityDynamic.push_back(0.25f + 0.01f * (float)i);
    }
    c.descriptorLabels = labels;
    const int rows = labelRows(labels);
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < rows; ++r) c.descriptors.push_back(x0 + 100.f * (float)r + (float)i);
    return c;
}

int main(int argc, char** argv) {
    const std::string folder = argc > 1 ? argv[1] : "/tmp/";

    // ---- descriptors
    DataPoints a = cloud(4, 0.f, true, true, {{"intensity", 1}, {"observationDirections", 3}, {"t", 1}});
    CHECK(a.getNbPoints() == 4 && a.getDescriptorRows() == 5);
    CHECK(a.descriptorExists("normals") && a.descriptorExists("t") && !a.descriptorExists("ring"));
    CHECK(labelStartingRow(a.descriptorLabels, "observationDirections") == 1 && labelStartingRow(a.descriptorLabels, "t") == 4);
    const std::vector<float> t = a.getDescriptorCopyByName("t");
    CHECK(t.size() == 4 && t[2] == 0.f + 100.f * 4 + 2.f);
    a.addDescriptor("ring", 1, {7.f, 8.f, 9.f, 10.f});
    CHECK(a.getDescriptorRows() == 6 && a.getDescriptorCopyByName("ring")[3] == 10.f && a.getDescriptorCopyByName("intensity")[1] == 1.f);
    a.removeDescriptor("observationDirections");
    CHECK(a.getDescriptorRows() == 3 && a.descriptorLabels[1].text == "t" && a.getDescriptorCopyByName("t")[2] == 402.f);
    bool threw = false;
    try {
        a.getDescriptorCopyByName("nope");
    } catch (const InvalidField&) {
        threw = true;
    }
    CHECK(threw);
    std::puts("ok descriptors");

    // ---- concatenate: only descriptors BOTH clouds carry survive, in the first cloud's order
    DataPoints m = cloud(3, 0.f, true, true, {{"intensity", 1}, {"t", 1}, {"ring", 1}});
    DataPoints s = cloud(2, 50.f, true, false, {{"ring", 1}, {"intensity", 1}});
    m.concatenate(s);
    CHECK(m.getNbPoints() == 5 && !m.normals.empty() && m.normals.size() == 15);
    CHECK(m.probabilityDynamic.empty());  // the second cloud has none
    CHECK(m.descriptorLabels.size() == 2 && m.descriptorLabels[0].text == "intensity" && m.descriptorLabels[1].text == "ring");
    CHECK(m.descriptors.size() == 10);
    CHECK(m.descriptors[0 * 2 + 0] == 0.f && m.descriptors[0 * 2 + 1] == 200.f);          // first cloud, point 0: intensity row 0, ring row 2
    CHECK(m.descriptors[3 * 2 + 0] == 150.f && m.descriptors[3 * 2 + 1] == 50.f);        // second cloud, point 0: its intensity is row 1, ring row 0
    DataPoints empty;
    empty.concatenate(s);  // an empty cloud takes everything
    CHECK(empty.getNbPoints() == 2 && empty.descriptorLabels.size() == 2 && empty.descriptorLabels[0].text == "ring");
    threw = false;
    try {
        DataPoints bad = cloud(1, 0.f, false, false, {{"intensity", 2}});
        DataPoints good = cloud(1, 0.f, false, false, {{"intensity", 1}});
        good.concatenate(bad);
    } catch (const InvalidField&) {
        threw = true;
    }
    CHECK(threw);  // same name, different dimension
    CHECK((rowsOf({{"a", 1}, {"b", 3}, {"c", 1}}, {{"c", 1}, {"b", 3}}) == std::vector<int32_t>{4, 1, 2, 3}));
    std::puts("ok concatenate");

    // ---- cell managers behind the reference's interface
    for (int kind = 0; kind < 2; ++kind) {
        std::unique_ptr<CellManager> cm;
        if (kind == 0) cm.reset(new RAMCellManager());
        else cm.reset(new HardDriveCellManager(3, folder));
        CHECK(cm->getAllCellIds().empty() && cm->retrieveCell("0_0_0").getNbPoints() == 0);
        const DataPoints c1 = cloud(5, 1.f, true, true, {{"intensity", 1}, {"t", 1}});
        const DataPoints c2 = cloud(2, 9.f, true, true, {{"intensity", 1}, {"t", 1}});
        cm->saveCell("3_-2_0", c1);
        cm->saveCell("-1_7_1", c2);
        CHECK(cm->getAllCellIds().size() == 2);
        const DataPoints r1 = cm->retrieveCell("3_-2_0");
        CHECK(r1.getNbPoints() == 5 && r1.features == c1.features && r1.normals == c1.normals && r1.probabilityDynamic == c1.probabilityDynamic);
        CHECK(r1.descriptorLabels.size() == 2 && r1.getDescriptorCopyByName("t") == c1.getDescriptorCopyByName("t") &&
              r1.getDescriptorCopyByName("intensity") == c1.getDescriptorCopyByName("intensity"));
        cm->saveCell("3_-2_0", c2);  // overwrite
        CHECK(cm->retrieveCell("3_-2_0").getNbPoints() == 2 && cm->getAllCellIds().size() == 2);
        if (kind == 1) {
            std::ifstream f(folder + (folder.back() == '/' ? "" : "/") + "cell_-1_7_1.vtk");
            CHECK(f.is_open());  // HardDriveCellManager.h: CELL_FOLDER + "cell_" + id + ".vtk"
        }
        cm->clearAllCells();
        CHECK(cm->getAllCellIds().empty() && cm->retrieveCell("3_-2_0").getNbPoints() == 0);
        if (kind == 1) {
            std::ifstream f(folder + (folder.back() == '/' ? "" : "/") + "cell_-1_7_1.vtk");
            CHECK(!f.is_open());  // clearAllCells removes the files (HardDriveCellManager.cpp:30-36)
        }
    }
    std::puts("ok cell managers");

    // ---- the YAML subset reader: tree shape
    std::istringstream y(
        "a:\n  - X:\n      p: 1\n      q: [1, 2.5]  # comment\n  - Y\nb: {k: v, n: 3}\nc:\n  d:\n    E:\n");
    const yaml::Node root = yaml::parse(y);
    CHECK(root.kind == yaml::Node::Map && root.map.size() == 3);
    const yaml::Node* an = root.find("a");
    CHECK(an && an->kind == yaml::Node::List && an->list.size() == 2);
    CHECK(an->list[0].kind == yaml::Node::Map && an->list[0].map[0].first == "X" && an->list[0].map[0].second.find("q")->list.size() == 2);
    CHECK(an->list[1].kind == yaml::Node::Scalar && an->list[1].scalar == "Y");
    CHECK(root.find("b")->find("n")->scalar == "3");
    CHECK(root.find("c")->find("d")->find("E")->isNull());
    std::puts("ok yaml tree");
    return 0;
}
