// IO.h -- the disk formats the reference's example pipeline exchanges with the world:
//   * legacy VTK POLYDATA clouds as libpointmatcher writes them (`DP::load` / `DP::save`,
//     /root/reference/examples/build_map_from_scans_and_trajectory.cpp:228,234; the bundled
//     examples/data/scans/*.vtk: "# vtk DataFile Version 3.0" / title / ASCII / DATASET POLYDATA / POINTS n float /
//     VERTICES n 2n / POINT_DATA n / SCALARS <name> float + LOOKUP_TABLE default ...), ASCII and BINARY;
//   * the trajectory CSV of the example (same file, :15-172): a header row naming the columns, of which
//     header.stamp.{sec,nanosec} and pose.pose.{position.{x,y,z},orientation.{x,y,z,w}} are used.
// Header-only, host-only, no dependencies.  Every point-data array becomes a descriptor of the cloud (SCALARS with their
// component count, VECTORS / NORMALS with three rows); COLOR_SCALARS and FIELD arrays are skipped.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "DataPoints.h"

namespace norlab_icp_mapper_b200 {
namespace io {

namespace detail {
inline void swap4(void* p) {
    unsigned char* b = static_cast<unsigned char*>(p);
    std::swap(b[0], b[3]);
    std::swap(b[1], b[2]);
}
inline void swap8(void* p) {
    unsigned char* b = static_cast<unsigned char*>(p);
    std::swap(b[0], b[7]);
    std::swap(b[1], b[6]);
    std::swap(b[2], b[5]);
    std::swap(b[3], b[4]);
}
// n values of `type` ("float" / "double" / integer types are skipped as floats of their width) -> float
inline std::vector<float> readValues(std::istream& in, bool binary, const std::string& type, size_t n) {
    std::vector<float> out(n);
    if (!binary) {
        for (size_t i = 0; i < n; ++i) {
            double v;
            if (!(in >> v)) throw std::runtime_error("VTK: unexpected end of data");
            out[i] = (float)v;
        }
        return out;
    }
    in.get();  // the single newline after the header line
    if (type == "double") {
        std::vector<double> tmp(n);
        in.read(reinterpret_cast<char*>(tmp.data()), (std::streamsize)(n * 8));
        for (size_t i = 0; i < n; ++i) {
            swap8(&tmp[i]);  // legacy VTK binary is big-endian
            out[i] = (float)tmp[i];
        }
    } else if (type == "float") {
        in.read(reinterpret_cast<char*>(out.data()), (std::streamsize)(n * 4));
        for (size_t i = 0; i < n; ++i) swap4(&out[i]);
    } else if (type == "int" || type == "unsigned_int") {
        std::vector<int32_t> tmp(n);
        in.read(reinterpret_cast<char*>(tmp.data()), (std::streamsize)(n * 4));
        for (size_t i = 0; i < n; ++i) {
            swap4(&tmp[i]);
            out[i] = (float)tmp[i];
        }
    } else {
        throw std::runtime_error("VTK: unsupported binary data type " + type);
    }
    if (!in) throw std::runtime_error("VTK: unexpected end of binary data");
    return out;
}
}  // namespace detail

//! DP::load for legacy VTK POLYDATA (3-D points; `dim` = 2 drops z).
inline DataPoints loadVTK(const std::string& path, int dim = 3) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Could not open " + path);
    std::string line;
    std::getline(in, line);
    if (line.rfind("# vtk DataFile", 0) != 0) throw std::runtime_error(path + ": not a legacy VTK file");
    std::getline(in, line);  // title
    std::getline(in, line);
    const bool binary = line.rfind("BINARY", 0) == 0;
    if (!binary && line.rfind("ASCII", 0) != 0) throw std::runtime_error(path + ": expected ASCII or BINARY");
    DataPoints out;
    out.dim = dim;
    size_t n = 0;
    std::string tok;
    while (in >> tok) {
        if (tok == "DATASET") {
            in >> tok;
            if (tok != "POLYDATA" && tok != "UNSTRUCTURED_GRID") throw std::runtime_error(path + ": unsupported DATASET " + tok);
        } else if (tok == "POINTS") {
            std::string type;
            in >> n >> type;
            const std::vector<float> xyz = detail::readValues(in, binary, type, n * 3);
            out.features.resize(n * (size_t)(dim + 1));
            for (size_t i = 0; i < n; ++i) {
                for (int d = 0; d < dim; ++d) out.features[i * (dim + 1) + d] = xyz[i * 3 + d];
                out.features[i * (dim + 1) + dim] = 1.f;
            }
        } else if (tok == "VERTICES" || tok == "LINES" || tok == "POLYGONS" || tok == "CELLS") {
            size_t cells, ints;
            in >> cells >> ints;
            detail::readValues(in, binary, "int", ints);
        } else if (tok == "CELL_TYPES") {
            size_t cells;
            in >> cells;
            detail::readValues(in, binary, "int", cells);
        } else if (tok == "POINT_DATA") {
            size_t m;
            in >> m;
            if (m != n) throw std::runtime_error(path + ": POINT_DATA size differs from POINTS");
        } else if (tok == "SCALARS") {
            std::string name, type;
            in >> name >> type;
            std::getline(in, line);  // optional numComp
            int comps = 1;
            {
                std::istringstream ls(line);
                int c;
                if (ls >> c) comps = c;
            }
            in >> tok;  // LOOKUP_TABLE
            if (tok != "LOOKUP_TABLE") throw std::runtime_error(path + ": SCALARS without LOOKUP_TABLE");
            in >> tok;  // table name
            const std::vector<float> v = detail::readValues(in, binary, type, n * (size_t)comps);
            out.addDescriptor(name, comps, v);  // (`probabilityDynamic` lands in its own member)
        } else if (tok == "NORMALS" || tok == "VECTORS" || tok == "COLOR_SCALARS") {
            std::string name, type = "float";
            size_t comps = 3;
            in >> name;
            if (tok == "COLOR_SCALARS") {
                in >> comps;  // ASCII floats / binary unsigned chars: not used on this path
                if (binary) throw std::runtime_error(path + ": binary COLOR_SCALARS are not supported");
            } else {
                in >> type;
            }
            const std::vector<float> v = detail::readValues(in, binary, type, n * comps);
            if (tok != "COLOR_SCALARS" && name == "normals") {
                out.normals.resize(n * (size_t)dim);
                for (size_t i = 0; i < n; ++i)
                    for (int d = 0; d < dim; ++d) out.normals[i * dim + d] = v[i * 3 + d];
            } else if (tok != "COLOR_SCALARS") {
                // a vector descriptor (observationDirections, eigVectors ...): dim rows, like libpointmatcher keeps them
                std::vector<float> d2(n * (size_t)dim);
                for (size_t i = 0; i < n; ++i)
                    for (int d = 0; d < dim; ++d) d2[i * dim + d] = v[i * 3 + d];
                out.addDescriptor(name, dim, d2);
            }
        } else if (tok == "FIELD") {
            std::string name;
            int arrays;
            in >> name >> arrays;
            for (int a = 0; a < arrays; ++a) {
                std::string aname, type;
                size_t comps, tuples;
                in >> aname >> comps >> tuples >> type;
                detail::readValues(in, binary, type, comps * tuples);
            }
        } else {
            throw std::runtime_error(path + ": unsupported VTK keyword " + tok);
        }
    }
    return out;
}

//! DP::save: the layout libpointmatcher writes (and loadVTK / ParaView read).
inline void saveVTK(const DataPoints& cloud, const std::string& path, bool binary = false) {
    if (cloud.onDevice) throw std::runtime_error("saveVTK: materialize the device-resident cloud first");
    std::ofstream out(path, std::ios::binary);
    if (!out.is_open()) throw std::runtime_error("Could not open " + path);
    const size_t n = (size_t)cloud.getNbPoints();
    const int dim = cloud.dim;
    out << "# vtk DataFile Version 3.0\nFile created by norlab_icp_mapper_b200\n" << (binary ? "BINARY" : "ASCII") << "\nDATASET POLYDATA\n";
    auto writeFloats = [&](const std::vector<float>& v, size_t per_line) {
        if (binary) {
            std::vector<float> be(v);
            for (float& f : be) detail::swap4(&f);
            out.write(reinterpret_cast<const char*>(be.data()), (std::streamsize)(be.size() * 4));
            out << "\n";
        } else {
            out.precision(9);
            for (size_t i = 0; i < v.size(); ++i) out << v[i] << ((i + 1) % per_line == 0 ? "\n" : " ");
        }
    };
    std::vector<float> xyz(n * 3, 0.f);
    for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d) xyz[i * 3 + d] = cloud.features[i * (dim + 1) + d];
    out << "POINTS " << n << " float\n";
    writeFloats(xyz, 3);
    out << "VERTICES " << n << " " << 2 * n << "\n";
    if (binary) {
        std::vector<int32_t> cells(2 * n);
        for (size_t i = 0; i < n; ++i) {
            cells[2 * i] = 1;
            cells[2 * i + 1] = (int32_t)i;
            detail::swap4(&cells[2 * i]);
            detail::swap4(&cells[2 * i + 1]);
        }
        out.write(reinterpret_cast<const char*>(cells.data()), (std::streamsize)(cells.size() * 4));
        out << "\n";
    } else {
        for (size_t i = 0; i < n; ++i) out << "1 " << i << "\n";
    }
    out << "POINT_DATA " << n << "\n";
    if (!cloud.normals.empty()) {
        std::vector<float> nr(n * 3, 0.f);
        for (size_t i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d) nr[i * 3 + d] = cloud.normals[i * dim + d];
        out << "NORMALS normals float\n";
        writeFloats(nr, 3);
    }
    if (!cloud.probabilityDynamic.empty()) {
        out << "SCALARS probabilityDynamic float\nLOOKUP_TABLE default\n";
        writeFloats(cloud.probabilityDynamic, 1);
    }
    {
        const int rows = cloud.getDescriptorRows();
        int r0 = 0;
        for (const Label& l : cloud.descriptorLabels) {
            const bool vec = dim == 3 && l.span == 3;  // three-row descriptors are written as VECTORS, the rest as SCALARS
            std::vector<float> v(n * (size_t)(vec ? 3 : l.span), 0.f);
            for (size_t i = 0; i < n; ++i)
                for (int c = 0; c < l.span; ++c) v[i * (vec ? 3 : l.span) + c] = cloud.descriptors[i * rows + r0 + c];
            if (vec) out << "VECTORS " << l.text << " float\n";
            else if (l.span == 1) out << "SCALARS " << l.text << " float\nLOOKUP_TABLE default\n";
            else out << "SCALARS " << l.text << " float " << l.span << "\nLOOKUP_TABLE default\n";
            writeFloats(v, vec ? 3 : (size_t)l.span);
            r0 += l.span;
        }
    }
    if (!out) throw std::runtime_error("write error on " + path);
}

struct StampedPose {
    TransformationParameters T;  // 4 x 4
    uint64_t stamp_ns = 0;
};

//! The example's trajectory CSV (build_map_from_scans_and_trajectory.cpp:15-172): columns found by name in the header.
inline std::vector<StampedPose> loadTrajectoryCSV(const std::string& path) {
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("Could not open file " + path);
    std::string header;
    std::getline(file, header);
    std::vector<std::string> cols;
    {
        std::istringstream hs(header);
        std::string c;
        while (std::getline(hs, c, ',')) {
            while (!c.empty() && (c.back() == '\r' || c.back() == ' ')) c.pop_back();
            cols.push_back(c);
        }
    }
    const char* wanted[9] = {"pose.pose.position.x",    "pose.pose.position.y",    "pose.pose.position.z",
                             "pose.pose.orientation.x", "pose.pose.orientation.y", "pose.pose.orientation.z",
                             "pose.pose.orientation.w", "header.stamp.sec",        "header.stamp.nanosec"};
    int idx[9];
    for (int k = 0; k < 9; ++k) {
        const auto it = std::find(cols.begin(), cols.end(), std::string(wanted[k]));
        if (it == cols.end()) throw std::runtime_error("Error: Required columns not found in the header.");
        idx[k] = (int)(it - cols.begin());
    }
    std::vector<StampedPose> out;
    std::string line;
    while (std::getline(file, line)) {
        if (line.empty()) continue;
        std::vector<std::string> tok;
        std::istringstream ls(line);
        std::string t;
        while (std::getline(ls, t, ',')) tok.push_back(t);
        if ((int)tok.size() <= *std::max_element(idx, idx + 9)) continue;  // short line
        double v[7];
        for (int k = 0; k < 7; ++k) v[k] = std::stod(tok[idx[k]]);
        const double x = v[3], y = v[4], z = v[5], w = v[6];  // Eigen::Quaterniond::toRotationMatrix (not renormalised)
        StampedPose p;
        p.T = TransformationParameters::Identity(4);
        p.T(0, 0) = (float)(1 - 2 * (y * y + z * z));
        p.T(0, 1) = (float)(2 * (x * y - z * w));
        p.T(0, 2) = (float)(2 * (x * z + y * w));
        p.T(1, 0) = (float)(2 * (x * y + z * w));
        p.T(1, 1) = (float)(1 - 2 * (x * x + z * z));
        p.T(1, 2) = (float)(2 * (y * z - x * w));
        p.T(2, 0) = (float)(2 * (x * z - y * w));
        p.T(2, 1) = (float)(2 * (y * z + x * w));
        p.T(2, 2) = (float)(1 - 2 * (x * x + y * y));
        p.T(0, 3) = (float)v[0];
        p.T(1, 3) = (float)v[1];
        p.T(2, 3) = (float)v[2];
        p.stamp_ns = (uint64_t)std::stoull(tok[idx[7]]) * 1000000000ull + (uint64_t)std::stoull(tok[idx[8]]);
        out.push_back(p);
    }
    return out;
}

}  // namespace io
}  // namespace norlab_icp_mapper_b200
