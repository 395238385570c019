// DataPoints.h -- the slice of libpointmatcher's `PM::DataPoints` / `TransformationParameters` this
// path needs on the host: column-major fp32 features (dim+1) x N with the homogeneous row last, and
// named descriptors.  Memory layout is Eigen's, so a maintainer with libpointmatcher can construct one
// from `cloud.features.data()` / `cloud.descriptors.data()` without copying semantics changing.
//
// Descriptors: `normals` and `probabilityDynamic` -- the two this path computes with -- have their own
// members; every other one (intensity, t, ring, observationDirections ... whatever the sensor driver
// attached) lives in `descriptors` with `descriptorLabels`, exactly libpointmatcher's
// descriptorLabels / descriptors pair.  They are carried through the input filters, the rigid
// transforms, the MapperModules and the map download with DataPoints::concatenate's rule: a descriptor
// survives a concatenation only if both clouds have it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace norlab_icp_mapper_b200 {

// C++ face of the libpointmatcher exception types the reference lets propagate out of processInput
struct ConvergenceError : std::runtime_error { using std::runtime_error::runtime_error; };
struct TransformationError : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidField : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidParameter : std::runtime_error { using std::runtime_error::runtime_error; };

struct Label {  // PM::DataPoints::Label
    std::string text;
    int span = 1;
    bool operator==(const Label& o) const { return text == o.text && span == o.span; }
};
typedef std::vector<Label> Labels;

inline int labelRows(const Labels& labels) {
    int r = 0;
    for (const Label& l : labels) r += l.span;
    return r;
}
// first row of `name` inside a block laid out by `labels`, or -1
inline int labelStartingRow(const Labels& labels, const std::string& name) {
    int r = 0;
    for (const Label& l : labels) {
        if (l.text == name) return r;
        r += l.span;
    }
    return -1;
}
// labels of `a` that `b` also carries (same span), in a's order: what DataPoints::concatenate keeps
inline Labels commonLabels(const Labels& a, const Labels& b) {
    Labels out;
    for (const Label& l : a) {
        const auto it = std::find_if(b.begin(), b.end(), [&](const Label& m) { return m.text == l.text; });
        if (it == b.end()) continue;
        if (it->span != l.span) throw InvalidField("DataPoints::concatenate: descriptor " + l.text + " has different dimensions in the two clouds");
        out.push_back(l);
    }
    return out;
}
// rows of `from` that make up `wanted` (a sub-list of from), in wanted's order
inline std::vector<int32_t> rowsOf(const Labels& from, const Labels& wanted) {
    std::vector<int32_t> rows;
    for (const Label& l : wanted) {
        const int r0 = labelStartingRow(from, l.text);
        for (int c = 0; c < l.span; ++c) rows.push_back(r0 + c);
    }
    return rows;
}

struct DataPoints {
    int dim = 3;                  // euclidean dimension (2 or 3); features has dim + 1 rows
    std::vector<float> features;  // (dim + 1) x N, column-major
    std::vector<float> normals;   // dim x N, column-major, or empty (descriptor absent)
    std::vector<float> probabilityDynamic;  // 1 x N or empty (descriptor absent)
    Labels descriptorLabels;                // the other descriptors: names and row spans, in storage order
    std::vector<float> descriptors;         // labelRows(descriptorLabels) x N, column-major
    // Device-resident scan: the vectors are empty and the cloud lives in the ICP context's scan slot (b200icp_scan_*).
    // Set by Mapper::applyInputFilters / processInput for their temporaries so that the filters, the transforms, icp(input)
    // and the map insert work on ONE upload; anything that needs the numbers on the host calls ICPSequence::materialize().
    bool onDevice = false;
    int64_t deviceCount = 0;
    int64_t getNbPoints() const { return onDevice ? deviceCount : (features.empty() ? 0 : (int64_t)features.size() / (dim + 1)); }
    int getDescriptorRows() const { return labelRows(descriptorLabels); }
    bool descriptorExists(const std::string& name) const {
        if (onDevice) throw std::runtime_error("descriptorExists: ask the ICPSequence about a device-resident cloud (scanHas)");
        if (name == "normals") return !normals.empty();
        if (name == "probabilityDynamic") return !probabilityDynamic.empty();
        return labelStartingRow(descriptorLabels, name) >= 0;
    }
    // PM::DataPoints::addDescriptor: `data` is rows x N, column-major; an existing descriptor of that name is replaced
    void addDescriptor(const std::string& name, int rows, const std::vector<float>& data) {
        const int64_t n = getNbPoints();
        if ((int64_t)data.size() != n * rows) throw InvalidField("addDescriptor: " + name + " has the wrong size");
        if (name == "normals") {
            normals = data;
            return;
        }
        if (name == "probabilityDynamic") {
            probabilityDynamic = data;
            return;
        }
        if (labelStartingRow(descriptorLabels, name) >= 0) removeDescriptor(name);
        const int old = getDescriptorRows();
        std::vector<float> merged((size_t)n * (old + rows));
        for (int64_t i = 0; i < n; ++i) {
            for (int c = 0; c < old; ++c) merged[i * (old + rows) + c] = descriptors[i * old + c];
            for (int c = 0; c < rows; ++c) merged[i * (old + rows) + old + c] = data[i * rows + c];
        }
        descriptors.swap(merged);
        descriptorLabels.push_back(Label{name, rows});
    }
    void removeDescriptor(const std::string& name) {
        if (name == "normals") {
            normals.clear();
            return;
        }
        if (name == "probabilityDynamic") {
            probabilityDynamic.clear();
            return;
        }
        Labels keep;
        for (const Label& l : descriptorLabels)
            if (l.text != name) keep.push_back(l);
        selectDescriptors(keep);
    }
    // keep the listed descriptors (a sub-list of descriptorLabels), in that order
    void selectDescriptors(const Labels& keep) {
        const std::vector<int32_t> rows = rowsOf(descriptorLabels, keep);
        const int old = getDescriptorRows(), now = (int)rows.size();
        const int64_t n = getNbPoints();
        std::vector<float> out((size_t)n * now);
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < now; ++c) out[i * now + c] = descriptors[i * old + rows[c]];
        descriptors.swap(out);
        descriptorLabels = keep;
    }
    std::vector<float> getDescriptorCopyByName(const std::string& name) const {
        if (name == "normals" && !normals.empty()) return normals;
        if (name == "probabilityDynamic" && !probabilityDynamic.empty()) return probabilityDynamic;
        const int r0 = labelStartingRow(descriptorLabels, name);
        if (r0 < 0) throw InvalidField("Descriptor " + name + " not found");
        int span = 1;
        for (const Label& l : descriptorLabels)
            if (l.text == name) span = l.span;
        const int rows = getDescriptorRows();
        const int64_t n = getNbPoints();
        std::vector<float> out((size_t)n * span);
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < span; ++c) out[i * span + c] = descriptors[i * rows + r0 + c];
        return out;
    }
    // PM::DataPoints::concatenate (host clouds): features appended; only descriptors both clouds carry survive
    void concatenate(const DataPoints& o) {
        if (onDevice || o.onDevice) throw std::runtime_error("concatenate: host clouds only");
        const bool first = getNbPoints() == 0 && features.empty();
        if (first) {
            *this = o;
            return;
        }
        if (normals.empty() || o.normals.empty()) normals.clear();
        else normals.insert(normals.end(), o.normals.begin(), o.normals.end());
        if (probabilityDynamic.empty() || o.probabilityDynamic.empty()) probabilityDynamic.clear();
        else probabilityDynamic.insert(probabilityDynamic.end(), o.probabilityDynamic.begin(), o.probabilityDynamic.end());
        const Labels common = commonLabels(descriptorLabels, o.descriptorLabels);
        if (!(common == descriptorLabels)) selectDescriptors(common);
        DataPoints other = o;  // (a copy only when the layouts differ)
        const DataPoints* src = &o;
        if (!(o.descriptorLabels == common)) {
            other.selectDescriptors(common);
            src = &other;
        }
        descriptors.insert(descriptors.end(), src->descriptors.begin(), src->descriptors.end());
        features.insert(features.end(), o.features.begin(), o.features.end());
    }
};

// (dim + 1) x (dim + 1), column-major -- PM::TransformationParameters
struct TransformationParameters {
    int n = 4;
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    static TransformationParameters Identity(int n_) {
        TransformationParameters T;
        T.n = n_;
        for (int i = 0; i < 16; ++i) T.m[i] = 0.f;
        for (int i = 0; i < n_; ++i) T.m[i * n_ + i] = 1.f;
        return T;
    }
    float& operator()(int r, int c) { return m[c * n + r]; }
    float operator()(int r, int c) const { return m[c * n + r]; }
    TransformationParameters operator*(const TransformationParameters& o) const {  // fp32, k-order like the oracle
        TransformationParameters R = Identity(n);
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r) {
                float acc = 0.f;
                for (int k = 0; k < n; ++k) acc += m[k * n + r] * o.m[c * n + k];
                R.m[c * n + r] = acc;
            }
        return R;
    }
};

}  // namespace norlab_icp_mapper_b200
