// DataPoints.h -- the slice of libpointmatcher's `PM::DataPoints` / `TransformationParameters` this
// path needs on the host: column-major fp32 features (dim+1) x N with the homogeneous row last, and
// an optional `normals` descriptor dim x N.  Memory layout is Eigen's, so a maintainer with
// libpointmatcher can construct one from `cloud.features.data()` without copying semantics changing.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace norlab_icp_mapper_b200 {

struct DataPoints {
    int dim = 3;                  // euclidean dimension (2 or 3); features has dim + 1 rows
    std::vector<float> features;  // (dim + 1) x N, column-major
    std::vector<float> normals;   // dim x N, column-major, or empty (descriptor absent)
    std::vector<float> probabilityDynamic;  // 1 x N or empty (descriptor absent)
    // Device-resident scan: `features` is empty and the (dim + 1) x deviceCount matrix lives in the ICP context's scan
    // slot (b200icp_scan_*).  Set by Mapper::processInput for its temporaries so that the transforms, icp(input) and the
    // map insert work on ONE upload; anything that needs the numbers on the host calls ICPSequence::materialize().
    bool onDevice = false;
    int64_t deviceCount = 0;
    int64_t getNbPoints() const { return onDevice ? deviceCount : (features.empty() ? 0 : (int64_t)features.size() / (dim + 1)); }
    bool descriptorExists(const std::string& name) const {
        return (name == "normals" && !normals.empty()) || (name == "probabilityDynamic" && !probabilityDynamic.empty());
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

// C++ face of the libpointmatcher exception types the reference lets propagate out of processInput
struct ConvergenceError : std::runtime_error { using std::runtime_error::runtime_error; };
struct TransformationError : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidField : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidParameter : std::runtime_error { using std::runtime_error::runtime_error; };

}  // namespace norlab_icp_mapper_b200
