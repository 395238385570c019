// ICPSequence.h -- `PM::ICPSequence icp` (reference Mapper.h:23) over the C-ABI of libb200icp.so.
// Same methods the reference calls: setMap (Map.cpp:111,178,528,581), operator() (Mapper.cpp:213),
// getOverlap (Mapper.cpp:219); libpointmatcher's exceptions are re-thrown from the status codes.
#pragma once
#include "DataPoints.h"
#include "b200icp.h"

namespace norlab_icp_mapper_b200 {

class ICPSequence {
    b200icp_ctx* ctx = nullptr;
    b200icp_result last{};
    int dim;

   public:
    static void check(b200icp_ctx* c, int32_t rc) {
        switch (rc) {
            case B200ICP_OK:
                return;
            case B200ICP_ERR_CONVERGENCE:
            case B200ICP_ERR_BOUND:
            case B200ICP_ERR_NAN:
                throw ConvergenceError(b200icp_last_error(c));
            case B200ICP_ERR_TRANSFORM:
                throw TransformationError(b200icp_last_error(c));
            case B200ICP_ERR_INVALID_FIELD:
                throw InvalidField(b200icp_last_error(c));
            default:
                throw std::runtime_error(b200icp_last_error(c));
        }
    }
    ICPSequence(const b200icp_config& cfg, int device) : dim(cfg.dim) { check(nullptr, b200icp_create(&cfg, device, &ctx)); }
    ~ICPSequence() { b200icp_destroy(ctx); }
    ICPSequence(const ICPSequence&) = delete;
    ICPSequence& operator=(const ICPSequence&) = delete;

    b200icp_ctx* context() { return ctx; }
    bool hasMap() const { return b200icp_map_size(ctx) > 0; }
    bool setMap(const DataPoints& map) {
        check(ctx, b200icp_set_map(ctx, map.features.data(), dim + 1, map.normals.empty() ? nullptr : map.normals.data(), map.getNbPoints()));
        return map.getNbPoints() > 0;
    }
    // host copy of a device-resident scan (modules without a device entry point use it)
    DataPoints materialize(const DataPoints& cloud) {
        if (!cloud.onDevice) return cloud;
        DataPoints out = cloud;
        out.onDevice = false;
        out.deviceCount = 0;
        int64_t n = 0;
        out.features.resize((size_t)cloud.deviceCount * (dim + 1));
        check(ctx, b200icp_scan_download(ctx, out.features.data(), cloud.deviceCount, &n));
        return out;
    }
    // upload once; the returned cloud refers to the context's scan slot
    DataPoints toDevice(const DataPoints& cloud) {
        check(ctx, b200icp_scan_upload(ctx, cloud.features.data(), dim + 1, cloud.getNbPoints()));
        DataPoints out;
        out.dim = cloud.dim;
        out.probabilityDynamic = cloud.probabilityDynamic;
        out.onDevice = true;
        out.deviceCount = cloud.getNbPoints();
        return out;
    }
    TransformationParameters operator()(const DataPoints& cloud) {
        TransformationParameters T = TransformationParameters::Identity(dim + 1);
        // (a reading that carries `normals` hands them over: SurfaceNormalOutlierFilter compares them with the map's)
        const int32_t rc = cloud.onDevice ? b200icp_scan_register(ctx, nullptr, T.m, &last)
                                          : b200icp_register_normals(ctx, cloud.features.data(), dim + 1, cloud.getNbPoints(),
                                                                     cloud.normals.empty() ? nullptr : cloud.normals.data(), nullptr, T.m, &last);
        if (rc == B200ICP_ERR_NO_MAP) return T;  // LPM: no map -> identity
        check(ctx, rc);
        return T;
    }
    float getOverlap() const { return last.overlap; }  // errorMinimizer->getOverlap()
    const b200icp_result& lastResult() const { return last; }
};

// PM::Transformation("RigidTransformation")::compute (Mapper.cpp:197,221)
inline DataPoints rigidTransform(ICPSequence& icp, const DataPoints& in, const TransformationParameters& T) {
    if (in.onDevice) {  // the slot is transformed in place: `in` and the result are the same device cloud
        ICPSequence::check(icp.context(), b200icp_scan_transform(icp.context(), T.m));
        return in;
    }
    DataPoints out = in;
    ICPSequence::check(icp.context(), b200icp_transform(icp.context(), out.features.data(), in.dim + 1,
                                                        out.normals.empty() ? nullptr : out.normals.data(), out.getNbPoints(), T.m));
    return out;
}

}  // namespace norlab_icp_mapper_b200
