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
    // icp.setMap(cloud): every descriptor comes along (the reference assigns / copies the whole DataPoints, Map.cpp:575-588)
    bool setMap(const DataPoints& map) {
        check(ctx, b200icp_set_map(ctx, map.features.data(), dim + 1, map.normals.empty() ? nullptr : map.normals.data(), map.getNbPoints()));
        if (map.getNbPoints() > 0) {
            if (!map.probabilityDynamic.empty()) check(ctx, b200icp_map_set_prob(ctx, map.probabilityDynamic.data(), 0.f));
            check(ctx, b200icp_map_set_extra(ctx, map.descriptors.empty() ? nullptr : map.descriptors.data(), map.getDescriptorRows()));
            mapLabels = map.descriptorLabels;
        }
        return map.getNbPoints() > 0;
    }
    // ---- the scan slot: ONE device-resident cloud per context; `scanLabels` names the rows of its `extra` block ----
    // descriptors of the scan slot other than normals / probabilityDynamic: the caller's slot, and the snapshot an asynchronous map
    // update works on (the worker thread marks itself with UpdateThreadScope, mirroring b200icp_map_begin_update's routing)
    Labels scanLabelsMain, scanLabelsUpd;
    Labels mapLabels;   // the same for the device-resident map
    static bool& onUpdateThread() {
        static thread_local bool flag = false;
        return flag;
    }
    struct UpdateThreadScope {
        UpdateThreadScope() { onUpdateThread() = true; }
        ~UpdateThreadScope() { onUpdateThread() = false; }
    };
    Labels& scanLabelsRef() { return onUpdateThread() ? scanLabelsUpd : scanLabelsMain; }
    const Labels& scanLabelsRef() const { return onUpdateThread() ? scanLabelsUpd : scanLabelsMain; }
    // the scan in the slot becomes the input of an asynchronous update (Mapper::updateMap, isOnline)
    void snapshotScan() {
        check(ctx, b200icp_scan_snapshot(ctx));
        scanLabelsUpd = scanLabelsMain;
        scanLabelsMain.clear();
    }
    bool scanHas(const std::string& name) const {
        int32_t hn = 0, hp = 0;
        b200icp_scan_info(ctx, &hn, &hp, nullptr);
        if (name == "normals") return hn != 0;
        if (name == "probabilityDynamic") return hp != 0;
        return labelStartingRow(scanLabelsRef(), name) >= 0;
    }
    // host copy of a device-resident scan (modules without a device entry point use it)
    DataPoints materialize(const DataPoints& cloud) {
        if (!cloud.onDevice) return cloud;
        DataPoints out;
        out.dim = cloud.dim;
        int64_t n = 0;
        int32_t hn = 0, hp = 0, xr = 0;
        check(ctx, b200icp_scan_info(ctx, &hn, &hp, &xr));
        out.features.resize((size_t)cloud.deviceCount * (dim + 1));
        check(ctx, b200icp_scan_download(ctx, out.features.data(), cloud.deviceCount, &n));
        if (hn) out.normals.resize((size_t)n * dim);
        if (hp) out.probabilityDynamic.resize((size_t)n);
        if (xr) out.descriptors.resize((size_t)n * xr);
        out.descriptorLabels = scanLabelsRef();
        if (n > 0 && (hn || hp || xr))
            check(ctx, b200icp_scan_download_descriptors(ctx, hn ? out.normals.data() : nullptr, hp ? out.probabilityDynamic.data() : nullptr,
                                                         xr ? out.descriptors.data() : nullptr, n));
        return out;
    }
    // upload once, features and descriptors; the returned cloud refers to the context's scan slot
    DataPoints toDevice(const DataPoints& cloud) {
        if (cloud.onDevice) return cloud;
        check(ctx, b200icp_scan_upload(ctx, cloud.features.data(), dim + 1, cloud.getNbPoints()));
        std::vector<int32_t> rotating;  // libpointmatcher rotates `normals` and `observationDirections` with the cloud
        const int r0 = labelStartingRow(cloud.descriptorLabels, "observationDirections");
        if (r0 >= 0) rotating.push_back(r0);
        check(ctx, b200icp_scan_set_descriptors(ctx, cloud.normals.empty() ? nullptr : cloud.normals.data(),
                                                cloud.probabilityDynamic.empty() ? nullptr : cloud.probabilityDynamic.data(),
                                                cloud.descriptors.empty() ? nullptr : cloud.descriptors.data(), cloud.getDescriptorRows(),
                                                rotating.empty() ? nullptr : rotating.data(), (int32_t)rotating.size()));
        scanLabelsRef() = cloud.descriptors.empty() ? Labels() : cloud.descriptorLabels;
        DataPoints out;
        out.dim = cloud.dim;
        out.onDevice = true;
        out.deviceCount = cloud.getNbPoints();
        return out;
    }
    // DataPoints::concatenate's descriptor rule ahead of `map.concatenate(scan)` on the device: both sides keep the descriptors
    // they have in common, in the map's order (an empty map takes the scan's)
    void reconcileScanWithMap(bool mapIsEmpty) {
        Labels& scanLabels = scanLabelsRef();
        if (mapIsEmpty) {
            mapLabels = scanLabels;
            return;
        }
        const Labels common = commonLabels(mapLabels, scanLabels);
        if (!(common == mapLabels)) {
            const std::vector<int32_t> rows = rowsOf(mapLabels, common);
            check(ctx, b200icp_map_select_extra(ctx, rows.data(), (int32_t)rows.size()));
            mapLabels = common;
        }
        if (!(common == scanLabels)) {
            const std::vector<int32_t> rows = rowsOf(scanLabels, common);
            check(ctx, b200icp_scan_select_extra(ctx, rows.data(), (int32_t)rows.size()));
            scanLabels = common;
        }
    }
    TransformationParameters operator()(const DataPoints& cloud) {
        TransformationParameters T = TransformationParameters::Identity(dim + 1);
        // (a reading that carries `normals` hands them over: SurfaceNormalOutlierFilter compares them with the map's)
        // and `maxSearchDist` gives every point its own search radius (KDTreeMatcher)
        std::vector<float> maxSearchDist;
        if (!cloud.onDevice && labelStartingRow(cloud.descriptorLabels, "maxSearchDist") >= 0) maxSearchDist = cloud.getDescriptorCopyByName("maxSearchDist");
        const int32_t rc = cloud.onDevice ? b200icp_scan_register(ctx, nullptr, T.m, &last)
                                          : b200icp_register_descriptors(ctx, cloud.features.data(), dim + 1, cloud.getNbPoints(),
                                                                         cloud.normals.empty() ? nullptr : cloud.normals.data(),
                                                                         maxSearchDist.empty() ? nullptr : maxSearchDist.data(), nullptr, T.m, &last);
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
    if (labelStartingRow(out.descriptorLabels, "observationDirections") >= 0 && out.getNbPoints() > 0) {
        // rotated like the normals (LPM TransformationsImpl.cpp): R applied through the same entry point on a dummy feature block
        std::vector<float> od = out.getDescriptorCopyByName("observationDirections");
        std::vector<float> dummy((size_t)out.getNbPoints() * (in.dim + 1), 0.f);
        TransformationParameters R = T;
        for (int r = 0; r < in.dim; ++r) R(r, in.dim) = 0.f;
        ICPSequence::check(icp.context(), b200icp_transform(icp.context(), dummy.data(), in.dim + 1, od.data(), out.getNbPoints(), R.m));
        const int r0 = labelStartingRow(out.descriptorLabels, "observationDirections"), rows = out.getDescriptorRows();
        for (int64_t i = 0; i < out.getNbPoints(); ++i)
            for (int c = 0; c < in.dim; ++c) out.descriptors[i * rows + r0 + c] = od[i * in.dim + c];
    }
    return out;
}

}  // namespace norlab_icp_mapper_b200
