"""Python restatement of the reference's Mapper::processInput / Map::updateLocalPointCloud /
Map::updatePose sequence on top of the CPU oracle (test infrastructure).  Follows
/root/reference/norlab_icp_mapper/Mapper.cpp:194-288 and Map.cpp:246-534 step by step, including the
sensor-frame round trip around the post filters (Map.cpp:523-525)."""
import math

import numpy as np

import os
import sys

import oracle_binding as ob

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

CELL, BUFFER = 20.0, 2


class RefMapper:
    def __init__(self, cfg, update_distance=1.0, min_dist_new_point=0.15, surface_normal_knn=10, sensor_max_range=200.0):
        self.cfg = cfg
        self.icp = ob.OracleICP(cfg)
        self.map = None
        self.normals = None
        self.pose = np.eye(4, dtype=np.float32)
        self.last_pose_update = np.eye(4, dtype=np.float32)
        self.d_upd, self.min_dist, self.knn, self.range = update_distance, min_dist_new_point, surface_normal_knn, sensor_max_range
        self.trajectory = []
        self.updated = False

    def _update_local_point_cloud(self, inp, pose):
        if self.map is None:
            self.map = inp.copy()
        else:
            kept, keep = ob.point_distance_keep(self.map, inp, self.min_dist)
            self.map = np.ascontiguousarray(np.r_[self.map, inp[keep]])
        if self.knn > 0:
            inv = np.linalg.inv(pose.astype(np.float32)).astype(np.float32)
            rc, in_sensor, _ = ob.transform(self.map, inv)
            rc2, nrm = ob.surface_normals(in_sensor, self.knn)
            rc3, back, nrm_back = ob.transform(in_sensor, pose, nrm)
            assert rc == rc2 == rc3 == 0
            self.map, self.normals = back, nrm_back
        self.icp.set_map(self.map, self.normals)

    def process_input(self, scan, T_est, stamp):
        T_est = np.asarray(T_est, np.float32)
        rc, inp, _ = ob.transform(scan, T_est)
        assert rc == 0
        self.updated = False
        if self.map is None:
            corrected = T_est
            self._update_local_point_cloud(inp, corrected)
            self.last_pose_update = corrected
            self.updated = True
        else:
            rc, corr, res, _, _ = self.icp.register(inp)
            assert rc == 0, self.icp.last_error()
            corrected = (corr.astype(np.float32) @ T_est).astype(np.float32)
            if np.linalg.norm(corrected[:3, 3] - self.last_pose_update[:3, 3]) > self.d_upd:
                rc, inp2, _ = ob.transform(inp, corr)
                self._update_local_point_cloud(inp2, corrected)
                self.last_pose_update = corrected
                self.updated = True
        self.pose = corrected
        self.trajectory.append(corrected)


class RefWindow:
    """Map::updatePose's cell-window state machine, written out per axis exactly as the reference
    spells it (Map.cpp:246-460) -- the independent check of the generic loop in host/Map.cpp."""

    def __init__(self, sensor_max_range, is3d=True):
        self.r, self.is3d, self.first = sensor_max_range, is3d, True
        self.inf = [0, 0, 0]
        self.sup = [0, 0, 0]

    def _inf(self, x):
        return int(math.ceil(((np.float32(x) - np.float32(self.r)) / np.float32(CELL)) - 1.0))

    def _sup(self, x):
        return int(math.floor((np.float32(x) + np.float32(self.r)) / np.float32(CELL)))

    def update(self, pos):
        ups = []
        if self.first:
            for a in range(3 if self.is3d else 2):
                self.inf[a], self.sup[a] = self._inf(pos[a]), self._sup(pos[a])
            ups.append("unload-all")
            ups.append((self.inf[0] - BUFFER, self.sup[0] + BUFFER, self.inf[1] - BUFFER, self.sup[1] + BUFFER,
                        self.inf[2] - BUFFER, self.sup[2] + BUFFER, 1))
            self.first = False
            return ups
        i, s, B = self.inf, self.sup, BUFFER
        # rows
        n = self._inf(pos[0])
        if abs(n - i[0]) >= 2:
            if n < i[0]:
                nb = i[0] - n
                ups.append((n - B, n - B + nb - 1, i[1] - B, s[1] + B, i[2] - B, s[2] + B, 1))
            if n > i[0]:
                nb = n - i[0]
                ups.append((i[0] - B, i[0] - B + nb - 1, i[1] - B, s[1] + B, i[2] - B, s[2] + B, 0))
            i[0] = n
        n = self._sup(pos[0])
        if abs(n - s[0]) >= 2:
            if n < s[0]:
                nb = s[0] - n
                ups.append((s[0] + B - nb + 1, s[0] + B, i[1] - B, s[1] + B, i[2] - B, s[2] + B, 0))
            if n > s[0]:
                nb = n - s[0]
                ups.append((n + B - nb + 1, n + B, i[1] - B, s[1] + B, i[2] - B, s[2] + B, 1))
            s[0] = n
        # columns
        n = self._inf(pos[1])
        if abs(n - i[1]) >= 2:
            if n < i[1]:
                nb = i[1] - n
                ups.append((i[0] - B, s[0] + B, n - B, n - B + nb - 1, i[2] - B, s[2] + B, 1))
            if n > i[1]:
                nb = n - i[1]
                ups.append((i[0] - B, s[0] + B, i[1] - B, i[1] - B + nb - 1, i[2] - B, s[2] + B, 0))
            i[1] = n
        n = self._sup(pos[1])
        if abs(n - s[1]) >= 2:
            if n < s[1]:
                nb = s[1] - n
                ups.append((i[0] - B, s[0] + B, s[1] + B - nb + 1, s[1] + B, i[2] - B, s[2] + B, 0))
            if n > s[1]:
                nb = n - s[1]
                ups.append((i[0] - B, s[0] + B, n + B - nb + 1, n + B, i[2] - B, s[2] + B, 1))
            s[1] = n
        if self.is3d:
            n = self._inf(pos[2])
            if abs(n - i[2]) >= 2:
                if n < i[2]:
                    nb = i[2] - n
                    ups.append((i[0] - B, s[0] + B, i[1] - B, s[1] + B, n - B, n - B + nb - 1, 1))
                if n > i[2]:
                    nb = n - i[2]
                    ups.append((i[0] - B, s[0] + B, i[1] - B, s[1] + B, i[2] - B, i[2] - B + nb - 1, 0))
                i[2] = n
            n = self._sup(pos[2])
            if abs(n - s[2]) >= 2:
                if n < s[2]:
                    nb = s[2] - n
                    ups.append((i[0] - B, s[0] + B, i[1] - B, s[1] + B, s[2] + B - nb + 1, s[2] + B, 0))
                if n > s[2]:
                    nb = n - s[2]
                    ups.append((i[0] - B, s[0] + B, i[1] - B, s[1] + B, n + B - nb + 1, n + B, 1))
                s[2] = n
        return ups


class RefExampleMapper:
    """The pipeline of the reference's bundled configuration (examples/config.yaml) restated on the
    oracle pieces: input BoundingBox x2 + AddDescriptor{probabilityDynamic 0.6}; mapper modules
    DynamicPoints then Octree; post SurfaceNormal{knn 10} + CutAtDescriptorThreshold{0.65}; delay
    update condition; Identity error minimiser.  samplingMethod is 0 (first) instead of the shipped 1
    (random), which no two implementations can reproduce."""

    def __init__(self, max_size=0.15, knn=10, cut=0.65, delay=0.05, dyn=None, boxes=()):
        import modules_oracle as mo
        self.mo, self.max_size, self.knn, self.cut, self.delay = mo, max_size, knn, cut, delay
        self.dyn = dyn or {}
        self.boxes = boxes
        self.map = self.normals = self.prob = None
        self.last_t = 0.0
        self.updated = False

    def apply_input_filters(self, scan, sensor_max_range=200.0):
        keep = self.mo.distance_limit_keep(scan, sensor_max_range)
        for lo, hi in self.boxes:
            keep &= self.mo.bounding_box_keep(scan, lo, hi, True)
        return np.ascontiguousarray(scan[keep])

    def _post(self, pose):
        inv = np.linalg.inv(pose.astype(np.float32)).astype(np.float32)
        rc, in_sensor, _ = ob.transform(self.map, inv)
        rc2, nrm = ob.surface_normals(in_sensor, self.knn)
        rc3, back, nrm_back = ob.transform(in_sensor, pose, nrm)
        self.map, self.normals = back, nrm_back
        keep = self.mo.cut_at_descriptor_threshold(self.prob, self.cut, True)
        self.map, self.normals, self.prob = self.map[keep], self.normals[keep], self.prob[keep]

    def _octree(self, inp, inp_prob):
        allp = np.r_[self.map, inp] if self.map is not None else inp
        allq = np.r_[self.prob, inp_prob] if self.prob is not None else inp_prob
        order, feat, desc = self.mo.octree_grid_filter(allp, self.max_size, 0, descriptors=allq[:, None])
        self.map, self.prob, self.normals = np.ascontiguousarray(feat), desc[:, 0].copy(), None

    def process_input(self, scan, T_est, stamp):
        T_est = np.asarray(T_est, np.float32)
        rc, inp, _ = ob.transform(scan, T_est)
        inp_prob = np.full(len(inp), 0.6, np.float32)
        self.updated = False
        if self.map is None:
            self.map, self.prob = inp.copy(), inp_prob.copy()  # DynamicPoints::createMap copies the input
            self._octree(inp, inp_prob)                         # Octree::inPlaceUpdateMap concatenates it again
            self._post(T_est)
            self.last_t, self.updated = stamp, True
        elif (stamp - self.last_t) > self.delay:                # Identity minimiser: correction = I, pose = T_est
            self.prob, _ = self.mo.dynamic_points_update(inp, self.map, self.normals, self.prob, T_est, **self.dyn)
            self._octree(inp, inp_prob)
            self._post(T_est)
            self.last_t, self.updated = stamp, True
        self.pose = T_est
