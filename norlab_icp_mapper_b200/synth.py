"""Seeded synthetic workloads for the ICP hot path.  numpy only.

The reference ships neither benchmarks nor inputs of the BASELINE.json sizes, so bench.py and the
tests generate them here: a 3-D outdoor-like world (undulating ground, boxes, bounding walls) sampled
area-uniformly with analytic normals, and a 2-D polygonal room.  All clouds are returned as
(N, dim+1) C-contiguous fp32 arrays, which is byte-for-byte the reference's column-major
(dim+1) x N `DataPoints::features` matrix.
"""
import numpy as np


# ----------------------------------------------------------------------------------------------
# transforms
# ----------------------------------------------------------------------------------------------
def rpy_to_R(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def make_T(t, rpy_deg=(0.0, 0.0, 0.0)):
    T = np.eye(4)
    T[:3, :3] = rpy_to_R(*np.deg2rad(rpy_deg))
    T[:3, 3] = t
    return T


def make_T2(t, yaw_deg=0.0):
    a = np.deg2rad(yaw_deg)
    T = np.eye(3)
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    T[:2, 2] = t
    return T


def apply_T(T, pts):
    d = T.shape[0] - 1
    return pts[:, :d] @ T[:d, :d].T + T[:d, d]


def homog(pts):
    return np.ascontiguousarray(np.c_[pts, np.ones(len(pts))], dtype=np.float32)


def pose_error(T_a, T_b):
    """(rotation angle in rad, translation norm in m) between two homogeneous transforms."""
    d = T_a.shape[0] - 1
    dT = np.linalg.inv(np.asarray(T_a, np.float64)) @ np.asarray(T_b, np.float64)
    if d == 3:
        c = np.clip((np.trace(dT[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
        ang = float(np.arccos(c))
        if ang < 1e-3:  # arccos is ill-conditioned near 0: use the skew part
            S = dT[:3, :3] - dT[:3, :3].T
            ang = float(0.5 * np.sqrt(S[2, 1] ** 2 + S[0, 2] ** 2 + S[1, 0] ** 2))
    else:
        ang = float(abs(np.arctan2(dT[1, 0], dT[0, 0])))
    return ang, float(np.linalg.norm(T_a[:d, d].astype(np.float64) - T_b[:d, d].astype(np.float64)))


# ----------------------------------------------------------------------------------------------
# 3-D world: ground z = 0.3 sin(0.05 x) cos(0.07 y), axis-aligned boxes, four bounding walls
# ----------------------------------------------------------------------------------------------
class World3D:
    def __init__(self, seed=1234, size=(200.0, 200.0), n_boxes=40, wall_height=10.0):
        rng = np.random.default_rng(seed)
        self.size = size
        hx, hy = size[0] / 2, size[1] / 2
        # rectangular patches: origin o, edge vectors u, v, unit normal n
        patches = []
        for _ in range(n_boxes):
            sx, sy = rng.uniform(4, 20, 2)
            h = rng.uniform(2, 12)
            cx = rng.uniform(-hx + sx, hx - sx)
            cy = rng.uniform(-hy + sy, hy - sy)
            x0, x1, y0, y1 = cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2
            z0 = -0.5
            patches += [
                ((x0, y0, z0), (sx, 0, 0), (0, 0, h - z0), (0, -1, 0)),
                ((x0, y1, z0), (sx, 0, 0), (0, 0, h - z0), (0, 1, 0)),
                ((x0, y0, z0), (0, sy, 0), (0, 0, h - z0), (-1, 0, 0)),
                ((x1, y0, z0), (0, sy, 0), (0, 0, h - z0), (1, 0, 0)),
                ((x0, y0, h), (sx, 0, 0), (0, sy, 0), (0, 0, 1)),
            ]
        z0 = -0.5
        patches += [
            ((-hx, -hy, z0), (size[0], 0, 0), (0, 0, wall_height - z0), (0, 1, 0)),
            ((-hx, hy, z0), (size[0], 0, 0), (0, 0, wall_height - z0), (0, -1, 0)),
            ((-hx, -hy, z0), (0, size[1], 0), (0, 0, wall_height - z0), (1, 0, 0)),
            ((hx, -hy, z0), (0, size[1], 0), (0, 0, wall_height - z0), (-1, 0, 0)),
        ]
        self.o = np.array([p[0] for p in patches], np.float64)
        self.u = np.array([p[1] for p in patches], np.float64)
        self.v = np.array([p[2] for p in patches], np.float64)
        self.n = np.array([p[3] for p in patches], np.float64)
        self.patch_area = np.linalg.norm(np.cross(self.u, self.v), axis=1)
        self.ground_area = size[0] * size[1]

    def ground_z(self, x, y):
        return 0.3 * np.sin(0.05 * x) * np.cos(0.07 * y)

    def sample(self, n, rng, noise=0.01, center=None, radius=None):
        """n area-uniform surface samples (+ N(0, noise) per axis) and their analytic unit normals
        with random sign.  With center/radius only samples within `radius` of `center` are kept
        (rejection, so the density stays uniform)."""
        pts_l, nrm_l, have = [], [], 0
        areas = np.r_[self.ground_area, self.patch_area]
        prob = areas / areas.sum()
        while have < n:
            m = int((n - have) * (1.3 if radius is None else 8.0)) + 1024
            which = rng.choice(len(areas), size=m, p=prob)
            a, b = rng.random(m), rng.random(m)
            P = np.empty((m, 3))
            N = np.empty((m, 3))
            g = which == 0
            x = (a[g] - 0.5) * self.size[0]
            y = (b[g] - 0.5) * self.size[1]
            P[g] = np.c_[x, y, self.ground_z(x, y)]
            dzdx = 0.3 * 0.05 * np.cos(0.05 * x) * np.cos(0.07 * y)
            dzdy = -0.3 * 0.07 * np.sin(0.05 * x) * np.sin(0.07 * y)
            ng = np.c_[-dzdx, -dzdy, np.ones_like(x)]
            N[g] = ng / np.linalg.norm(ng, axis=1, keepdims=True)
            k = which[~g] - 1
            P[~g] = self.o[k] + a[~g, None] * self.u[k] + b[~g, None] * self.v[k]
            N[~g] = self.n[k]
            if radius is not None:
                keep = np.linalg.norm(P - np.asarray(center), axis=1) < radius
                P, N = P[keep], N[keep]
            pts_l.append(P)
            nrm_l.append(N)
            have += len(P)
        P = np.concatenate(pts_l)[:n]
        N = np.concatenate(nrm_l)[:n]
        P = P + rng.normal(0.0, noise, P.shape)
        N = N * rng.choice([-1.0, 1.0], size=(len(N), 1))
        return P, N


def make_pair_3d(n_map=2_000_000, n_scan=100_000, seed=1234, world_size=(200.0, 200.0), n_boxes=40,
                 sensor=(3.0, -2.0, 1.5), sensor_rpy_deg=(0.0, 0.0, 20.0), scan_radius=80.0,
                 dt=(0.30, -0.20, 0.10), drpy_deg=(0.3, -0.3, 1.0), noise=0.01, scan_seed=None):
    """BASELINE.json config 2 (SURVEY.md 8d cfg 2) at the default sizes.

    Returns a dict with
      map        (n_map, 4) fp32, map frame         normals (n_map, 3) fp32, unit, random sign
      scan       (n_scan, 4) fp32, SENSOR frame     (what Mapper::processInput receives)
      T_true     sensor pose in the map             T_est = T_true o perturbation (initial guess)
      reading    (n_scan, 4) fp32 = T_est * scan    (what icp(input) receives, Mapper.cpp:197)
      correction_true = T_true * T_est^-1           (what icp(input) should return)
    """
    world = World3D(seed=seed, size=world_size, n_boxes=n_boxes)
    rng_m = np.random.default_rng(seed)
    rng_s = np.random.default_rng(seed + 1 if scan_seed is None else scan_seed)  # (scan_seed: another scan of the same world and map)
    P, N = world.sample(n_map, rng_m, noise=noise)
    T_true = make_T(sensor, sensor_rpy_deg)
    S, _ = world.sample(n_scan, rng_s, noise=noise, center=sensor, radius=scan_radius)
    scan = apply_T(np.linalg.inv(T_true), S)
    T_est = T_true @ make_T(dt, drpy_deg)
    reading = apply_T(T_est, scan)
    return dict(map=homog(P), normals=np.ascontiguousarray(N, np.float32), scan=homog(scan),
                reading=homog(reading), T_true=T_true, T_est=T_est,
                correction_true=T_true @ np.linalg.inv(T_est))


# ----------------------------------------------------------------------------------------------
# 2-D world: rectangular room with polygonal obstacles (BASELINE.json config 4)
# ----------------------------------------------------------------------------------------------
class World2D:
    def __init__(self, seed=3000, size=(60.0, 40.0), n_obstacles=8):
        rng = np.random.default_rng(seed)
        hx, hy = size[0] / 2, size[1] / 2
        segs = [((-hx, -hy), (hx, -hy)), ((hx, -hy), (hx, hy)), ((hx, hy), (-hx, hy)), ((-hx, hy), (-hx, -hy))]
        for _ in range(n_obstacles):
            c = np.array([rng.uniform(-hx + 6, hx - 6), rng.uniform(-hy + 6, hy - 6)])
            m = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(0, 2 * np.pi, m))
            r = rng.uniform(1.5, 4.0, m)
            v = c + np.c_[r * np.cos(ang), r * np.sin(ang)]
            for i in range(m):
                segs.append((tuple(v[i]), tuple(v[(i + 1) % m])))
        self.a = np.array([s[0] for s in segs], np.float64)
        self.b = np.array([s[1] for s in segs], np.float64)
        self.len = np.linalg.norm(self.b - self.a, axis=1)

    def sample(self, n, rng, noise=0.01, center=None, radius=None):
        pts_l, nrm_l, have = [], [], 0
        prob = self.len / self.len.sum()
        while have < n:
            m = int((n - have) * (1.2 if radius is None else 4.0)) + 256
            k = rng.choice(len(self.len), size=m, p=prob)
            t = rng.random(m)[:, None]
            P = self.a[k] + t * (self.b[k] - self.a[k])
            d = (self.b[k] - self.a[k]) / self.len[k, None]
            N = np.c_[-d[:, 1], d[:, 0]]
            if radius is not None:
                keep = np.linalg.norm(P - np.asarray(center), axis=1) < radius
                P, N = P[keep], N[keep]
            pts_l.append(P)
            nrm_l.append(N)
            have += len(P)
        P = np.concatenate(pts_l)[:n] + rng.normal(0.0, noise, (n, 2))
        N = np.concatenate(nrm_l)[:n] * rng.choice([-1.0, 1.0], size=(n, 1))
        return P, N


def make_pair_2d(n_map=200_000, n_scan=10_000, seed=3000, sensor=(2.0, -1.0), sensor_yaw_deg=15.0,
                 scan_radius=25.0, dt=(0.15, -0.10), dyaw_deg=2.0, noise=0.01):
    world = World2D(seed=seed)
    P, N = world.sample(n_map, np.random.default_rng(seed), noise=noise)
    T_true = make_T2(sensor, sensor_yaw_deg)
    S, _ = world.sample(n_scan, np.random.default_rng(seed + 1), noise=noise, center=sensor, radius=scan_radius)
    scan = apply_T(np.linalg.inv(T_true), S)
    T_est = T_true @ make_T2(dt, dyaw_deg)
    reading = apply_T(T_est, scan)
    return dict(map=homog(P), normals=np.ascontiguousarray(N, np.float32), scan=homog(scan),
                reading=homog(reading), T_true=T_true, T_est=T_est,
                correction_true=T_true @ np.linalg.inv(T_est))
