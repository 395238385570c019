"""Second-opinion ICP in numpy/scipy (float64, scipy.spatial.cKDTree), written independently of
oracle/icp_oracle.c from the same upstream semantics.  Used to pin the C oracle
and to generate tests/golden/.  Test infrastructure only."""
import numpy as np
from scipy.spatial import cKDTree


def knn(ref_xyz, q_xyz, k, max_dist=np.inf):
    tree = cKDTree(np.asarray(ref_xyz, np.float64))
    d, i = tree.query(np.asarray(q_xyz, np.float64), k=k, distance_upper_bound=max_dist if np.isfinite(max_dist) else np.inf)
    d = d.reshape(len(q_xyz), k)
    i = i.reshape(len(q_xyz), k)
    return np.where(np.isinf(d), -1, i), d ** 2


def quantile_lpm(values, q):
    """LPM Matches::getDistsQuantile: nth_element at index size_t(n * q) (fp32 product), q == 1 -> max."""
    v = np.sort(values[np.isfinite(values)])
    if len(v) == 0:
        raise ValueError("no outlier to filter")
    if q == 1.0:
        return v[-1]
    idx = int(np.float32(len(v)) * np.float32(q))
    return v[min(idx, len(v) - 1)]


def rodrigues(r):
    ang = np.linalg.norm(r)
    if ang == 0:
        return np.eye(3)
    a = r / ang
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def robust_weights(prm, iteration, scale, d2, ids, p_all, ref, map_normals):
    """RobustOutlierFilter (libpointmatcher OutlierFiltersImpl.cpp) written from its documentation in fp64: returns
    (scale, weights).  prm: dict with libpointmatcher's parameter names."""
    prm = dict(dict(robustFct="cauchy", tuning=1.0, scaleEstimator="mad", nbIterationForScale=0, distanceType="point2point",
                    approximation=np.inf), **prm)
    k, est, nb = prm["tuning"], prm["scaleEstimator"], prm["nbIterationForScale"]
    if est == "berg":
        k = {"cauchy": 4.3040, "tukey": 7.0589, "huber": 2.0138}.get(prm["robustFct"], k)
    if iteration <= nb or nb == 0:
        fin = np.sort(d2[np.isfinite(d2)])
        if est == "mad":
            med = fin[len(fin) // 2]
            scale = np.sqrt(np.sort(np.abs(fin - med))[len(fin) // 2])
        elif est == "std":
            scale = np.sqrt(np.std(d2, ddof=1))
        elif est == "berg":
            scale = 1.9 * np.sqrt(fin[int(len(fin) * 0.5)]) if iteration == 1 else 0.85 * (scale - prm["tuning"]) + prm["tuning"]
    s = 1.0 if est == "none" else scale
    dist = d2
    if prm["distanceType"] == "point2plane":
        n = np.asarray(map_normals, np.float64)
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        safe = np.where(np.isfinite(d2), ids, 0)
        diff = p_all[:, None, :] - ref[safe]
        dist = np.einsum("ijk,ijk->ij", n[safe], diff) ** 2
    with np.errstate(all="ignore"):
        e2 = dist / (s * s)
        k2 = k * k
        fct = prm["robustFct"]
        if fct == "cauchy":
            w = 1.0 / (1.0 + e2 / k2)
        elif fct == "welsch":
            w = np.exp(-e2 / k2)
        elif fct == "sc":
            w = np.where(e2 >= k, 4.0 * k2 / (k + e2) ** 2, 1.0)
        elif fct == "gm":
            w = k2 / (k + e2) ** 2
        elif fct == "tukey":
            w = np.where(e2 >= k2, 0.0, (1.0 - e2 / k2) ** 2)
        elif fct == "huber":
            w = np.where(e2 >= k2, k / np.sqrt(e2), 1.0)
        elif fct == "L1":
            w = 1.0 / np.sqrt(e2)
        elif fct == "student":
            w = (1.0 + e2 / k) ** (-(k + 3.0) / 2.0) * (k + 3.0) / (k + e2)
        else:
            raise ValueError(fct)
    w = np.where(w <= 0.0, 0.0, w)
    if np.isfinite(prm["approximation"]):
        w = np.where(e2 >= prm["approximation"] ** 2, 0.0, w)
    return scale, w


def icp(map_xyz, map_normals, reading_xyz, knn_k=1, max_dist=np.inf, outliers=(("trimmed", 0.85),),
        minimizer="point_to_plane", iterations=30, force2D=False, force4DOF=False):
    """Counter-checker-only ICP in the mean-centred frame; returns the (dim+1)x(dim+1) correction."""
    dim = map_xyz.shape[1]
    map_xyz = np.asarray(map_xyz, np.float64)
    mean = map_xyz.mean(axis=0)
    ref = map_xyz - mean
    tree = cKDTree(ref)
    reading = np.asarray(reading_xyz, np.float64) - mean
    T = np.eye(dim + 1)
    robust_scale = 0.0
    for it in range(iterations):
        p_all = reading @ T[:dim, :dim].T + T[:dim, dim]
        d, ids = tree.query(p_all, k=knn_k, distance_upper_bound=max_dist if np.isfinite(max_dist) else np.inf)
        d = d.reshape(len(p_all), knn_k)
        ids = ids.reshape(len(p_all), knn_k)
        d2 = d ** 2
        w = np.ones_like(d2)
        for name, prm in outliers:
            if name == "trimmed":
                w *= d2 <= quantile_lpm(d2.ravel(), prm)
            elif name == "median":
                w *= d2 <= prm * quantile_lpm(d2.ravel(), 0.5)
            elif name == "max_dist":
                w *= d2 <= prm * prm
            elif name == "min_dist":
                w *= d2 >= prm * prm
            elif name == "robust":
                robust_scale, wr = robust_weights(prm, it + 1, robust_scale, d2, ids, p_all, ref, map_normals)
                w *= wr
        keep = np.isfinite(d2) & (w != 0)
        qi, kk = np.nonzero(keep)
        p = p_all[qi]
        q = ref[ids[qi, kk]]
        ww = w[qi, kk]
        dT = np.eye(dim + 1)
        if minimizer == "point_to_plane":
            n = np.asarray(map_normals, np.float64)[ids[qi, kk]]
            resid = np.einsum("ij,ij->i", p - q, n)
            if dim == 3 and force2D:      # LPM PointToPlane.cpp: clouds cut down to x, y
                F = np.c_[p[:, 0] * n[:, 1] - p[:, 1] * n[:, 0], n[:, :2]]
                resid = np.einsum("ij,ij->i", (p - q)[:, :2], n[:, :2])
            elif dim == 3 and force4DOF:  # rotation about z only, full translation
                F = np.c_[p[:, 0] * n[:, 1] - p[:, 1] * n[:, 0], n]
            elif dim == 3:
                F = np.c_[np.cross(p, n), n]
            else:
                F = np.c_[p[:, 0] * n[:, 1] - p[:, 1] * n[:, 0], n]
            A = (F * ww[:, None]).T @ F
            b = -(F * ww[:, None]).T @ resid
            # solvePossiblyUnderdeterminedLinearSystem: the minimum-norm least-squares solution when A is rank deficient
            x = np.linalg.solve(A, b) if np.linalg.matrix_rank(A, tol=1e-6 * np.abs(A).max()) == len(A) else np.linalg.pinv(A, rcond=1e-6) @ b
            if dim == 3 and (force2D or force4DOF):
                c, s = np.cos(x[0]), np.sin(x[0])
                dT[:2, :2] = [[c, -s], [s, c]]
                dT[:len(x) - 1, 3] = x[1:]
            elif dim == 3:
                dT[:3, :3] = rodrigues(x[:3])
                dT[:3, 3] = x[3:]
            else:
                c, s = np.cos(x[0]), np.sin(x[0])
                dT[:2, :2] = [[c, -s], [s, c]]
                dT[:2, 2] = x[1:]
        elif minimizer == "point_to_point":
            mp = (p * ww[:, None]).sum(0) / ww.sum()
            mq = (q * ww[:, None]).sum(0) / ww.sum()
            M = ((q - mq) * ww[:, None]).T @ (p - mp)
            U, _, Vt = np.linalg.svd(M)
            R = U @ Vt
            if np.linalg.det(R) < 0:
                Vt[dim - 1] *= -1
                R = U @ Vt
            dT[:dim, :dim] = R
            dT[:dim, dim] = mq - R @ mp
        T = dT @ T
    Tm = np.eye(dim + 1)
    Tm[:dim, dim] = mean
    Tmi = np.eye(dim + 1)
    Tmi[:dim, dim] = -mean
    return Tm @ T @ Tmi


def surface_normals(xyz, k):
    xyz = np.asarray(xyz, np.float64)
    tree = cKDTree(xyz)
    _, ids = tree.query(xyz, k=k)
    nb = xyz[ids]
    c = nb - nb.mean(axis=1, keepdims=True)
    C = np.einsum("nki,nkj->nij", c, c)
    w, v = np.linalg.eigh(C)
    return v[:, :, 0]
