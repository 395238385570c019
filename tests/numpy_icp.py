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


def icp(map_xyz, map_normals, reading_xyz, knn_k=1, max_dist=np.inf, outliers=(("trimmed", 0.85),),
        minimizer="point_to_plane", iterations=30):
    """Counter-checker-only ICP in the mean-centred frame; returns the (dim+1)x(dim+1) correction."""
    dim = map_xyz.shape[1]
    map_xyz = np.asarray(map_xyz, np.float64)
    mean = map_xyz.mean(axis=0)
    ref = map_xyz - mean
    tree = cKDTree(ref)
    reading = np.asarray(reading_xyz, np.float64) - mean
    T = np.eye(dim + 1)
    for _ in range(iterations):
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
        keep = np.isfinite(d2) & (w != 0)
        qi, kk = np.nonzero(keep)
        p = p_all[qi]
        q = ref[ids[qi, kk]]
        ww = w[qi, kk]
        dT = np.eye(dim + 1)
        if minimizer == "point_to_plane":
            n = np.asarray(map_normals, np.float64)[ids[qi, kk]]
            if dim == 3:
                F = np.c_[np.cross(p, n), n]
            else:
                F = np.c_[p[:, 0] * n[:, 1] - p[:, 1] * n[:, 0], n]
            A = (F * ww[:, None]).T @ F
            b = -(F * ww[:, None]).T @ np.einsum("ij,ij->i", p - q, n)
            x = np.linalg.solve(A, b)
            if dim == 3:
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
