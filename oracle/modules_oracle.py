"""numpy restatements of the map-update steps whose arithmetic is per-point (no tree search needed).
TEST INFRASTRUCTURE ONLY, parity unpinned (see oracle/icp_oracle.h): the checker for the CUDA map
modules.  Citations are into /root/reference/norlab_icp_mapper unless prefixed LPM (libpointmatcher
1.4.x, restated from its published algorithm).
"""
import numpy as np
from scipy.spatial import cKDTree

f32 = np.float32


def octree_leaf_keys(xyz, max_size_by_node):
    """LPM utils/octree.hpp Octree_::build with maxPointByNode = 1: root box = AABB centre, radius =
    half the largest extent; a node splits while 2 * radius > maxSizeByNode; child octant of a point =
    (p > centre) per axis; child centre = centre +- radius / 2.  All in fp32 like the reference.
    Returns one integer key per point identifying its leaf at the final depth."""
    p = np.asarray(xyz, f32)
    dim = p.shape[1]
    lo, hi = p.min(axis=0), p.max(axis=0)
    radii = (hi - lo).astype(f32)
    centre = (lo + radii * f32(0.5)).astype(f32)
    radius = f32(radii.max() * f32(0.5))
    depth, r = 0, radius
    while float(r) * 2.0 > float(f32(max_size_by_node)):
        r = f32(r * f32(0.5))
        depth += 1
    c = np.tile(centre, (len(p), 1)).astype(f32)
    keys = np.zeros(len(p), np.uint64)
    r = radius
    for _ in range(depth):
        h = f32(r * f32(0.5))
        bits = p > c
        code = np.zeros(len(p), np.uint64)
        for d in range(dim):
            code |= bits[:, d].astype(np.uint64) << np.uint64(d)
        keys = (keys << np.uint64(3)) | code
        c = np.where(bits, c + h, c - h).astype(f32)
        r = h
    return keys, depth


def _mix64(x):
    """splitmix64 finaliser on Python ints (the random sampler's counter-based generator)."""
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def octree_grid_filter(features, max_size_by_node, sampling_method=0, descriptors=None, seed=0x0c7ee5eed):
    """LPM DataPointsFilters/OctreeGrid.cpp with maxPointByNode 1 (ref: OctreeMapperModule.cpp:35-39 after
    map.concatenate).  One survivor per leaf: method 0 = first point (lowest index), 2 = centroid, 1 = random
    member (upstream: its own generator; here member hash(leaf key, seed) mod count of the leaf, members in
    input order), 3 = medoid (member closest to the leaf's centroid, first one on ties).
    Returns (indices of the leaf representatives in input order, survivor features, survivor descriptors)."""
    feat = np.asarray(features, f32)
    dim = feat.shape[1] - 1
    keys, depth = octree_leaf_keys(feat[:, :dim], max_size_by_node)
    uniq, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    if sampling_method in (1, 3):
        members = {}
        for i, k in enumerate(keys):
            members.setdefault(int(k), []).append(i)
        picks = []
        for k, idx in members.items():
            if sampling_method == 1:
                picks.append(idx[_mix64(k ^ _mix64(seed)) % len(idx)])
            else:
                acc = np.zeros(dim, f32)
                for i in idx:  # sequential fp32 sums in index order, like the device's run walk
                    acc = (acc + feat[i, :dim]).astype(f32)
                mean = (acc * f32(f32(1.0) / f32(len(idx)))).astype(f32)
                best, pick = np.inf, idx[0]
                for i in idx:
                    dlt = (feat[i, :dim] - mean).astype(f32)
                    sq = f32(0)
                    for c in range(dim):
                        sq = f32(sq + f32(dlt[c] * dlt[c]))
                    d = np.sqrt(sq, dtype=f32)
                    if d < best:
                        best, pick = d, i
                picks.append(pick)
        order = np.sort(np.array(picks, np.int64))
        out_desc = None if descriptors is None else np.asarray(descriptors, f32)[order].copy()
        return order, feat[order].copy(), out_desc
    order = np.sort(first)
    out_feat = feat[order].copy()
    out_desc = None if descriptors is None else np.asarray(descriptors, f32)[order].copy()
    if sampling_method == 2:
        rep_of_key = {int(k): i for i, k in enumerate(keys[order])}
        slot = np.array([rep_of_key[int(k)] for k in keys])
        cnt = np.bincount(slot, minlength=len(order)).astype(f32)
        for d in range(dim):  # sequential fp32 sums in index order, like the device's run walk
            acc = np.zeros(len(order), f32)
            np.add.at(acc, slot, feat[:, d])
            out_feat[:, d] = acc / cnt
        if descriptors is not None:
            desc = np.asarray(descriptors, f32)
            for d in range(desc.shape[1]):
                acc = np.zeros(len(order), f32)
                np.add.at(acc, slot, desc[:, d])
                out_desc[:, d] = acc / cnt
    return order, out_feat, out_desc


def cut_at_descriptor_threshold(values, threshold, use_larger_than=True):
    """LPM CutAtDescriptorThresholdDataPointsFilter: drop points with value > threshold
    (useLargerThan 1) or < threshold (0).  Returns the keep mask.  ref: examples/config.yaml:29-32."""
    v = np.asarray(values, f32)
    return ~(v > f32(threshold)) if use_larger_than else ~(v < f32(threshold))


def spherical(p, dim):
    """DynamicPointsMapperModule::convertToSphericalCoordinates -- DynamicPointsMapperModule.cpp:156-172."""
    p = np.asarray(p, f32)
    r = np.sqrt((p[:, :dim].astype(f32) ** 2).sum(axis=1, dtype=f32)).astype(f32)
    el = np.arcsin((p[:, 2] / r).astype(f32)).astype(f32) if dim == 3 else np.zeros(len(p), f32)
    az = np.arctan2(p[:, 1], p[:, 0]).astype(f32)
    return r, np.c_[el, az].astype(f32)


def dynamic_points_update(input_map_frame, map_feat, map_normals, map_prob, pose, thresholdDynamic=0.6, alpha=0.8, beta=0.99,
                          beamHalfAngle=0.01, epsilonA=0.01, epsilonD=0.01, sensorMaxRange=200.0):
    """DynamicPointsMapperModule::inPlaceUpdateMap -- DynamicPointsMapperModule.cpp:34-151.
    Returns the updated probabilityDynamic of the map (float32) and the mask of points that found a scan neighbour."""
    dim = map_feat.shape[1] - 1
    eps = f32(0.0001)
    P = np.asarray(pose, np.float64)
    R, t = P[:dim, :dim], P[:dim, dim]
    Rinv = R.T.astype(f32)
    tinv = (-(R.T @ t)).astype(f32)

    def to_sensor(x):
        return (np.asarray(x, f32)[:, :dim] @ Rinv.T + tinv).astype(f32)

    inp = to_sensor(input_map_frame)
    in_r, in_ang = spherical(np.c_[inp, np.zeros((len(inp), 3 - dim), f32)], dim)
    mp = to_sensor(map_feat)
    mp3 = np.c_[mp, np.zeros((len(mp), 3 - dim), f32)]
    m_r, m_ang = spherical(mp3, dim)
    in_range = m_r < f32(sensorMaxRange)
    tree = cKDTree(in_ang.astype(np.float64))
    d, ids = tree.query(m_ang.astype(np.float64), k=1, distance_upper_bound=2 * float(f32(beamHalfAngle)))
    matched = in_range & np.isfinite(d)
    prob = np.asarray(map_prob, f32).copy()
    nrm_s = (np.asarray(map_normals, f32) @ Rinv.T).astype(f32)
    for i in np.nonzero(matched)[0]:
        ip = inp[ids[i]]
        lp = mp[i]
        inN, lpN = f32(np.linalg.norm(ip)), f32(np.linalg.norm(lp))
        delta = f32(np.linalg.norm(ip - lp))
        d_max = f32(f32(epsilonA) * inN)
        dist2 = f32(((m_ang[i] - in_ang[ids[i]]).astype(f32) ** 2).sum())
        w_v = f32(float(eps) + (1. - float(eps)) * abs(float(np.dot(nrm_s[i], lp / lpN))))
        w_d1 = f32(float(eps) + (1. - float(eps)) * (1. - float(np.sqrt(dist2)) / float(f32(2 * f32(beamHalfAngle)))))
        offset = f32(delta - f32(epsilonD))
        w_d2 = f32(1.)
        if delta < f32(epsilonD) or lpN > inN:
            w_d2 = eps
        elif offset < d_max:
            w_d2 = f32(eps + (f32(1) - eps) * offset / d_max)
        w_p2 = eps
        if delta < f32(epsilonD):
            w_p2 = f32(1)
        elif offset < d_max:
            w_p2 = f32(float(eps) + (1. - float(eps)) * (1. - float(f32(offset / d_max))))
        if f32(inN + f32(epsilonD) + d_max) >= lpN:
            last = prob[i]
            c1 = f32(f32(1) - w_v * w_d1)
            c2 = f32(w_v * w_d1)
            if last < f32(thresholdDynamic):
                pd = f32(c1 * last + c2 * w_d2 * (f32(f32(1) - f32(alpha)) * f32(f32(1) - last) + f32(beta) * last))
                ps = f32(c1 * f32(f32(1) - last) + c2 * w_p2 * (f32(alpha) * f32(f32(1) - last) + f32(f32(1) - f32(beta)) * last))
            else:
                pd, ps = f32(f32(1) - eps), eps
            prob[i] = f32(pd / f32(pd + ps))
    return prob, matched


def bounding_box_keep(features, lo, hi, remove_inside=True):
    """LPM BoundingBoxDataPointsFilter{xMin..zMax, removeInside}: a point is inside when every
    coordinate lies in [min, max]; ref: examples/config.yaml:1-17."""
    p = np.asarray(features, f32)
    dim = len(lo)
    inside = np.all((p[:, :dim] >= np.asarray(lo, f32)) & (p[:, :dim] <= np.asarray(hi, f32)), axis=1)
    return ~inside if remove_inside else inside


def distance_limit_keep(features, dist, dim_index=-1, remove_inside=False, n_dim=3):
    """LPM DistanceLimitDataPointsFilter{dim, dist, removeInside}: dim -1 -> radial test on the norm;
    removeInside 0 keeps ||p|| < dist (this is Mapper's radiusFilter, Mapper.cpp:27-31)."""
    p = np.asarray(features, f32)
    if dim_index == -1:
        v = np.sqrt((p[:, :n_dim] ** 2).sum(axis=1, dtype=f32)).astype(f32)
        lim = abs(f32(dist))
    else:
        v = p[:, dim_index]
        lim = f32(dist)
    return (v > lim) if remove_inside else (v < lim)


def random_sampling_keep(n, prob, seed=0, slot=0):
    """RandomSamplingDataPointsFilter{prob} as this project defines its generator (upstream: rand() < prob * RAND_MAX, any
    stream): point i of the cloud entering the chain survives when u < prob, u = top 24 bits of
    splitmix64((seed << 40) ^ (chain slot << 36) ^ i) / 2^24."""
    keep = np.zeros(n, bool)
    p = f32(prob)
    for i in range(n):
        x = _mix64(((seed & 0xFFFFFFFF) << 40) ^ (slot << 36) ^ i)
        keep[i] = f32(x >> 40) * f32(1.0 / 16777216.0) < p
    return keep
