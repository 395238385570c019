"""Disk formats and host-side containers of the reference's Python surface (python/src/*.cpp), on top of the C++ mirror's VTK
reader / writer (host/IO.h through libb200mapper.so; no GPU involved):

  save_vtk / load_vtk       PointMatcher DataPoints::save / ::load for legacy VTK POLYDATA (ASCII or binary)
  RamCellManager            python/src/ram_cell_manager.cpp   (RAMCellManager.cpp)
  HardDriveCellManager      python/src/hard_drive_cell_manager.cpp (HardDriveCellManager.cpp: one cell_<id>.vtk per cell)
  Trajectory                python/src/trajectory.cpp (Trajectory.cpp: positions + orientation columns, saved as a point cloud)

A cloud is a dict: features (n, dim + 1) fp32 with the homogeneous 1, optional normals (n, dim), optional probabilityDynamic (n,).
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._lib import B200ICPError


def _L():
    from . import mapper
    return mapper.load()


def save_vtk(path, features, normals=None, probabilityDynamic=None, binary=False):
    L = _L()
    features = np.ascontiguousarray(features, np.float32)
    n, rows = features.shape
    nptr = pptr = None
    if normals is not None:
        normals = np.ascontiguousarray(normals, np.float32)
        if normals.shape != (n, rows - 1):
            raise ValueError("normals must be (n, dim)")
        nptr = normals.ctypes.data
    if probabilityDynamic is not None:
        probabilityDynamic = np.ascontiguousarray(probabilityDynamic, np.float32).reshape(-1)
        if len(probabilityDynamic) != n:
            raise ValueError("probabilityDynamic must have one value per point")
        pptr = probabilityDynamic.ctypes.data
    rc = L.b200mapper_vtk_save(os.fspath(path).encode(), features.ctypes.data, rows, n, nptr, pptr, int(binary))
    if rc != _abi.OK:
        raise B200ICPError(rc, L.b200mapper_last_error(None).decode())


def load_vtk(path, dim=3):
    """-> dict(features, normals or None, probabilityDynamic or None)."""
    L = _L()
    p = os.fspath(path).encode()
    n, has_n, has_p = C.c_int64(), C.c_int32(), C.c_int32()
    rc = L.b200mapper_vtk_load(p, dim, None, None, None, 0, C.byref(n), C.byref(has_n), C.byref(has_p))
    if rc != _abi.OK:
        raise B200ICPError(rc, L.b200mapper_last_error(None).decode())
    feat = np.zeros((n.value, dim + 1), np.float32)
    nrm = np.zeros((n.value, dim), np.float32) if has_n.value else None
    prob = np.zeros(n.value, np.float32) if has_p.value else None
    if n.value:
        rc = L.b200mapper_vtk_load(p, dim, feat.ctypes.data, nrm.ctypes.data if nrm is not None else None,
                                   prob.ctypes.data if prob is not None else None, n.value, C.byref(n), C.byref(has_n), C.byref(has_p))
        if rc != _abi.OK:
            raise B200ICPError(rc, L.b200mapper_last_error(None).decode())
    return dict(features=feat, normals=nrm, probabilityDynamic=prob)


def _empty(dim):
    return dict(features=np.zeros((0, dim + 1), np.float32), normals=None, probabilityDynamic=None)


class RamCellManager:
    """CellManager interface (CellManager.h:15-18) over a dict, as RAMCellManager.cpp: saveCell overwrites, retrieveCell of an
    unknown id returns an empty cloud."""

    def __init__(self, dim=3):
        self.dim = dim
        self._cells = {}

    def getAllCellIds(self):
        return list(self._cells)

    def saveCell(self, cellId, cell):
        self._cells[str(cellId)] = {k: (None if v is None else np.array(v, np.float32, copy=True)) for k, v in cell.items()}

    def retrieveCell(self, cellId):
        c = self._cells.get(str(cellId))
        return _empty(self.dim) if c is None else {k: (None if v is None else v.copy()) for k, v in c.items()}

    def clearAllCells(self):
        self._cells.clear()


class HardDriveCellManager:
    """One VTK file per cell, `<folder>cell_<id>.vtk` (HardDriveCellManager.h / .cpp:14-36); the folder defaults to /tmp/ as upstream."""

    def __init__(self, dim=3, folder="/tmp/", binary=True):
        self.dim = dim
        self.folder = os.fspath(folder)
        if not self.folder.endswith(os.sep):
            self.folder += os.sep
        self.binary = binary
        self._ids = set()

    def _path(self, cellId):
        return f"{self.folder}cell_{cellId}.vtk"

    def getAllCellIds(self):
        return sorted(self._ids)

    def saveCell(self, cellId, cell):
        save_vtk(self._path(cellId), cell["features"], cell.get("normals"), cell.get("probabilityDynamic"), binary=self.binary)
        self._ids.add(str(cellId))

    def retrieveCell(self, cellId):
        if str(cellId) not in self._ids:
            return _empty(self.dim)
        return load_vtk(self._path(cellId), self.dim)

    def clearAllCells(self):
        for cellId in self._ids:
            try:
                os.remove(self._path(cellId))
            except FileNotFoundError:
                pass
        self._ids.clear()


class Trajectory:
    """Trajectory.cpp: poses with time stamps; save() writes them as a point cloud -- positions as the points, the columns of the
    rotation as the point's normals-like orientation (upstream stores orientationX / Y / Z descriptors; this writer keeps the
    first column, the heading, as `normals`) -- through the same VTK writer."""

    def __init__(self, dimension=3):
        self.dimension = dimension
        self.poses = []
        self.timeStamps = []

    def addPose(self, pose, timeStamp):
        pose = np.asarray(pose, np.float32)
        if pose.shape != (self.dimension + 1, self.dimension + 1):
            raise ValueError("pose must be (dim + 1) x (dim + 1)")
        self.poses.append(pose.copy())
        self.timeStamps.append(float(timeStamp))

    def save(self, filename, binary=False):
        d = self.dimension
        n = len(self.poses)
        feat = np.ones((n, d + 1), np.float32)
        heading = np.zeros((n, d), np.float32)
        for i, P in enumerate(self.poses):
            feat[i, :d] = P[:d, d]
            heading[i] = P[:d, 0]
        save_vtk(filename, feat, heading if n else None, None, binary=binary)

    def clear(self):
        self.poses.clear()
        self.timeStamps.clear()
