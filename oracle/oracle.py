"""ctypes wrapper around oracle/libpcr_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference path (see pcr_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under pointclouds_rs_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpcr_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle (gcc, a second or two)."""
    src = os.path.join(_HERE, "pcr_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "pcr_oracle.h"))
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "all"])
    return _LIB_PATH


_lib = None

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class _Transform(C.Structure):
    _fields_ = [("rotation", C.c_float * 9), ("translation", C.c_float * 3)]


class _IcpResult(C.Structure):
    _fields_ = [
        ("transform", _Transform),
        ("fitness", C.c_float),
        ("rmse", C.c_float),
        ("converged", C.c_int),
        ("num_iterations", C.c_size_t),
    ]


class _IcpParams(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_size_t),
        ("tolerance", C.c_float),
        ("max_correspondence_distance", C.c_float),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_tree_build.restype = C.c_void_p
        L.orc_tree_build.argtypes = [_f32p, _f32p, _f32p, C.c_size_t]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_tree_len.restype = C.c_size_t
        L.orc_tree_len.argtypes = [C.c_void_p]
        L.orc_tree_knn.restype = C.c_size_t
        L.orc_tree_knn.argtypes = [C.c_void_p, _f32p, C.c_size_t, _u32p, _f32p]
        L.orc_knn_brute.restype = C.c_size_t
        L.orc_knn_brute.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, _f32p, C.c_size_t, _u32p, _f32p]
        L.orc_knn_batch.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, _u32p, _f32p, _u32p, C.c_int]
        L.orc_tree_radius_search.restype = C.c_size_t
        L.orc_tree_radius_search.argtypes = [C.c_void_p, _f32p, C.c_float, _u32p, C.c_size_t]
        L.orc_tree_radius_count.restype = C.c_size_t
        L.orc_tree_radius_count.argtypes = [C.c_void_p, _f32p, C.c_float]
        L.orc_radius_count_batch.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_size_t, C.c_float, _u32p, C.c_int]
        L.orc_sor.restype = C.c_size_t
        L.orc_sor.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_float, _u8p, _f32p, _f32p, C.c_int]
        L.orc_ror.restype = C.c_size_t
        L.orc_ror.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_float, C.c_size_t, _u8p, C.c_int]
        L.orc_normals.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, _f32p, _f32p, _f32p, _f32p, C.c_int]
        L.orc_smallest_eigenvector_3x3.argtypes = [C.c_float] * 6 + [_f32p]
        L.orc_apply_transform.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.POINTER(_Transform), _f32p, _f32p, _f32p]
        L.orc_compose.argtypes = [C.POINTER(_Transform)] * 3
        L.orc_find_correspondences.restype = C.c_size_t
        L.orc_find_correspondences.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_void_p, C.c_float, _u32p, _u32p, _f32p, C.c_int]
        L.orc_icp_point_to_point.argtypes = [_f32p] * 3 + [C.c_size_t] + [_f32p] * 3 + [C.c_size_t, C.POINTER(_IcpParams), C.POINTER(_IcpResult), C.c_int]
        L.orc_icp_point_to_plane.restype = C.c_int
        L.orc_icp_point_to_plane.argtypes = [_f32p] * 3 + [C.c_size_t] + [_f32p] * 3 + [C.c_size_t] + [_f32p] * 3 + [C.c_size_t, C.POINTER(_IcpParams), C.POINTER(_IcpResult), C.c_int]
        for fn in (L.orc_euclidean_cluster, L.orc_euclidean_cluster_brute):
            fn.restype = C.c_size_t
            fn.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_float, C.c_size_t, C.c_size_t, _u32p, _u32p]
        L.orc_ransac_plane_samples.restype = C.c_size_t
        L.orc_ransac_plane_samples.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_float, _u32p, C.c_size_t, _f32p, _u32p]
        L.orc_voxel_downsample.restype = C.c_size_t
        L.orc_voxel_downsample.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_float, _f32p, _f32p, _f32p]
        L.orc_read_pcd_ascii.restype = C.c_long
        L.orc_read_pcd_ascii.argtypes = [C.c_char_p, _f32p, _f32p, _f32p, C.c_size_t]
        L.orc_grid_knn_model.argtypes = [_f32p] * 3 + [C.c_size_t] + [_f32p] * 3 + [C.c_size_t, C.c_size_t, C.c_float, _u32p, _f32p, _u32p, _f64p]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _xyz(pts):
    """(N,3) array -> three contiguous SoA f32 arrays (crates/core/src/cloud.rs:53-71)."""
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 3)
    x = np.ascontiguousarray(pts[:, 0])
    y = np.ascontiguousarray(pts[:, 1])
    z = np.ascontiguousarray(pts[:, 2])
    return x, y, z


def _p(a, t):
    return a.ctypes.data_as(t)


class Tree:
    """pointclouds_spatial::KdTree (crates/spatial/src/kdtree.rs:14-164)."""

    def __init__(self, pts):
        self.x, self.y, self.z = _xyz(pts)
        self.n = len(self.x)
        self._h = lib().orc_tree_build(_p(self.x, _f32p), _p(self.y, _f32p), _p(self.z, _f32p), self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_free(self._h)
            self._h = None

    def __len__(self):
        return lib().orc_tree_len(self._h)

    def knn(self, q, k):
        qa = np.asarray(q, dtype=np.float32).reshape(3)
        idx = np.empty(max(k, 1), np.uint32)
        dist = np.empty(max(k, 1), np.float32)
        r = lib().orc_tree_knn(self._h, _p(qa, _f32p), k, _p(idx, _u32p), _p(dist, _f32p))
        return idx[:r].copy(), dist[:r].copy()

    def knn_batch(self, queries, k, threads=1):
        qx, qy, qz = _xyz(queries)
        nq = len(qx)
        idx = np.empty((nq, k), np.uint32)
        dist = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint32)
        lib().orc_knn_batch(self._h, _p(qx, _f32p), _p(qy, _f32p), _p(qz, _f32p), nq, k, _p(idx, _u32p), _p(dist, _f32p), _p(cnt, _u32p), threads)
        return idx, dist, cnt

    def radius_search(self, q, radius):
        qa = np.asarray(q, dtype=np.float32).reshape(3)
        cap = max(self.n, 1)
        idx = np.empty(cap, np.uint32)
        r = lib().orc_tree_radius_search(self._h, _p(qa, _f32p), float(radius), _p(idx, _u32p), cap)
        return idx[:r].copy()

    def radius_count_batch(self, queries, radius, threads=1):
        qx, qy, qz = _xyz(queries)
        cnt = np.empty(len(qx), np.uint32)
        lib().orc_radius_count_batch(self._h, _p(qx, _f32p), _p(qy, _f32p), _p(qz, _f32p), len(qx), float(radius), _p(cnt, _u32p), threads)
        return cnt

    def find_correspondences(self, source, max_distance=np.inf, threads=1):
        sx, sy, sz = _xyz(source)
        ns = len(sx)
        si = np.empty(max(ns, 1), np.uint32)
        ti = np.empty(max(ns, 1), np.uint32)
        dd = np.empty(max(ns, 1), np.float32)
        m = lib().orc_find_correspondences(_p(sx, _f32p), _p(sy, _f32p), _p(sz, _f32p), ns, self._h, float(max_distance), _p(si, _u32p), _p(ti, _u32p), _p(dd, _f32p), threads)
        return si[:m].copy(), ti[:m].copy(), dd[:m].copy()


def knn_brute(pts, q, k):
    x, y, z = _xyz(pts)
    qa = np.asarray(q, dtype=np.float32).reshape(3)
    idx = np.empty(max(k, 1), np.uint32)
    dist = np.empty(max(k, 1), np.float32)
    r = lib().orc_knn_brute(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), len(x), _p(qa, _f32p), k, _p(idx, _u32p), _p(dist, _f32p))
    return idx[:r].copy(), dist[:r].copy()


def grid_knn_model(pts, queries, k, cell):
    """Scalar model of the CUDA engine's ring search (same results as Tree.knn_batch)."""
    x, y, z = _xyz(pts)
    qx, qy, qz = _xyz(queries)
    nq = len(qx)
    idx = np.empty((nq, max(k, 1)), np.uint32)
    dist = np.empty((nq, max(k, 1)), np.float32)
    cnt = np.empty(nq, np.uint32)
    stats = np.zeros(4, np.float64)
    lib().orc_grid_knn_model(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), len(x), _p(qx, _f32p), _p(qy, _f32p), _p(qz, _f32p), nq, k, float(cell), _p(idx, _u32p), _p(dist, _f32p), _p(cnt, _u32p), _p(stats, _f64p))
    return idx[:, :k], dist[:, :k], cnt, stats


def sor(pts, k, std_mul, threads=1):
    """statistical_outlier_removal -> (keep mask u8, mean_d, (mean, std, thr))."""
    x, y, z = _xyz(pts)
    n = len(x)
    keep = np.zeros(max(n, 1), np.uint8)
    md = np.full(max(n, 1), np.inf, np.float32)
    st = np.zeros(3, np.float32)
    kept = lib().orc_sor(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, k, float(std_mul), _p(keep, _u8p), _p(md, _f32p), _p(st, _f32p), threads)
    assert kept == int(keep[:n].sum())
    return keep[:n].copy(), md[:n].copy(), st


def ror(pts, radius, min_neighbors, threads=1):
    x, y, z = _xyz(pts)
    n = len(x)
    keep = np.zeros(max(n, 1), np.uint8)
    lib().orc_ror(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, float(radius), min_neighbors, _p(keep, _u8p), threads)
    return keep[:n].copy()


def normals(pts, k, viewpoint=(0.0, 0.0, 0.0), threads=1):
    x, y, z = _xyz(pts)
    n = len(x)
    if n == 0 or k == 0:
        return np.zeros((0, 3), np.float32)
    vp = np.asarray(viewpoint, np.float32)
    nx = np.empty(n, np.float32)
    ny = np.empty(n, np.float32)
    nz = np.empty(n, np.float32)
    lib().orc_normals(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, k, _p(vp, _f32p), _p(nx, _f32p), _p(ny, _f32p), _p(nz, _f32p), threads)
    return np.stack([nx, ny, nz], axis=1)


def smallest_eigenvector(a00, a01, a02, a11, a12, a22):
    out = np.empty(3, np.float32)
    lib().orc_smallest_eigenvector_3x3(a00, a01, a02, a11, a12, a22, _p(out, _f32p))
    return out


def _mk_transform(rotation, translation):
    t = _Transform()
    r = np.asarray(rotation, np.float32).reshape(9)
    tr = np.asarray(translation, np.float32).reshape(3)
    for i in range(9):
        t.rotation[i] = float(r[i])
    for i in range(3):
        t.translation[i] = float(tr[i])
    return t


def apply_transform(pts, rotation, translation):
    x, y, z = _xyz(pts)
    n = len(x)
    ox = np.empty(n, np.float32)
    oy = np.empty(n, np.float32)
    oz = np.empty(n, np.float32)
    t = _mk_transform(rotation, translation)
    lib().orc_apply_transform(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, C.byref(t), _p(ox, _f32p), _p(oy, _f32p), _p(oz, _f32p))
    return np.stack([ox, oy, oz], axis=1)


def compose(first, then):
    """RigidTransform::compose: apply `first`, then `then` (icp.rs:52-73). Args: (R, t) pairs."""
    a = _mk_transform(*first)
    b = _mk_transform(*then)
    o = _Transform()
    lib().orc_compose(C.byref(a), C.byref(b), C.byref(o))
    return np.array(o.rotation, np.float32).reshape(3, 3), np.array(o.translation, np.float32)


@dataclass
class IcpResult:
    rotation: np.ndarray
    translation: np.ndarray
    fitness: float
    rmse: float
    converged: bool
    num_iterations: int


def _icp_out(res):
    return IcpResult(
        rotation=np.array(res.transform.rotation, np.float32).reshape(3, 3),
        translation=np.array(res.transform.translation, np.float32),
        fitness=float(np.float32(res.fitness)),
        rmse=float(np.float32(res.rmse)),
        converged=bool(res.converged),
        num_iterations=int(res.num_iterations),
    )


def icp_point_to_point(source, target, max_iterations=50, tolerance=1e-5, max_correspondence_distance=np.inf, threads=1):
    sx, sy, sz = _xyz(source)
    tx, ty, tz = _xyz(target)
    p = _IcpParams(max_iterations, tolerance, max_correspondence_distance)
    res = _IcpResult()
    lib().orc_icp_point_to_point(_p(sx, _f32p), _p(sy, _f32p), _p(sz, _f32p), len(sx), _p(tx, _f32p), _p(ty, _f32p), _p(tz, _f32p), len(tx), C.byref(p), C.byref(res), threads)
    return _icp_out(res)


def icp_point_to_plane(source, target, target_normals, max_iterations=50, tolerance=1e-5, max_correspondence_distance=np.inf, threads=1):
    sx, sy, sz = _xyz(source)
    tx, ty, tz = _xyz(target)
    nx, ny, nz = _xyz(target_normals)
    p = _IcpParams(max_iterations, tolerance, max_correspondence_distance)
    res = _IcpResult()
    rc = lib().orc_icp_point_to_plane(_p(sx, _f32p), _p(sy, _f32p), _p(sz, _f32p), len(sx), _p(tx, _f32p), _p(ty, _f32p), _p(tz, _f32p), len(tx), _p(nx, _f32p), _p(ny, _f32p), _p(nz, _f32p), len(nx), C.byref(p), C.byref(res), threads)
    if rc != 0:
        raise ValueError(
            f"target_normals length ({len(nx)}) does not match target cloud length ({len(tx)})"
        )
    return _icp_out(res)


def euclidean_cluster(pts, distance_threshold, min_size, max_size, brute=False):
    """euclidean_cluster.rs:96-187 -> list of index arrays (reference order).  brute=True: the O(n^2)
    checker of tests/cluster_differential.rs."""
    x, y, z = _xyz(pts)
    n = len(x)
    off = np.zeros(n + 1, np.uint32)
    idx = np.zeros(max(n, 1), np.uint32)
    fn = lib().orc_euclidean_cluster_brute if brute else lib().orc_euclidean_cluster
    nc = fn(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, float(distance_threshold), int(min_size), int(max_size), _p(off, _u32p), _p(idx, _u32p))
    return [idx[off[c]:off[c + 1]].copy() for c in range(nc)]


def ransac_plane_samples(pts, threshold, samples):
    """ransac_plane.rs:56-129 for given (m,3) sample triples -> (model [nx,ny,nz,d] f32, inlier indices u32)."""
    x, y, z = _xyz(pts)
    n = len(x)
    smp = np.ascontiguousarray(np.asarray(samples, np.uint32).reshape(-1, 3))
    model = np.zeros(4, np.float32)
    inl = np.zeros(max(n, 1), np.uint32)
    k = lib().orc_ransac_plane_samples(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, float(threshold), _p(smp, _u32p), len(smp), _p(model, _f32p), _p(inl, _u32p))
    return model, inl[:k].copy()


def voxel_downsample(pts, voxel_size):
    x, y, z = _xyz(pts)
    n = len(x)
    ox = np.empty(max(n, 1), np.float32)
    oy = np.empty(max(n, 1), np.float32)
    oz = np.empty(max(n, 1), np.float32)
    m = lib().orc_voxel_downsample(_p(x, _f32p), _p(y, _f32p), _p(z, _f32p), n, float(voxel_size), _p(ox, _f32p), _p(oy, _f32p), _p(oz, _f32p))
    if m == C.c_size_t(-1).value:
        raise ValueError("voxel_size must be > 0 and finite")
    return np.stack([ox[:m], oy[:m], oz[:m]], axis=1)


def read_pcd_ascii(path, cap=1 << 22):
    x = np.empty(cap, np.float32)
    y = np.empty(cap, np.float32)
    z = np.empty(cap, np.float32)
    n = lib().orc_read_pcd_ascii(path.encode(), _p(x, _f32p), _p(y, _f32p), _p(z, _f32p), cap)
    if n < 0:
        raise IOError(path)
    return np.stack([x[:n], y[:n], z[:n]], axis=1)
