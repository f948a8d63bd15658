"""pointclouds_rs_b200 -- B200-native drop-in for the KNN hot path of pointclouds-rs ("pcrs").

The module mirrors the names and argument meaning of the reference's PyO3 module `pointclouds_rs`
(crates/python/src/lib.rs:12-49) for the functions that sit on the KNN path:

    PointCloud                                    crates/python/src/cloud.rs:10-88
    statistical_outlier_removal(cloud, k, std_mul)      crates/python/src/filters.rs:36-49
    radius_outlier_removal(cloud, radius, min_neighbors) crates/python/src/filters.rs:51-66
    estimate_normals(cloud, k)                          crates/python/src/normals.rs:4-10
    icp_point_to_point / icp_point_to_plane / apply_transform / IcpResult
                                                        crates/python/src/registration.rs:4-109
    KdTree (batched)                                    crates/spatial/src/kdtree.rs:14-164

and for the steps either side of it:

    voxel_downsample(cloud, voxel_size)                 crates/python/src/filters.rs (voxel_downsample.rs:12-65)
    euclidean_cluster(cloud, threshold, min, max)       crates/segmentation/src/euclidean_cluster.rs:96-187
    ransac_plane(cloud, threshold, iterations)          crates/segmentation/src/ransac_plane.rs:56-129 (after the sampling step)
    DeviceCloud                                         a PointCloud that stays in HBM between the steps of a pipeline

Everything runs through the C ABI (include/pcr_b200.h) of lib/libpcr_b200.so: hand-written sm_100a
CUDA, no PyTorch on the path, no CPU fallback.  Out of scope: passthrough_filter and file IO.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import PcrError  # noqa: F401

__all__ = [
    "PointCloud", "KdTree", "IcpResult", "Context",
    "statistical_outlier_removal", "radius_outlier_removal", "estimate_normals",
    "icp_point_to_point", "icp_point_to_plane", "apply_transform", "find_correspondences",
    "sor_normals_batch", "default_context", "PcrError", "euclidean_cluster", "voxel_downsample", "DeviceCloud", "ransac_plane", "PlaneResult",
]


def _p(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(t) if a is not None else None


class Context:
    """Owns a pcr_ctx (device, stream, scratch).  One per host thread."""

    def __init__(self, device: Optional[int] = None, stream: Optional[int] = None):
        lib = _ffi.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            if lib.pcr_device_count() and device >= lib.pcr_device_count():
                device = 0
        h = C.c_void_p()
        if stream is None:
            st = lib.pcr_ctx_create(device, C.byref(h))
        else:
            st = lib.pcr_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        _ffi.check(st)
        self._h = h
        self.device = device
        self._clouds = weakref.WeakSet()  # live DeviceClouds: they must be freed before the context is destroyed

    def close(self):
        if getattr(self, "_h", None):
            for cl in list(getattr(self, "_clouds", ())):
                cl.free()
            _ffi.load().pcr_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _ffi.check(_ffi.load().pcr_ctx_synchronize(self._h), self._h)

    @property
    def launch_count(self) -> int:
        return int(_ffi.load().pcr_ctx_launch_count(self._h))

    def set_timing(self, enable: bool):
        _ffi.check(_ffi.load().pcr_ctx_set_timing(self._h, 1 if enable else 0), self._h)

    def get_timing(self):
        """-> {tag: (milliseconds, spans)} accumulated since the previous call (synchronises)."""
        ms = (C.c_double * _ffi.PCR_NUM_TIMING_TAGS)()
        cnt = (C.c_uint64 * _ffi.PCR_NUM_TIMING_TAGS)()
        _ffi.check(_ffi.load().pcr_ctx_get_timing(self._h, ms, cnt), self._h)
        return {_ffi.TIMING_TAGS[i]: (ms[i], int(cnt[i])) for i in range(_ffi.PCR_NUM_TIMING_TAGS)}

    def set_frame_stream(self, enable: bool = True):
        """Consecutive clouds are frames of one sensor stream: reuse the probed cell size between similar frames."""
        _ffi.check(_ffi.load().pcr_ctx_set_frame_stream(self._h, 1 if enable else 0), self._h)

    def knn_counters(self) -> dict:
        """{queries, candidates}: work of the level-0 KNN kernel since the previous call (counted while timing is on)."""
        out = (C.c_uint64 * 2)()
        _ffi.check(_ffi.load().pcr_ctx_get_knn_counters(self._h, out), self._h)
        return {"queries": int(out[0]), "candidates": int(out[1])}

    def hint_stats(self) -> dict:
        """How often the frame-stream hints held (pcr_ctx_get_hint_stats)."""
        out = (C.c_uint64 * 8)()
        _ffi.check(_ffi.load().pcr_ctx_get_hint_stats(self._h, out), self._h)
        names = ("cell_size_reused", "cell_size_probed", "voxel_box_guess_held", "voxel_box_guess_missed", "coarser_level_ahead_needed",
                 "coarser_level_ahead_unneeded", "deferred_counts_not_awaited_zero", "deferred_counts_not_awaited_nonzero")
        return {k: int(v) for k, v in zip(names, out)}

    def set_query_sharding(self, enable: bool = True):
        """With a communicator (comm_init): SOR / normals / radius outlier removal of ONE cloud, given in full on every rank,
        search only this rank's share of the queries and merge the results over NCCL."""
        _ffi.check(_ffi.load().pcr_ctx_set_query_sharding(self._h, 1 if enable else 0), self._h)

    def debug_set_shard(self, rank: int, world_size: int):
        """Test hook (pcr_ctx_debug_set_shard): play one rank of a query-sharded call without a communicator."""
        _ffi.check(_ffi.load().pcr_ctx_debug_set_shard(self._h, int(rank), int(world_size)), self._h)

    def set_cell_size(self, cell: float):
        _ffi.check(_ffi.load().pcr_ctx_set_cell_size(self._h, float(cell)), self._h)

    # multi-GPU: the host exchanges the id (e.g. torch.distributed.broadcast_object_list)
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_ffi.PCR_UNIQUE_ID_BYTES)
        _ffi.check(_ffi.load().pcr_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world_size: int):
        buf = C.create_string_buffer(uid, _ffi.PCR_UNIQUE_ID_BYTES)
        _ffi.check(_ffi.load().pcr_ctx_comm_init(self._h, buf, rank, world_size), self._h)

    @property
    def comm_kind(self) -> str:
        """How the sharded ICP all-reduces its normal equations on this context: 'none', 'nccl' or 'peer' (NVLink peer memory)."""
        return ("none", "nccl", "peer")[_ffi.load().pcr_ctx_comm_kind(self._h)]


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def _soa(a: np.ndarray):
    x = np.ascontiguousarray(a[:, 0])
    y = np.ascontiguousarray(a[:, 1])
    z = np.ascontiguousarray(a[:, 2])
    return x, y, z


class PointCloud:
    """SoA point cloud (crates/core/src/cloud.rs:4-11) with the PyO3 class's surface."""

    def __init__(self):
        self.x = np.empty(0, np.float32)
        self.y = np.empty(0, np.float32)
        self.z = np.empty(0, np.float32)
        self.normals: Optional[np.ndarray] = None  # (N,3) f32
        self._nsoa = None  # (the array `normals` was when split, (nx, ny, nz)): the reference's Normals are SoA (cloud.rs:13-18)

    @staticmethod
    def _from_xyz(x, y, z, normals=None) -> "PointCloud":
        pc = PointCloud()
        pc.x, pc.y, pc.z = x, y, z
        pc.normals = normals
        return pc

    @staticmethod
    def from_numpy(array) -> "PointCloud":
        # crates/python/src/cloud.rs:25-37,91-137
        if not isinstance(array, np.ndarray) or array.dtype not in (np.float32, np.float64):
            raise TypeError("expected NumPy array with dtype float32 or float64, shape (N, 3)")
        if array.ndim != 2:
            raise TypeError("expected NumPy array with dtype float32 or float64, shape (N, 3)")
        if not array.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous (row-major). Use numpy.ascontiguousarray(arr) to convert.")
        if array.shape[1] != 3:
            raise ValueError("expected shape (N, 3)")
        a = array.astype(np.float32, copy=False)
        return PointCloud._from_xyz(*_soa(a))

    def _normals_soa(self):
        """nx, ny, nz as contiguous arrays, split once per normals array (a 1 M-point split costs more than an ICP setup)."""
        if self._nsoa is None or self._nsoa[0] is not self.normals:
            self._nsoa = (self.normals, _soa(np.asarray(self.normals, np.float32).reshape(-1, 3)))
        return self._nsoa[1]

    def to_numpy(self) -> np.ndarray:
        return np.stack([self.x, self.y, self.z], axis=1) if len(self.x) else np.zeros((0, 3), np.float32)

    def normals_to_numpy(self) -> Optional[np.ndarray]:
        """Superset of the reference surface (it has no accessor for normals)."""
        return None if self.normals is None else self.normals.copy()

    def len(self) -> int:
        return int(len(self.x))

    def is_empty(self) -> bool:
        return len(self.x) == 0

    def __len__(self) -> int:
        return int(len(self.x))

    def __repr__(self) -> str:
        return f"PointCloud(n={len(self.x)})"

    def select(self, indices: Sequence[int]) -> "PointCloud":
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        n = len(self.x)
        bad = idx[(idx >= n) | (idx < 0)]
        if len(bad):
            raise IndexError(f"index {int(bad[0])} out of bounds for cloud with {n} points")
        nr = None if self.normals is None else self.normals[idx]
        return PointCloud._from_xyz(self.x[idx], self.y[idx], self.z[idx], nr)

    def select_inverse(self, indices: Sequence[int]) -> "PointCloud":
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        n = len(self.x)
        bad = idx[(idx >= n) | (idx < 0)]
        if len(bad):
            raise IndexError(f"index {int(bad[0])} out of bounds for cloud with {n} points")
        keep = np.ones(n, bool)
        keep[idx] = False
        return self.select(np.nonzero(keep)[0])


class KdTree:
    """Batched form of pointclouds_spatial::KdTree: a uniform-grid index resident in HBM."""

    def __init__(self, cloud: PointCloud, k_hint: int = 0, ctx: Optional[Context] = None):
        self._ctx = ctx or default_context()
        lib = _ffi.load()
        h = C.c_void_p()
        self._n = len(cloud)
        st = lib.pcr_index_build(self._ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p),
                                 self._n, k_hint, C.byref(h))
        _ffi.check(st, self._ctx._h)
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _ffi.load().pcr_index_free(self._h)
                self._h = None
        except Exception:
            pass

    def len(self) -> int:
        return int(_ffi.load().pcr_index_len(self._h))

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def info(self):
        cell = C.c_float()
        dims = (C.c_int32 * 3)()
        n_idx = C.c_size_t()
        _ffi.check(_ffi.load().pcr_index_info(self._h, C.byref(cell), dims, C.byref(n_idx)))
        return {"cell_size": cell.value, "dims": list(dims), "n_indexed": n_idx.value}

    def knn(self, queries: np.ndarray, k: int):
        """-> (idx (nq,k) u32, dist (nq,k) f32, counts (nq,) u32); rows padded with UINT32_MAX / inf."""
        q = np.ascontiguousarray(np.asarray(queries, np.float32).reshape(-1, 3))
        qx, qy, qz = _soa(q)
        nq = len(qx)
        idx = np.full((nq, max(k, 0)), 0xFFFFFFFF, np.uint32)
        dist = np.full((nq, max(k, 0)), np.inf, np.float32)
        cnt = np.zeros(nq, np.uint32)
        st = _ffi.load().pcr_knn(self._h, _p(qx, _ffi.f32p), _p(qy, _ffi.f32p), _p(qz, _ffi.f32p), nq, k,
                                 _p(idx, _ffi.u32p), _p(dist, _ffi.f32p), _p(cnt, _ffi.u32p))
        _ffi.check(st, self._ctx._h)
        return idx, dist, cnt

    def knn_indices(self, queries: np.ndarray, k: int):
        q = np.ascontiguousarray(np.asarray(queries, np.float32).reshape(-1, 3))
        qx, qy, qz = _soa(q)
        nq = len(qx)
        idx = np.full((nq, max(k, 0)), 0xFFFFFFFF, np.uint32)
        cnt = np.zeros(nq, np.uint32)
        st = _ffi.load().pcr_knn(self._h, _p(qx, _ffi.f32p), _p(qy, _ffi.f32p), _p(qz, _ffi.f32p), nq, k,
                                 _p(idx, _ffi.u32p), None, _p(cnt, _ffi.u32p))
        _ffi.check(st, self._ctx._h)
        return idx, cnt

    def radius_count(self, queries: np.ndarray, radius: float) -> np.ndarray:
        q = np.ascontiguousarray(np.asarray(queries, np.float32).reshape(-1, 3))
        qx, qy, qz = _soa(q)
        cnt = np.zeros(len(qx), np.uint32)
        st = _ffi.load().pcr_radius_count(self._h, _p(qx, _ffi.f32p), _p(qy, _ffi.f32p), _p(qz, _ffi.f32p), len(qx),
                                          float(radius), _p(cnt, _ffi.u32p))
        _ffi.check(st, self._ctx._h)
        return cnt

    def radius_search(self, queries: np.ndarray, radius: float):
        """CSR: (offsets (nq+1,) u64, idx u32) with every row ascending by index (kdtree.rs:132)."""
        q = np.ascontiguousarray(np.asarray(queries, np.float32).reshape(-1, 3))
        qx, qy, qz = _soa(q)
        nq = len(qx)
        off = np.zeros(nq + 1, np.uint64)
        total = C.c_size_t()
        cap = max(16 * nq, 1024)
        lib = _ffi.load()
        for _ in range(2):
            idx = np.empty(cap, np.uint32)
            st = lib.pcr_radius_search(self._h, _p(qx, _ffi.f32p), _p(qy, _ffi.f32p), _p(qz, _ffi.f32p), nq, float(radius),
                                       _p(off, _ffi.u64p), _p(idx, _ffi.u32p), cap, C.byref(total))
            if st == _ffi.PCR_ERR_CAPACITY:
                cap = int(total.value)
                continue
            _ffi.check(st, self._ctx._h)
            return off, idx[: total.value].copy()
        raise PcrError(_ffi.PCR_ERR_CAPACITY, "radius_search: capacity retry failed")

    def find_correspondences(self, source: PointCloud, max_distance: float = math.inf):
        ns = len(source)
        si = np.empty(max(ns, 1), np.uint32)
        ti = np.empty(max(ns, 1), np.uint32)
        dd = np.empty(max(ns, 1), np.float32)
        m = C.c_size_t()
        st = _ffi.load().pcr_find_correspondences(self._h, _p(source.x, _ffi.f32p), _p(source.y, _ffi.f32p), _p(source.z, _ffi.f32p),
                                                  ns, float(max_distance), _p(si, _ffi.u32p), _p(ti, _ffi.u32p), _p(dd, _ffi.f32p),
                                                  C.byref(m))
        _ffi.check(st, self._ctx._h)
        return si[: m.value].copy(), ti[: m.value].copy(), dd[: m.value].copy()


def find_correspondences(source: PointCloud, target_tree: KdTree, max_distance: float = math.inf):
    return target_tree.find_correspondences(source, max_distance)


# ---------------------------------------------------------------------------------------------------
# filters
# ---------------------------------------------------------------------------------------------------

def sor_mask(cloud: PointCloud, k: int, std_mul: float, ctx: Optional[Context] = None, want_mean: bool = False):
    """keep mask (u8), kept count, [mean distances], (mean, std, threshold)."""
    ctx = ctx or default_context()
    if not math.isfinite(std_mul) or std_mul < 0.0:  # crates/python/src/filters.rs:42-46
        raise ValueError("std_mul must be >= 0 and finite")
    n = len(cloud)
    keep = np.zeros(max(n, 1), np.uint8)
    kept = C.c_size_t()
    mean_d = np.empty(max(n, 1), np.float32) if want_mean else None
    stats = np.zeros(3, np.float32)
    st = _ffi.load().pcr_sor(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n, k, float(std_mul),
                             _p(keep, _ffi.u8p), C.byref(kept), _p(mean_d, _ffi.f32p), _p(stats, _ffi.f32p))
    _ffi.check(st, ctx._h)
    return keep[:n], int(kept.value), (mean_d[:n] if want_mean else None), stats


def statistical_outlier_removal(cloud: PointCloud, k: int, std_mul: float, ctx: Optional[Context] = None) -> PointCloud:
    keep, _, _, _ = sor_mask(cloud, k, std_mul, ctx)
    return cloud.select(np.nonzero(keep)[0])  # statistical_outlier.rs:68


def voxel_downsample(cloud: PointCloud, voxel_size: float, ctx: Optional[Context] = None) -> PointCloud:
    """crates/python/src/filters.rs:4-13 (ValueError unless voxel_size > 0 and finite)."""
    ctx = ctx or default_context()
    if not math.isfinite(voxel_size) or voxel_size <= 0.0:
        raise ValueError("voxel_size must be > 0 and finite")
    n = len(cloud)
    ox, oy, oz = (np.zeros(max(n, 1), np.float32) for _ in range(3))
    m = C.c_size_t()
    st = _ffi.load().pcr_voxel_downsample(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                          float(voxel_size), _p(ox, _ffi.f32p), _p(oy, _ffi.f32p), _p(oz, _ffi.f32p), C.byref(m))
    _ffi.check(st, ctx._h)
    k = int(m.value)
    return PointCloud._from_xyz(ox[:k].copy(), oy[:k].copy(), oz[:k].copy())


def ror_mask(cloud: PointCloud, radius: float, min_neighbors: int, ctx: Optional[Context] = None):
    ctx = ctx or default_context()
    if not math.isfinite(radius) or radius <= 0.0:  # crates/python/src/filters.rs:57-61
        raise ValueError("radius must be > 0 and finite")
    n = len(cloud)
    keep = np.zeros(max(n, 1), np.uint8)
    kept = C.c_size_t()
    st = _ffi.load().pcr_radius_outlier(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                        float(radius), int(min_neighbors), _p(keep, _ffi.u8p), C.byref(kept))
    _ffi.check(st, ctx._h)
    return keep[:n], int(kept.value)


def radius_outlier_removal(cloud: PointCloud, radius: float, min_neighbors: int, ctx: Optional[Context] = None) -> PointCloud:
    keep, _ = ror_mask(cloud, radius, min_neighbors, ctx)
    return cloud.select(np.nonzero(keep)[0])


# ---------------------------------------------------------------------------------------------------
# segmentation
# ---------------------------------------------------------------------------------------------------

def euclidean_cluster(cloud: PointCloud, distance_threshold: float, min_size: int, max_size: int,
                      ctx: Optional[Context] = None) -> List[List[int]]:
    """crates/python/src/segmentation.rs:4-17 -> list of index lists (largest cluster first)."""
    ctx = ctx or default_context()
    n = len(cloud)
    off = np.zeros(n + 1, np.uint32)
    idx = np.zeros(max(n, 1), np.uint32)
    nc = C.c_size_t()
    st = _ffi.load().pcr_euclidean_cluster(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                           float(distance_threshold), int(min_size), int(max_size), _p(off, _ffi.u32p),
                                           _p(idx, _ffi.u32p), C.byref(nc))
    _ffi.check(st, ctx._h)
    return [idx[off[c]:off[c + 1]].tolist() for c in range(int(nc.value))]


class PlaneResult:
    """crates/python/src/segmentation.rs:19-36."""

    def __init__(self, normal, d, inliers):
        self.normal = [float(v) for v in normal]
        self.d = float(d)
        self.inliers = inliers

    def __repr__(self) -> str:
        return f"PlaneResult(normal={self.normal}, d={self.d:.4f}, inliers={len(self.inliers)})"


def draw_plane_samples(n: int, iterations: int, seed=None) -> np.ndarray:
    """The reference's sample_three_distinct (ransac_plane.rs:140-163) with numpy's generator: same procedure,
    NOT the same stream as Rust's StdRng (ChaCha12).  A Rust host passes its own triples to the C ABI."""
    rng = np.random.default_rng(seed)
    out = []
    if n < 3:
        return np.zeros((0, 3), np.uint32)
    for _ in range(int(iterations)):
        i0 = int(rng.integers(n))
        i1, tries = int(rng.integers(n)), 0
        while i1 == i0 and tries <= 100:
            i1, tries = int(rng.integers(n)), tries + 1
        if i1 == i0:
            continue
        i2, tries = int(rng.integers(n)), 0
        while i2 in (i0, i1) and tries <= 100:
            i2, tries = int(rng.integers(n)), tries + 1
        if i2 in (i0, i1):
            continue
        out.append((i0, i1, i2))
    return np.asarray(out, np.uint32).reshape(-1, 3)


def ransac_plane_samples(cloud: PointCloud, distance_threshold: float, samples, ctx: Optional[Context] = None) -> PlaneResult:
    """ransac_plane_seeded (ransac_plane.rs:56-129) for given index triples."""
    ctx = ctx or default_context()
    n = len(cloud)
    smp = np.ascontiguousarray(np.asarray(samples, np.uint32).reshape(-1, 3))
    if len(smp) and int(smp.max()) >= n:
        raise IndexError(f"sample index {int(smp.max())} out of bounds for cloud with {n} points")
    model = np.zeros(4, np.float32)
    inl = np.zeros(max(n, 1), np.uint32)
    k = C.c_size_t()
    st = _ffi.load().pcr_ransac_plane_samples(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                              float(distance_threshold), _p(smp, _ffi.u32p), len(smp), _p(model, _ffi.f32p),
                                              _p(inl, _ffi.u32p), C.byref(k))
    _ffi.check(st, ctx._h)
    return PlaneResult(model[:3], model[3], inl[:int(k.value)].tolist())


def ransac_plane(cloud: PointCloud, distance_threshold: float, iterations: int, ctx: Optional[Context] = None) -> PlaneResult:
    """crates/python/src/segmentation.rs:42-55: a random seed per call, like the reference's ransac_plane (:36-43)."""
    return ransac_plane_samples(cloud, distance_threshold, draw_plane_samples(len(cloud), iterations), ctx)


def cluster_arrays(cloud: PointCloud, distance_threshold: float, min_size: int, max_size: int, ctx: Optional[Context] = None):
    """Same call, CSR form (offsets, indices) without building Python lists."""
    ctx = ctx or default_context()
    n = len(cloud)
    off = np.zeros(n + 1, np.uint32)
    idx = np.zeros(max(n, 1), np.uint32)
    nc = C.c_size_t()
    st = _ffi.load().pcr_euclidean_cluster(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                           float(distance_threshold), int(min_size), int(max_size), _p(off, _ffi.u32p),
                                           _p(idx, _ffi.u32p), C.byref(nc))
    _ffi.check(st, ctx._h)
    k = int(nc.value)
    return off[:k + 1].copy(), idx[:int(off[k])].copy()


# ---------------------------------------------------------------------------------------------------
# normals
# ---------------------------------------------------------------------------------------------------

def normals_array(cloud: PointCloud, k: int, viewpoint=(0.0, 0.0, 0.0), ctx: Optional[Context] = None) -> np.ndarray:
    """estimate_normals_with_viewpoint -> (N,3) f32 (empty if the cloud is empty or k == 0)."""
    ctx = ctx or default_context()
    n = len(cloud)
    if n == 0 or k == 0:
        return np.zeros((0, 3), np.float32)
    vp = np.asarray(viewpoint, np.float32).reshape(3)
    nx = np.empty(n, np.float32)
    ny = np.empty(n, np.float32)
    nz = np.empty(n, np.float32)
    st = _ffi.load().pcr_estimate_normals(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n, k,
                                          _p(vp, _ffi.f32p), _p(nx, _ffi.f32p), _p(ny, _ffi.f32p), _p(nz, _ffi.f32p))
    _ffi.check(st, ctx._h)
    return np.stack([nx, ny, nz], axis=1)


def estimate_normals(cloud: PointCloud, k: int, ctx: Optional[Context] = None) -> PointCloud:
    """Returns a copy of the cloud with normals attached (crates/python/src/normals.rs:4-10)."""
    nr = normals_array(cloud, k, (0.0, 0.0, 0.0), ctx)
    return PointCloud._from_xyz(cloud.x.copy(), cloud.y.copy(), cloud.z.copy(), nr)


# ---------------------------------------------------------------------------------------------------
# registration
# ---------------------------------------------------------------------------------------------------

class IcpResult:
    def __init__(self, r: _ffi.IcpResultC):
        self.converged = bool(r.converged)
        self.fitness = float(np.float32(r.fitness))
        self.rmse = float(np.float32(r.rmse))
        self.num_iterations = int(r.num_iterations)
        self.translation = [float(v) for v in r.translation]
        rot = [float(v) for v in r.rotation]
        self.rotation = [rot[0:3], rot[3:6], rot[6:9]]

    def __repr__(self) -> str:
        return f"IcpResult(converged={str(self.converged).lower()}, rmse={self.rmse:.6f}, iterations={self.num_iterations})"


def _icp_params(max_iterations, tolerance, max_correspondence_distance):
    # crates/python/src/registration.rs:67-72,85
    if math.isnan(tolerance) or tolerance < 0:
        raise ValueError("tolerance must be >= 0")
    if math.isnan(max_correspondence_distance) or max_correspondence_distance < 0:
        raise ValueError("max_correspondence_distance must be >= 0")
    return _ffi.IcpParams(int(max_iterations), float(tolerance), float(max_correspondence_distance))


def icp_point_to_point(source: PointCloud, target: PointCloud, max_iterations: int = 50, tolerance: float = 1e-5,
                       max_correspondence_distance: float = math.inf, ctx: Optional[Context] = None) -> IcpResult:
    ctx = ctx or default_context()
    p = _icp_params(max_iterations, tolerance, max_correspondence_distance)
    res = _ffi.IcpResultC()
    st = _ffi.load().pcr_icp_point_to_point(ctx._h, _p(source.x, _ffi.f32p), _p(source.y, _ffi.f32p), _p(source.z, _ffi.f32p),
                                            len(source), _p(target.x, _ffi.f32p), _p(target.y, _ffi.f32p), _p(target.z, _ffi.f32p),
                                            len(target), C.byref(p), C.byref(res))
    _ffi.check(st, ctx._h)
    return IcpResult(res)


def icp_point_to_plane(source: PointCloud, target: PointCloud, max_iterations: int = 50, tolerance: float = 1e-5,
                       max_correspondence_distance: float = math.inf, ctx: Optional[Context] = None) -> IcpResult:
    ctx = ctx or default_context()
    if target.normals is None:  # crates/python/src/registration.rs:66-71
        raise ValueError("target cloud must have normals for point-to-plane ICP. Use estimate_normals(target, k) first.")
    p = _icp_params(max_iterations, tolerance, max_correspondence_distance)
    nx, ny, nz = target._normals_soa()
    res = _ffi.IcpResultC()
    st = _ffi.load().pcr_icp_point_to_plane(ctx._h, _p(source.x, _ffi.f32p), _p(source.y, _ffi.f32p), _p(source.z, _ffi.f32p),
                                            len(source), _p(target.x, _ffi.f32p), _p(target.y, _ffi.f32p), _p(target.z, _ffi.f32p),
                                            len(target), _p(nx, _ffi.f32p), _p(ny, _ffi.f32p), _p(nz, _ffi.f32p), len(nx),
                                            C.byref(p), C.byref(res))
    _ffi.check(st, ctx._h)
    return IcpResult(res)


def apply_transform(cloud: PointCloud, rotation, translation, ctx: Optional[Context] = None) -> PointCloud:
    """R*p + t for every point; the result carries xyz only (icp.rs:77-92)."""
    ctx = ctx or default_context()
    n = len(cloud)
    r = np.ascontiguousarray(np.asarray(rotation, np.float32).reshape(9))
    t = np.ascontiguousarray(np.asarray(translation, np.float32).reshape(3))
    ox = np.empty(n, np.float32)
    oy = np.empty(n, np.float32)
    oz = np.empty(n, np.float32)
    st = _ffi.load().pcr_apply_transform(ctx._h, _p(cloud.x, _ffi.f32p), _p(cloud.y, _ffi.f32p), _p(cloud.z, _ffi.f32p), n,
                                         _p(r, _ffi.f32p), _p(t, _ffi.f32p), _p(ox, _ffi.f32p), _p(oy, _ffi.f32p), _p(oz, _ffi.f32p))
    _ffi.check(st, ctx._h)
    return PointCloud._from_xyz(ox, oy, oz)



# ---------------------------------------------------------------------------------------------------
# device-resident clouds (superset of the reference surface: the same calls without the PCIe round
# trips between the steps of a pipeline)
# ---------------------------------------------------------------------------------------------------

class DeviceCloud:
    """A PointCloud that lives in HBM.  Every filter returns a new DeviceCloud; nothing crosses PCIe
    until to_numpy() / normals_to_numpy() / the cluster index lists / the ICP result."""

    def __init__(self, handle, ctx: Context):
        self._h = handle
        self._ctx = ctx
        ctx._clouds.add(self)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @staticmethod
    def from_numpy(array, ctx: Optional[Context] = None) -> "DeviceCloud":
        """(N, 3) float32 / float64, C-contiguous (crates/python/src/cloud.rs:25-37).  A float32 array crosses PCIe as it is
        and is split into SoA on the device (pcr_cloud_upload_rows): splitting 122 K points with numpy costs more than the
        whole voxel -> SOR -> normals pipeline."""
        if isinstance(array, np.ndarray) and array.dtype == np.float32 and array.ndim == 2 and array.shape[1] == 3 and array.flags["C_CONTIGUOUS"]:
            ctx = ctx or default_context()
            h = C.c_void_p()
            _ffi.check(_ffi.load().pcr_cloud_upload_rows(ctx._h, array.ctypes.data, len(array), C.byref(h)), ctx._h)
            return DeviceCloud(h, ctx)
        return DeviceCloud.from_cloud(PointCloud.from_numpy(array), ctx)  # (same checks and errors as the host class)

    @staticmethod
    def from_cloud(cloud: PointCloud, ctx: Optional[Context] = None) -> "DeviceCloud":
        ctx = ctx or default_context()
        h = C.c_void_p()
        st = _ffi.load().pcr_cloud_upload(ctx._h, cloud.x.ctypes.data, cloud.y.ctypes.data, cloud.z.ctypes.data, len(cloud), C.byref(h))
        _ffi.check(st, ctx._h)
        return DeviceCloud(h, ctx)

    @staticmethod
    def upload_raw(ctx: Context, x_ptr: int, y_ptr: int, z_ptr: int, n: int) -> "DeviceCloud":
        """Upload from raw host addresses (e.g. pinned buffers)."""
        h = C.c_void_p()
        _ffi.check(_ffi.load().pcr_cloud_upload(ctx._h, x_ptr, y_ptr, z_ptr, n, C.byref(h)), ctx._h)
        return DeviceCloud(h, ctx)

    @staticmethod
    def upload_block(ctx: Context, ptr: int, stride: int, n: int, wait: bool = True) -> "DeviceCloud":
        """Upload x | y | z from ONE host block (rows `stride` floats apart): a single strided transfer.
        wait=False does not wait for the copy (pinned block, left unchanged until a result that depends on the
        cloud has come back): the next steps queue right behind it."""
        h = C.c_void_p()
        lib = _ffi.load()
        fn = lib.pcr_cloud_upload_block if wait else lib.pcr_cloud_upload_block_nowait
        _ffi.check(fn(ctx._h, ptr, stride, n, C.byref(h)), ctx._h)
        return DeviceCloud(h, ctx)

    def download_block(self, ptr: int, stride: int, with_normals: bool = False):
        """Download x | y | z [| nx | ny | nz] into ONE host block (rows `stride` floats apart)."""
        _ffi.check(_ffi.load().pcr_cloud_download_block(self._h, ptr, stride, 1 if with_normals else 0), self._ctx._h)

    def download_raw(self, x_ptr: int, y_ptr: int, z_ptr: int, nx_ptr: int = 0, ny_ptr: int = 0, nz_ptr: int = 0):
        """Download into raw host addresses (arrays of len()); normals too if the three pointers are given."""
        _ffi.check(_ffi.load().pcr_cloud_download(self._h, x_ptr, y_ptr, z_ptr), self._ctx._h)
        if nx_ptr:
            _ffi.check(_ffi.load().pcr_cloud_download_normals(self._h, nx_ptr, ny_ptr, nz_ptr), self._ctx._h)

    def free(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._ctx._h:  # a closed context has already freed its clouds
            _ffi.load().pcr_cloud_free(h)

    def _new(self, fn, *args) -> "DeviceCloud":
        h = C.c_void_p()
        _ffi.check(fn(self._h, *args, C.byref(h)), self._ctx._h)
        return DeviceCloud(h, self._ctx)

    def len(self) -> int:
        return int(_ffi.load().pcr_cloud_len(self._h))

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def has_normals(self) -> bool:
        return bool(_ffi.load().pcr_cloud_has_normals(self._h))

    def __repr__(self) -> str:
        return f"DeviceCloud(n={self.len()}, normals={self.has_normals()})"

    def to_numpy(self) -> np.ndarray:
        out = np.empty((self.len(), 3), np.float32)  # (interleaved on the device: one contiguous transfer)
        _ffi.check(_ffi.load().pcr_cloud_download_rows(self._h, out.ctypes.data, None), self._ctx._h)
        return out

    def normals_to_numpy(self) -> Optional[np.ndarray]:
        if not self.has_normals():
            return None
        out = np.empty((self.len(), 3), np.float32)
        _ffi.check(_ffi.load().pcr_cloud_download_rows(self._h, None, out.ctypes.data), self._ctx._h)
        return out

    def to_cloud(self) -> PointCloud:
        a = self.to_numpy()
        return PointCloud._from_xyz(*_soa(a), self.normals_to_numpy())

    def select(self, indices: Sequence[int]) -> "DeviceCloud":
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        n = self.len()
        bad = idx[(idx >= n) | (idx < 0)]
        if len(bad):
            raise IndexError(f"index {int(bad[0])} out of bounds for cloud with {n} points")
        i32 = np.ascontiguousarray(idx, np.uint32)
        return self._new(_ffi.load().pcr_cloud_select, _p(i32, _ffi.u32p), len(i32))

    def voxel_downsample(self, voxel_size: float) -> "DeviceCloud":
        if not math.isfinite(voxel_size) or voxel_size <= 0.0:
            raise ValueError("voxel_size must be > 0 and finite")
        return self._new(_ffi.load().pcr_cloud_voxel_downsample, float(voxel_size))

    def statistical_outlier_removal(self, k: int, std_mul: float) -> "DeviceCloud":
        if not math.isfinite(std_mul) or std_mul < 0.0:  # crates/python/src/filters.rs:42-46
            raise ValueError("std_mul must be >= 0 and finite")
        return self._new(_ffi.load().pcr_cloud_statistical_outlier_removal, int(k), float(std_mul))

    def radius_outlier_removal(self, radius: float, min_neighbors: int) -> "DeviceCloud":
        if not math.isfinite(radius) or radius <= 0.0:
            raise ValueError("radius must be > 0 and finite")
        return self._new(_ffi.load().pcr_cloud_radius_outlier_removal, float(radius), int(min_neighbors))

    def estimate_normals(self, k: int, viewpoint=(0.0, 0.0, 0.0)) -> "DeviceCloud":
        vp = np.asarray(viewpoint, np.float32)
        return self._new(_ffi.load().pcr_cloud_estimate_normals, int(k), _p(vp, _ffi.f32p))

    def sor_normals(self, k_sor: int, std_mul: float, k_normals: int, viewpoint=(0.0, 0.0, 0.0)) -> "DeviceCloud":
        """statistical_outlier_removal(k_sor, std_mul) then estimate_normals(k_normals) on the kept points, one index."""
        vp = np.asarray(viewpoint, np.float32)
        return self._new(_ffi.load().pcr_cloud_sor_normals, int(k_sor), float(std_mul), int(k_normals), _p(vp, _ffi.f32p))

    def euclidean_cluster(self, distance_threshold: float, min_size: int, max_size: int) -> List[List[int]]:
        n = self.len()
        off = np.zeros(n + 1, np.uint32)
        idx = np.zeros(max(n, 1), np.uint32)
        nc = C.c_size_t()
        st = _ffi.load().pcr_cloud_euclidean_cluster(self._h, float(distance_threshold), int(min_size), int(max_size), _p(off, _ffi.u32p),
                                                     _p(idx, _ffi.u32p), C.byref(nc))
        _ffi.check(st, self._ctx._h)
        return [idx[off[c]:off[c + 1]].tolist() for c in range(int(nc.value))]

    def ransac_plane_samples(self, distance_threshold: float, samples) -> "PlaneResult":
        n = self.len()
        smp = np.ascontiguousarray(np.asarray(samples, np.uint32).reshape(-1, 3))
        if len(smp) and int(smp.max()) >= n:
            raise IndexError(f"sample index {int(smp.max())} out of bounds for cloud with {n} points")
        model = np.zeros(4, np.float32)
        inl = np.zeros(max(n, 1), np.uint32)
        k = C.c_size_t()
        st = _ffi.load().pcr_cloud_ransac_plane_samples(self._h, float(distance_threshold), _p(smp, _ffi.u32p), len(smp),
                                                        _p(model, _ffi.f32p), _p(inl, _ffi.u32p), C.byref(k))
        _ffi.check(st, self._ctx._h)
        return PlaneResult(model[:3], model[3], inl[:int(k.value)].tolist())

    def ransac_plane(self, distance_threshold: float, iterations: int) -> "PlaneResult":
        return self.ransac_plane_samples(distance_threshold, draw_plane_samples(self.len(), iterations))

    def apply_transform(self, rotation, translation) -> "DeviceCloud":
        r = np.ascontiguousarray(np.asarray(rotation, np.float32).reshape(9))
        t = np.ascontiguousarray(np.asarray(translation, np.float32).reshape(3))
        return self._new(_ffi.load().pcr_cloud_apply_transform, _p(r, _ffi.f32p), _p(t, _ffi.f32p))

    def icp_point_to_point(self, target: "DeviceCloud", max_iterations: int = 50, tolerance: float = 1e-5,
                           max_correspondence_distance: float = math.inf) -> "IcpResult":
        prm = _icp_params(max_iterations, tolerance, max_correspondence_distance)
        res = _ffi.IcpResultC()
        _ffi.check(_ffi.load().pcr_cloud_icp_point_to_point(self._h, target._h, C.byref(prm), C.byref(res)), self._ctx._h)
        return IcpResult(res)

    def icp_point_to_plane(self, target: "DeviceCloud", max_iterations: int = 50, tolerance: float = 1e-5,
                           max_correspondence_distance: float = math.inf) -> "IcpResult":
        prm = _icp_params(max_iterations, tolerance, max_correspondence_distance)
        res = _ffi.IcpResultC()
        _ffi.check(_ffi.load().pcr_cloud_icp_point_to_plane(self._h, target._h, C.byref(prm), C.byref(res)), self._ctx._h)
        return IcpResult(res)

# ---------------------------------------------------------------------------------------------------
# multi-frame batch (BASELINE config 5)
# ---------------------------------------------------------------------------------------------------

def sor_normals_batch(points: np.ndarray, frame_offsets: Sequence[int], k_sor: int, std_mul: float, k_normals: int,
                      viewpoint=(0.0, 0.0, 0.0), ctx: Optional[Context] = None):
    """points: (N,3) f32 of all frames back to back. -> (keep u8 (N,), normals (N,3), kept per frame)."""
    ctx = ctx or default_context()
    pts = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1, 3))
    off = np.ascontiguousarray(np.asarray(frame_offsets, np.uint64))
    nf = len(off) - 1
    n = len(pts)
    vp = np.asarray(viewpoint, np.float32).reshape(3)
    keep = np.empty(max(n, 1), np.uint8)            # (every entry is written: removed points get keep 0 and a zero normal)
    nrm = np.empty((max(n, 1), 3), np.float32) if k_normals else np.zeros((max(n, 1), 3), np.float32)
    kept = np.zeros(max(nf, 1), np.uint64)
    # the rows go to the device as they are (pcr_sor_normals_batch_rows): numpy's split of 8 M points into SoA and back
    # cost 90 of the 103 ms this call took, the work itself 11
    st = _ffi.load().pcr_sor_normals_batch_rows(ctx._h, _p(pts, _ffi.f32p), _p(off, _ffi.u64p), nf, k_sor, float(std_mul), k_normals,
                                                _p(vp, _ffi.f32p), _p(keep, _ffi.u8p), _p(nrm, _ffi.f32p), _p(kept, _ffi.u64p))
    _ffi.check(st, ctx._h)
    return keep[:n], nrm[:n], kept[:nf]


def sor_normals_batch_raw(ctx: Context, x, y, z, n: int, frame_offsets: np.ndarray, k_sor: int, std_mul: float,
                          k_normals: int, viewpoint: np.ndarray, keep, nx, ny, nz, kept=None, device: bool = False):
    """Thin call for benchmarks: x..nz are raw addresses (ints).  device=True -> *_dev entry point
    (device pointers, nothing copied, asynchronous on the context's stream)."""
    lib = _ffi.load()
    nf = len(frame_offsets) - 1
    if device:
        st = lib.pcr_sor_normals_batch_dev(ctx._h, x, y, z, _p(frame_offsets, _ffi.u64p), nf, k_sor, float(std_mul), k_normals,
                                           _p(viewpoint, _ffi.f32p), keep, nx, ny, nz)
    else:
        cast = lambda a, t: C.cast(C.c_void_p(a), t)
        st = lib.pcr_sor_normals_batch(ctx._h, cast(x, _ffi.f32p), cast(y, _ffi.f32p), cast(z, _ffi.f32p),
                                       _p(frame_offsets, _ffi.u64p), nf, k_sor, float(std_mul), k_normals,
                                       _p(viewpoint, _ffi.f32p), cast(keep, _ffi.u8p), cast(nx, _ffi.f32p), cast(ny, _ffi.f32p),
                                       cast(nz, _ffi.f32p), _p(kept, _ffi.u64p) if kept is not None else None)
    _ffi.check(st, ctx._h)
