"""ransac_plane: the oracle pinned to the reference's tests (crates/segmentation/src/ransac_plane.rs:192-423) and the
CUDA path bit-exact against it FOR THE SAME SAMPLES.  The reference draws its samples from StdRng (ChaCha12), which
cannot be reproduced without Rust: the tests use numpy's generator with the reference's sampling procedure, and the
C ABI takes the triples from the caller."""
import numpy as np
import pytest

import pointclouds_rs_b200 as pcr_pkg
from pointclouds_rs_b200 import scenes


def grid_plane(nx, ny, step, zfn):
    i, j = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), indexing="ij")
    x, y = (i * np.float32(step)).ravel(), (j * np.float32(step)).ravel()
    return np.stack([x, y, zfn(x, y).astype(np.float32)], 1).astype(np.float32)


def samples(n, iters, seed=42):
    return pcr_pkg.draw_plane_samples(n, iters, seed)


# ---- oracle vs the reference's own tests --------------------------------------------------------
def test_oracle_fit_xy_plane(oracle):  # :198-221
    pts = grid_plane(20, 20, 0.1, lambda x, y: 0 * x)
    m, inl = oracle.ransac_plane_samples(pts, 0.01, samples(400, 100))
    assert abs(m[2]) > 0.99 and abs(m[3]) < 0.01 and len(inl) == 400


def test_oracle_fit_offset_plane(oracle):  # :224-253
    pts = grid_plane(10, 10, 1.0, lambda x, y: 0 * x + 5)
    m, inl = oracle.ransac_plane_samples(pts, 0.01, samples(100, 100))
    assert abs(m[2]) > 0.99 and abs(abs(m[3]) - 5.0) < 0.01 and len(inl) == 100


def test_oracle_fit_tilted_plane(oracle):  # :256-304
    pts = grid_plane(10, 10, 0.1, lambda x, y: 1.0 - x - y)
    m, inl = oracle.ransac_plane_samples(pts, 0.01, samples(100, 100))
    assert np.all(np.abs(np.abs(m[:3]) - 1 / np.sqrt(3)) < 0.05) and len(inl) == 100


def test_oracle_plane_with_outliers(oracle):  # :307-354
    pts = np.vstack([grid_plane(7, 7, 1.0, lambda x, y: 0 * x),
                     np.stack([np.arange(10), np.arange(10), np.full(10, 100.0)], 1)]).astype(np.float32)
    m, inl = oracle.ransac_plane_samples(pts, 0.1, samples(59, 200))
    assert abs(m[2]) > 0.9 and len(inl) >= 49 and np.all(np.abs(pts[inl, 2]) < 1.0)


def test_oracle_degenerate(oracle):  # :357-378
    for n in (0, 1, 2):
        pts = np.zeros((n, 3), np.float32)
        m, inl = oracle.ransac_plane_samples(pts, 0.1, np.zeros((0, 3), np.uint32))
        assert m.tolist() == [0.0, 0.0, 1.0, 0.0] and len(inl) == 0
    line = np.stack([np.arange(50), np.zeros(50), np.zeros(50)], 1).astype(np.float32)  # every sample collinear
    m, inl = oracle.ransac_plane_samples(line, 0.1, samples(50, 30))
    assert m.tolist() == [0.0, 0.0, 1.0, 0.0] and len(inl) == 50  # default plane z = 0 holds the whole line


def test_oracle_both_paths_pick_the_first_maximum(oracle):
    """The sequential (:95-121) and the parallel (:82-93) path agree whenever the early exit does not fire."""
    rng = np.random.default_rng(1)
    plane = np.concatenate([rng.uniform(-5, 5, (6000, 2)), rng.normal(0, 0.02, (6000, 1))], 1)
    noise = rng.uniform(-5, 5, (6000, 3))
    small = np.vstack([plane[:4000], noise[:5000]]).astype(np.float32)       # n < 10000: sequential
    big = np.vstack([plane, noise]).astype(np.float32)                       # n >= 10000: parallel rule
    for pts in (small, big):
        s = samples(len(pts), 64, 3)
        m, inl = oracle.ransac_plane_samples(pts, 0.05, s)
        d = np.abs(pts @ m[:3] + m[3])
        assert len(inl) == int((d <= 0.05 + 1e-6).sum()) or abs(len(inl) - int((d <= 0.05).sum())) <= 2
        assert len(inl) > 0.3 * len(pts) and abs(m[2]) > 0.95


# ---- CUDA path ------------------------------------------------------------------------------------
def _gpu_equal(pcr, oracle, pts, thr, s):
    pts = np.ascontiguousarray(pts, np.float32)
    got = pcr.ransac_plane_samples(pcr.PointCloud.from_numpy(pts), thr, s)
    m, inl = oracle.ransac_plane_samples(pts, thr, s)
    assert np.array_equal(np.array(got.normal + [got.d], np.float32).view(np.uint32), m.view(np.uint32))
    assert got.inliers == inl.tolist()
    dev = pcr.DeviceCloud.from_numpy(pts).ransac_plane_samples(thr, s)
    assert dev.normal == got.normal and dev.d == got.d and dev.inliers == got.inliers


@pytest.mark.gpu
def test_gpu_reference_kats(pcr, oracle):
    _gpu_equal(pcr, oracle, grid_plane(20, 20, 0.1, lambda x, y: 0 * x), 0.01, samples(400, 100))
    _gpu_equal(pcr, oracle, grid_plane(10, 10, 1.0, lambda x, y: 0 * x + 5), 0.01, samples(100, 100))
    _gpu_equal(pcr, oracle, grid_plane(10, 10, 0.1, lambda x, y: 1.0 - x - y), 0.01, samples(100, 100))
    for n in (0, 1, 2):
        r = pcr.ransac_plane_samples(pcr.PointCloud.from_numpy(np.zeros((n, 3), np.float32)), 0.1, np.zeros((0, 3), np.uint32))
        assert r.normal == [0.0, 0.0, 1.0] and r.d == 0.0 and r.inliers == []
    line = np.stack([np.arange(50), np.zeros(50), np.zeros(50)], 1).astype(np.float32)
    _gpu_equal(pcr, oracle, line, 0.1, samples(50, 30))
    with pytest.raises(IndexError):
        pcr.ransac_plane_samples(pcr.PointCloud.from_numpy(line), 0.1, [[0, 1, 50]])


@pytest.mark.gpu
def test_gpu_both_paths_bit_exact(pcr, oracle):
    rng = np.random.default_rng(1)
    plane = np.concatenate([rng.uniform(-5, 5, (6000, 2)), rng.normal(0, 0.02, (6000, 1))], 1)
    noise = rng.uniform(-5, 5, (6000, 3))
    _gpu_equal(pcr, oracle, np.vstack([plane[:4000], noise[:5000]]), 0.05, samples(9000, 64, 3))     # sequential rule + early exit
    _gpu_equal(pcr, oracle, np.vstack([plane, noise]), 0.05, samples(12000, 700, 4))                 # parallel rule, two model chunks
    bad = np.vstack([plane[:3000], [[np.nan, 0, 0], [np.inf, 1, 1]]]).astype(np.float32)             # non-finite points never count
    _gpu_equal(pcr, oracle, bad, 0.05, samples(3000, 40, 5))


@pytest.mark.gpu
def test_gpu_kitti_ground_plane_full_size(pcr, oracle):
    """The ground-removal step of examples/python/kitti_obstacle_detection.py on the 122 K frame, 500 hypotheses."""
    pts = scenes.kitti_scene()
    s = samples(len(pts), 500, 7)
    _gpu_equal(pcr, oracle, pts, 0.15, s)
    r = pcr.ransac_plane(pcr.PointCloud.from_numpy(pts), 0.15, 200)     # random seed per call, like the reference
    assert abs(r.normal[2]) > 0.99 and len(r.inliers) > 100_000
