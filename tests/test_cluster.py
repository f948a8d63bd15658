"""euclidean_cluster: the oracle pinned to the reference's own tests, and the CUDA path against it.

CPU part: every known-answer test of crates/segmentation/src/euclidean_cluster.rs:189-436 and
tests/cluster_differential.rs (file:line given), run on the oracle's restatement of the grid
algorithm AND on its restatement of the reference's brute-force checker.
GPU part (-m gpu): the C-ABI call against the oracle, bit-exact cluster lists.
"""
import numpy as np
import pytest

from pointclouds_rs_b200 import scenes


def cloud(x, y=None, z=None):
    if y is None:
        return np.asarray(x, np.float32).reshape(-1, 3)
    return np.stack([np.asarray(x, np.float32), np.asarray(y, np.float32), np.asarray(z, np.float32)], 1)


def as_lists(cl):
    return [list(map(int, c)) for c in cl]


KATS = [
    # (name, points, threshold, min, max, expected lists or None, expected sizes)
    ("two_separated_clusters :195-214", cloud([0, .1, .2, 100, 100.1, 100.2], [0, .1, 0, 100, 100.1, 100], [0, 0, .1, 100, 100, 100.1]),
     1.0, 1, 100, [[0, 1, 2], [3, 4, 5]]),
    ("single_dense_cluster :217-228", cloud([0, .1, .2, .3], [0] * 4, [0] * 4), 0.5, 1, 100, [[0, 1, 2, 3]]),
    ("min_size_filter :238-249", cloud([0, .1, 50], [0] * 3, [0] * 3), 1.0, 2, 100, [[0, 1]]),
    ("max_size_filter :252-262", cloud([0, .1, .2, .3], [0] * 4, [0] * 4), 1.0, 1, 2, []),
    ("clusters_sorted_largest_first :265-277", cloud([0, .1, 50, 50.1, 50.2, 100], [0] * 6, [0] * 6), 1.0, 1, 100, [[2, 3, 4], [0, 1], [5]]),
    ("zero_threshold :280-284", cloud([0], [0], [0]), 0.0, 1, 100, []),
    ("negative_threshold :287-291", cloud([0], [0], [0]), -1.0, 1, 100, []),
    ("zero_min_size :294-298", cloud([0], [0], [0]), 1.0, 0, 100, []),
    ("transitive_connectivity :319-331", cloud([0, .4, .8], [0] * 3, [0] * 3), 0.5, 1, 100, [[0, 1, 2]]),
    ("cross_cell_boundary :335-347", cloud([.99, 1.01], [0, 0], [0, 0]), 1.0, 1, 100, [[0, 1]]),
    ("nan_and_inf_excluded :366-381", cloud([0, .1, np.nan, np.inf, .2], [0] * 5, [0] * 5), 0.5, 1, 100, [[0, 1, 4], [2], [3]]),
    ("points_exactly_at_threshold cluster_differential.rs:151-163", cloud([0, 1], [0, 0], [0, 0]), 1.0, 1, 100, [[0, 1]]),
    ("points_just_beyond_threshold :166-177", cloud([0, np.float32(1.0) + np.float32(1e-4)], [0, 0], [0, 0]), 1.0, 1, 100, [[0], [1]]),
    ("points_on_cell_boundaries :180-192", cloud([1 - 1e-5, 1 + 1e-5], [0, 0], [0, 0]), 1.0, 1, 100, [[0, 1]]),
    ("very_large_coordinates :195-208", cloud([1e6, 1e6 + .1, 1e6 + .2, 1e6 + 100], [1e6] * 4, [0] * 4), 0.5, 1, 100, None),
    ("very_small_threshold :211-224", cloud([0, 1, 2], [0] * 3, [0] * 3), 1e-6, 1, 100, [[0], [1], [2]]),
    ("duplicate_points_stable :311-325", cloud([0, 0, 0, 10, 10], [0] * 5, [0] * 5), 1.0, 1, 100, [[0, 1, 2], [3, 4]]),
]


@pytest.mark.parametrize("kat", KATS, ids=[k[0].split()[0] for k in KATS])
@pytest.mark.parametrize("brute", [False, True])
def test_oracle_reference_kats(oracle, kat, brute):
    _, pts, thr, mn, mx, want = kat
    got = as_lists(oracle.euclidean_cluster(pts, thr, mn, mx, brute=brute))
    if want is not None:
        assert got == want
    else:  # very_large_coordinates: the reference asserts the sizes only
        assert [len(c) for c in got] == [3, 1]


def test_oracle_empty_cloud(oracle):  # euclidean_cluster.rs:231-235
    assert oracle.euclidean_cluster(np.zeros((0, 3), np.float32), 1.0, 1, 100) == []


def test_oracle_dense_cell_stress(oracle):  # :351-362
    n = 5000
    pts = cloud(np.arange(n, dtype=np.float32) * np.float32(0.001), np.zeros(n), np.zeros(n))
    cl = oracle.euclidean_cluster(pts, 0.01, 1, n + 1)
    assert sum(len(c) for c in cl) == n


def _random_cases(seed, trials, lo, hi, tlo, thi, span):
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        n = int(rng.integers(lo, hi))
        thr = float(rng.uniform(tlo, thi))
        yield rng.uniform(-span, span, (n, 3)).astype(np.float32), thr


def test_oracle_differential_small(oracle):  # cluster_differential.rs:106-126 (PCG64 instead of ChaCha12)
    for pts, thr in _random_cases(42, 200, 2, 80, 0.5, 5.0, 20.0):
        a, b = oracle.euclidean_cluster(pts, thr, 1, len(pts)), oracle.euclidean_cluster(pts, thr, 1, len(pts), brute=True)
        assert as_lists(a) == as_lists(b)


def test_oracle_differential_medium(oracle):  # :129-147
    for pts, thr in _random_cases(99, 20, 500, 2000, 1.0, 8.0, 50.0):
        a, b = oracle.euclidean_cluster(pts, thr, 1, len(pts)), oracle.euclidean_cluster(pts, thr, 1, len(pts), brute=True)
        assert as_lists(a) == as_lists(b)


def test_oracle_translation_invariance(oracle):  # :283-308
    rng = np.random.default_rng(66)
    pts = rng.uniform(-10, 10, (150, 3)).astype(np.float32)
    a = oracle.euclidean_cluster(pts, 3.0, 1, 150)
    b = oracle.euclidean_cluster((pts + np.float32(12345.0)).astype(np.float32), 3.0, 1, 150)
    assert as_lists(a) == as_lists(b)


# ---- CUDA path ------------------------------------------------------------------------------------
def _gpu_equal(pcr, oracle, pts, thr, mn, mx):
    got = pcr.euclidean_cluster(pcr.PointCloud.from_numpy(np.ascontiguousarray(pts, np.float32)), thr, mn, mx)
    want = as_lists(oracle.euclidean_cluster(pts, thr, mn, mx))
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("kat", KATS, ids=[k[0].split()[0] for k in KATS])
def test_gpu_reference_kats(pcr, oracle, kat):
    _, pts, thr, mn, mx, want = kat
    got = pcr.euclidean_cluster(pcr.PointCloud.from_numpy(pts), thr, mn, mx)
    if want is not None:
        assert got == want
    assert got == as_lists(oracle.euclidean_cluster(pts, thr, mn, mx))


@pytest.mark.gpu
def test_gpu_differential_random(pcr, oracle):
    for pts, thr in _random_cases(7, 60, 2, 400, 0.5, 5.0, 20.0):
        _gpu_equal(pcr, oracle, pts, thr, 1, len(pts))
    for pts, thr in _random_cases(8, 6, 3000, 9000, 1.0, 6.0, 50.0):
        _gpu_equal(pcr, oracle, pts, thr, 2, len(pts))


@pytest.mark.gpu
def test_gpu_edge_cases(pcr, oracle):
    empty = pcr.PointCloud.from_numpy(np.zeros((0, 3), np.float32))
    assert pcr.euclidean_cluster(empty, 1.0, 1, 100) == []
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    _gpu_equal(pcr, oracle, one, 1.0, 1, 1)
    nonfinite = np.array([[np.nan, 0, 0], [np.inf, 1, 1], [0, -np.inf, 2]], np.float32)
    _gpu_equal(pcr, oracle, nonfinite, 1.0, 1, 10)  # three singletons
    n = 5000  # dense_cell_stress
    line = cloud(np.arange(n, dtype=np.float32) * np.float32(0.001), np.zeros(n), np.zeros(n))
    _gpu_equal(pcr, oracle, line, 0.01, 1, n + 1)
    dup = np.repeat(scenes.uniform_cube(300, 9, 0, 10), 5, axis=0)
    _gpu_equal(pcr, oracle, dup, 0.7, 1, len(dup))
    lattice = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    _gpu_equal(pcr, oracle, lattice, 1.0, 1, len(lattice))   # d == r exactly between lattice neighbours: one component
    _gpu_equal(pcr, oracle, lattice, 0.999, 1, len(lattice))  # all singletons
    far = np.vstack([scenes.uniform_cube(2000, 3, 0, 10), [[3e6, -2e6, 1e6]]]).astype(np.float32)
    _gpu_equal(pcr, oracle, far, 0.5, 1, len(far))
    huge = np.array([[0, 0, 0], [3e9, 3e9, 3e9], [3.1e9, 3e9, 3e9], [-4e12, 0, 0]], np.float32)  # saturated i32 keys
    _gpu_equal(pcr, oracle, huge, 1e-3, 1, 10)


@pytest.mark.gpu
def test_gpu_kitti_scene_full_size(pcr, oracle):
    """The clustering step of the KITTI pipeline (examples/python/kitti_obstacle_detection.py): 122 K points."""
    pts = scenes.kitti_scene()
    for thr, mn, mx in ((0.5, 30, 25000), (0.3, 10, 200000)):
        _gpu_equal(pcr, oracle, pts, thr, mn, mx)


@pytest.mark.gpu
def test_gpu_determinism(pcr):  # cluster_differential.rs:330-360
    pts = scenes.uniform_cube(20000, 77, 0, 30)
    c = pcr.PointCloud.from_numpy(pts)
    first = pcr.cluster_arrays(c, 0.8, 1, len(pts))
    for _ in range(10):
        again = pcr.cluster_arrays(c, 0.8, 1, len(pts))
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])


def test_oracle_against_scipy_components(oracle):
    """Independent pin: connected components of the r-ball graph from scipy (f64 distances).  Clouds are drawn so that no
    pair sits within 1e-4 of the threshold, where f32 and f64 could disagree."""
    scipy_spatial = pytest.importorskip("scipy.spatial")
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    rng = np.random.default_rng(12)
    checked = 0
    for _ in range(40):
        n = int(rng.integers(50, 1500))
        thr = float(rng.uniform(0.5, 4.0))
        pts = rng.uniform(-25, 25, (n, 3)).astype(np.float32)
        tree = scipy_spatial.cKDTree(pts.astype(np.float64))
        near = tree.query_pairs(thr * (1 + 1e-4), output_type="ndarray")
        sure = tree.query_pairs(thr * (1 - 1e-4), output_type="ndarray")
        if len(near) != len(sure):
            continue  # a pair too close to the threshold: skip this draw
        checked += 1
        g = coo_matrix((np.ones(len(sure)), (sure[:, 0], sure[:, 1])), shape=(n, n))
        _, lab = connected_components(g, directed=False)
        comps = {}
        for i, l in enumerate(lab):
            comps.setdefault(l, []).append(i)
        want = sorted(comps.values(), key=lambda c: (-len(c), c[0]))
        got = as_lists(oracle.euclidean_cluster(pts, thr, 1, n))
        assert got == want
    assert checked >= 20
