"""voxel_downsample: the oracle pinned to the reference's tests (crates/filters/src/voxel_downsample.rs:67-121)
and the CUDA path (stable radix sort + sequential per-voxel sums) bit-exact against it."""
import numpy as np
import pytest

from pointclouds_rs_b200 import scenes


def cloud(x, y, z):
    return np.stack([np.asarray(x, np.float32), np.asarray(y, np.float32), np.asarray(z, np.float32)], 1)


# ---- oracle vs the reference's own tests ---------------------------------------------------------
def test_oracle_reduces_points(oracle):  # :73-84
    pts = cloud([0, .5, 0, .5, 0, .5, 0, .5], [0, 0, .5, .5, 0, 0, .5, .5], [0, 0, 0, 0, .5, .5, .5, .5])
    out = oracle.voxel_downsample(pts, 1.0)
    assert out.shape == (1, 3) and np.allclose(out[0], 0.25, atol=1e-6)


def test_oracle_empty_and_single(oracle):  # :87-99
    assert len(oracle.voxel_downsample(np.zeros((0, 3), np.float32), 1.0)) == 0
    out = oracle.voxel_downsample(cloud([1], [2], [3]), 1.0)
    assert out.tolist() == [[1.0, 2.0, 3.0]]


def test_oracle_never_increases_and_matches_numpy(oracle):  # :102-119 (proptest envelope) + numpy restatement
    rng = np.random.default_rng(5)
    for _ in range(30):
        n = int(rng.integers(1, 3000))
        v = float(rng.uniform(0.01, 10.0))
        pts = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
        out = oracle.voxel_downsample(pts, v)
        assert len(out) <= n
        assert np.array_equal(out, scenes.voxel_downsample_np(pts, v))


def test_oracle_rejects_bad_voxel(oracle):  # :13-16
    for v in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError):
            oracle.voxel_downsample(cloud([1], [2], [3]), v)


# ---- CUDA path ------------------------------------------------------------------------------------
def _gpu_equal(pcr, oracle, pts, v):
    pts = np.ascontiguousarray(pts, np.float32)
    got = pcr.voxel_downsample(pcr.PointCloud.from_numpy(pts), v).to_numpy()
    want = oracle.voxel_downsample(pts, v)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_reference_kats(pcr, oracle):
    pts = cloud([0, .5, 0, .5, 0, .5, 0, .5], [0, 0, .5, .5, 0, 0, .5, .5], [0, 0, 0, 0, .5, .5, .5, .5])
    _gpu_equal(pcr, oracle, pts, 1.0)
    _gpu_equal(pcr, oracle, cloud([1], [2], [3]), 1.0)
    assert len(pcr.voxel_downsample(pcr.PointCloud.from_numpy(np.zeros((0, 3), np.float32)), 1.0)) == 0
    for v in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(ValueError):
            pcr.voxel_downsample(pcr.PointCloud.from_numpy(cloud([1], [2], [3])), v)


@pytest.mark.gpu
def test_gpu_random_bit_exact(pcr, oracle):
    rng = np.random.default_rng(6)
    for _ in range(25):
        n = int(rng.integers(1, 6000))
        v = float(rng.uniform(0.01, 10.0))
        _gpu_equal(pcr, oracle, rng.uniform(-100, 100, (n, 3)), v)


@pytest.mark.gpu
def test_gpu_edge_cases(pcr, oracle):
    rng = np.random.default_rng(8)
    base = rng.uniform(-5, 5, (4000, 3)).astype(np.float32)
    with_bad = base.copy()
    with_bad[::37, 0] = np.nan
    with_bad[5::91, 2] = np.inf
    _gpu_equal(pcr, oracle, with_bad, 0.3)                      # non-finite points are skipped (:28-30)
    _gpu_equal(pcr, oracle, np.full((3, 3), np.nan, np.float32), 1.0)  # nothing finite -> empty cloud (:45-47)
    _gpu_equal(pcr, oracle, base, 100.0)                        # one or a few voxels: a 4000-long sequential f32 chain
    _gpu_equal(pcr, oracle, np.repeat(base[:500], 9, axis=0), 0.05)     # duplicated points
    far = np.vstack([base, [[3e6, -2e6, 1e6]], [[-4e7, 5e7, 9e6]]]).astype(np.float32)
    _gpu_equal(pcr, oracle, far, 0.01)                          # key box > 63 bits: three-pass (z, y, x) sort
    sat = np.vstack([base[:100], [[3e9, 3e9, 3e9]], [[3.5e9, 3e9, 3e9]], [[-1e12, 0, 0]]]).astype(np.float32)
    _gpu_equal(pcr, oracle, sat, 0.5)                           # saturated i32 keys merge, as in the reference
    _gpu_equal(pcr, oracle, (base * 1e-3).astype(np.float32), 1e-4)


@pytest.mark.gpu
def test_gpu_config2_input_full_size(pcr, oracle):
    """BASELINE configs[1]: the voxel 0.05 step in front of SOR on the 122 K KITTI-shaped frame, and a coarse grid."""
    pts = scenes.kitti_scene()
    _gpu_equal(pcr, oracle, pts, 0.05)
    _gpu_equal(pcr, oracle, pts, 0.5)
    _gpu_equal(pcr, oracle, scenes.aerial_scene(), 0.5)


@pytest.mark.gpu
def test_gpu_frame_stream_guessed_key_box(pcr, oracle):
    """A frame stream sizes the column table from the previous frame's key box (padded) without measuring the new frame;
    a point outside the box must be noticed and the frame redone the exact way.  Every frame bit-exact, hit or miss."""
    ctx = pcr.Context(device=0)
    ctx.set_frame_stream(True)
    rng = np.random.default_rng(12)
    a = scenes.kitti_scene(41, (9_000, 400, 80, 200))
    inside = a[rng.integers(0, len(a), 500)] + rng.uniform(-0.01, 0.01, (500, 3)).astype(np.float32)
    bad = a.copy()
    bad[::53, 1] = np.nan
    frames = [
        a,                                                               # measured (first frame)
        a[::-1].copy(),                                                  # hit
        np.vstack([a, [[500.0, -3.0, 1.0]]]).astype(np.float32),         # miss in x
        a,                                                               # measured again after the miss ... then hits
        np.vstack([a, [[0.0, 0.0, -90.0]]]).astype(np.float32),          # miss in z
        inside.astype(np.float32),                                       # hit, table mostly empty
        bad,                                                             # hit with non-finite points
        (a + np.float32(300.0)).astype(np.float32),                      # miss: everything outside
        np.full((5, 3), np.nan, np.float32),                             # nothing finite
        a,
    ] + [a] * 34                                                         # long enough to pass the periodic re-measure
    for i, f in enumerate(frames):
        want = oracle.voxel_downsample(f, 0.05)
        got = pcr.voxel_downsample(pcr.PointCloud.from_numpy(f), 0.05, ctx).to_numpy()
        assert np.array_equal(got, want), f"frame {i}"
        dev = pcr.DeviceCloud.from_numpy(f, ctx).voxel_downsample(0.05)
        assert np.array_equal(dev.to_numpy(), want), f"frame {i} (device cloud)"
        if len(want) > 50:                                               # the box handed to the index build must be right too
            s = dev.sor_normals(10, 1.0, 20)
            s2 = pcr.DeviceCloud.from_numpy(want).sor_normals(10, 1.0, 20)
            assert np.array_equal(s.to_numpy(), s2.to_numpy()) and np.array_equal(s.normals_to_numpy(), s2.normals_to_numpy())
    got = pcr.voxel_downsample(pcr.PointCloud.from_numpy(a), 0.2, ctx).to_numpy()   # another voxel size: the box is not reused
    assert np.array_equal(got, oracle.voxel_downsample(a, 0.2))
    ctx.close()
