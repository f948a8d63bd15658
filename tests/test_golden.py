"""Golden vectors (tests/golden/kitti_small.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitti_small.npz"))


def _ang(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), np.abs(np.sum(a * b, axis=1)))


def test_oracle_reproduces_golden(oracle):
    t = oracle.Tree(G["pts"])
    idx, dist, cnt = t.knn_batch(G["q"], 11)
    assert np.array_equal(idx, G["knn11_idx"]) and np.array_equal(dist.view(np.uint32), G["knn11_dist"].view(np.uint32))
    assert np.array_equal(cnt, G["knn11_cnt"])
    keep, mean, stats = oracle.sor(G["pts"], 10, 1.0)
    assert np.array_equal(keep, G["sor_keep"]) and np.array_equal(mean.view(np.uint32), G["sor_mean"].view(np.uint32))
    assert np.array_equal(stats.view(np.uint32), G["sor_stats"].view(np.uint32))
    assert np.array_equal(oracle.ror(G["pts"], 0.5, 5), G["ror_keep"])
    assert np.array_equal(oracle.normals(G["pts"], 20).view(np.uint32), G["normals"].view(np.uint32))
    r = oracle.icp_point_to_plane(G["icp_src"], G["icp_tgt"], G["icp_tgt_normals"], 30, 0.0)
    assert np.array_equal(r.rotation, G["p2l_R"]) and np.array_equal(r.translation, G["p2l_t"])


@pytest.mark.gpu
def test_cuda_matches_golden(pcr):
    cloud = pcr.PointCloud.from_numpy(G["pts"])
    tree = pcr.KdTree(cloud, 11)
    idx, dist, cnt = tree.knn(G["q"], 11)
    assert np.array_equal(idx, G["knn11_idx"]) and np.array_equal(dist.view(np.uint32), G["knn11_dist"].view(np.uint32))
    idx, dist, cnt = tree.knn(G["q"], 40)
    assert np.array_equal(idx, G["knn40_idx"]) and np.array_equal(dist.view(np.uint32), G["knn40_dist"].view(np.uint32))
    assert np.array_equal(tree.radius_count(G["q"], 0.75), G["radius_count"])
    keep, kept, mean, stats = pcr.sor_mask(cloud, 10, 1.0, want_mean=True)
    assert np.array_equal(keep, G["sor_keep"]) and np.array_equal(mean.view(np.uint32), G["sor_mean"].view(np.uint32))
    assert np.array_equal(stats.view(np.uint32), G["sor_stats"].view(np.uint32))
    assert np.array_equal(pcr.ror_mask(cloud, 0.5, 5)[0], G["ror_keep"])
    assert _ang(pcr.normals_array(cloud, 20), G["normals"]).max() < 1e-4  # tolerance: 1e-4 rad (north_star)
    src, tgt = pcr.PointCloud.from_numpy(G["icp_src"]), pcr.PointCloud.from_numpy(G["icp_tgt"])
    r = pcr.icp_point_to_point(src, tgt, 30, 0.0)
    assert r.num_iterations == int(G["p2p_iters"])
    assert np.allclose(r.rotation, G["p2p_R"], atol=1e-4) and np.allclose(r.translation, G["p2p_t"], atol=1e-4)
    tgt.normals = G["icp_tgt_normals"]
    r = pcr.icp_point_to_plane(src, tgt, 30, 0.0)
    assert r.num_iterations == int(G["p2l_iters"])
    assert np.allclose(r.rotation, G["p2l_R"], atol=1e-4) and np.allclose(r.translation, G["p2l_t"], atol=1e-4)
    assert abs(r.rmse - float(G["p2l_rmse"])) < 1e-5
