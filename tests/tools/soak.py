"""Randomised soak of the newer paths against the oracle (cluster, voxel, RANSAC, fused SOR->normals, KNN k<=12 pruned
walk, warp-pruned deferred levels).  usage: python tests/tools/soak.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
T = min(16, os.cpu_count() or 1)
t0 = time.time()
n_cases = 0
fs = pcr.Context()          # a second context marked as a frame stream: guessed voxel key boxes (hits and misses on
fs.set_frame_stream(True)   # these unrelated clouds), reused cell sizes, the coarser level queued ahead of its count
while time.time() - t0 < budget:
    kind = rng.integers(0, 4)
    n = int(rng.integers(50, 6000))
    if kind == 0:
        pts = rng.uniform(-10, 10, (n, 3))
    elif kind == 1:   # clustered blobs + sparse background (dense cells next to empty space)
        c = rng.uniform(-20, 20, (int(rng.integers(1, 6)), 3))
        pts = np.vstack([c[rng.integers(0, len(c), n)] + rng.normal(0, rng.uniform(0.01, 0.5), (n, 3)), rng.uniform(-30, 30, (n // 10 + 1, 3))])
    elif kind == 2:   # sheet with outliers
        pts = np.vstack([np.concatenate([rng.uniform(-15, 15, (n, 2)), rng.normal(0, 0.02, (n, 1))], 1), rng.uniform(-15, 15, (n // 20 + 1, 3))])
    else:             # duplicates and lattice-like ties
        base = np.round(rng.uniform(-3, 3, (n // 3 + 1, 3)) * 4) / 4
        pts = np.repeat(base, 3, axis=0)
    pts = np.ascontiguousarray(pts, np.float32)
    if rng.random() < 0.2:
        pts[rng.integers(0, len(pts), 3)] = np.nan
    cloud = pcr.PointCloud.from_numpy(pts)
    # cluster
    thr = float(rng.uniform(0.05, 2.0))
    got = pcr.euclidean_cluster(cloud, thr, 1, len(pts))
    want = [list(map(int, c)) for c in O.euclidean_cluster(pts, thr, 1, len(pts))]
    assert got == want, ("cluster", kind, n, thr)
    # voxel
    v = float(rng.uniform(0.02, 3.0))
    assert np.array_equal(pcr.voxel_downsample(cloud, v).to_numpy().view(np.uint32), O.voxel_downsample(pts, v).view(np.uint32)), ("voxel", kind, n, v)
    dv = pcr.DeviceCloud.from_numpy(pts, fs).voxel_downsample(v)
    assert np.array_equal(dv.to_numpy().view(np.uint32), O.voxel_downsample(pts, v).view(np.uint32)), ("voxel, frame stream", kind, n, v)
    # ransac
    smp = pcr.draw_plane_samples(len(pts), int(rng.integers(1, 80)), int(rng.integers(1 << 30)))
    r = pcr.ransac_plane_samples(cloud, float(rng.uniform(0.01, 0.5)) if False else 0.1, smp)
    m, inl = O.ransac_plane_samples(pts, 0.1, smp)
    assert np.array_equal(np.array(r.normal + [r.d], np.float32).view(np.uint32), m.view(np.uint32)) and r.inliers == inl.tolist(), ("ransac", kind, n)
    # fused SOR -> normals vs two steps
    ks, kn, std = int(rng.integers(1, 31)), int(rng.integers(1, 31)), float(rng.uniform(0, 2))
    keep, nrm, kept = pcr.sor_normals_batch(pts, np.array([0, len(pts)], np.uint64), ks, std, kn)
    o_keep, o_mean, _ = O.sor(pts, ks, std, threads=T)
    assert np.array_equal(keep, o_keep), ("sor", kind, n, ks, std)
    sel = np.nonzero(o_keep)[0]
    if len(sel):
        two = pcr.normals_array(pcr.PointCloud.from_numpy(np.ascontiguousarray(pts[sel])), kn)
        assert np.array_equal(nrm[sel].view(np.uint32), two.view(np.uint32)), ("fused normals", kind, n, ks, kn, std)
    ds = pcr.DeviceCloud.from_numpy(pts, fs).sor_normals(ks, std, kn)
    assert np.array_equal(ds.to_numpy().view(np.uint32), pts[sel].view(np.uint32)), ("device sor_normals points", kind, n, ks, kn, std)
    if len(sel):
        assert np.array_equal(ds.normals_to_numpy().view(np.uint32), nrm[sel].view(np.uint32)), ("device sor_normals normals", kind, n, ks, kn, std)
    if len(dv) > 1:  # the box the voxel step hands to the index build
        a = dv.sor_normals(ks, std, kn)
        b = pcr.DeviceCloud.from_numpy(dv.to_numpy()).sor_normals(ks, std, kn)
        assert np.array_equal(a.to_numpy().view(np.uint32), b.to_numpy().view(np.uint32)) and \
            np.array_equal(a.normals_to_numpy().view(np.uint32), b.normals_to_numpy().view(np.uint32)), ("voxel -> sor_normals", kind, n, v, ks, kn)
    # KNN lists (k <= 12 uses the pruned walk; far queries exercise the deferred, warp-pruned levels)
    k = int(rng.integers(1, 13))
    q = np.vstack([pts[rng.integers(0, len(pts), 200)], rng.uniform(-60, 60, (50, 3))]).astype(np.float32)
    q = q[np.isfinite(q).all(1)]
    idx, dist, cnt = pcr.KdTree(cloud, k).knn(q, k)
    o_idx, o_dist, o_cnt = O.Tree(pts).knn_batch(q, k, threads=T)
    assert np.array_equal(cnt, o_cnt) and np.array_equal(idx, o_idx) and np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32)), ("knn", kind, n, k)
    n_cases += 1
print(f"soak ok: {n_cases} random cases in {time.time() - t0:.1f} s")
