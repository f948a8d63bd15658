"""Developer timing script (not the judged bench): host-API wall times of the main calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3, np.min(ts) * 1e3


ctx = pcr.default_context()
kitti = pcr.PointCloud.from_numpy(scenes.kitti_scene())
l0 = ctx.launch_count
print("SOR 122K k=10: med %.3f ms min %.3f ms" % timeit(lambda: pcr.sor_mask(kitti, 10, 1.0)), "launches/call", (ctx.launch_count - l0) / 13)
print("normals 122K k=20: med %.3f ms min %.3f ms" % timeit(lambda: pcr.normals_array(kitti, 20)))
t = pcr.KdTree(kitti, 11)
print("index info", t.info())
q = kitti.to_numpy()
print("knn 122K k=11 (host api): med %.3f ms min %.3f ms" % timeit(lambda: t.knn(q, 11), 5, 2))
aer_np = scenes.aerial_scene()
aer = pcr.PointCloud.from_numpy(aer_np)
print("normals 241K k=20: med %.3f ms min %.3f ms" % timeit(lambda: pcr.normals_array(aer, 20)))
print("ROR 241K r=2 min5: med %.3f ms min %.3f ms" % timeit(lambda: pcr.ror_mask(aer, 2.0, 5)))
cube = pcr.PointCloud.from_numpy(scenes.uniform_cube(100000))
print("SOR cube100K k=10: med %.3f ms min %.3f ms" % timeit(lambda: pcr.sor_mask(cube, 10, 1.0)))
n_icp = int(os.environ.get("ICP_SCALE_PCT", "10"))
tgt_np = scenes.aerial_scene(42, 0.415 * n_icp / 100)
from oracle import oracle as O
src_np = O.apply_transform(tgt_np, scenes.rot_z(0.05), [0.3, -0.2, 0.1])
tgt = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt_np), 20)
src = pcr.PointCloud.from_numpy(src_np)
for it in (30,):
    med, mn = timeit(lambda: pcr.icp_point_to_plane(src, tgt, it, 0.0), 3, 1)
    r = pcr.icp_point_to_plane(src, tgt, it, 0.0)
    print(f"ICP p2plane {len(src)} pts {it} iters: med {med:.3f} ms ({med/it:.3f} ms/iter) rmse {r.rmse:.5f} t {r.translation}")
