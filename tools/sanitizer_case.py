"""Small end-to-end case for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
pts = scenes.kitti_scene(3, (3000, 150, 30, 120))
far = np.random.default_rng(0).uniform(-300, 300, (40, 3)).astype(np.float32)
pts = np.vstack([pts, far]).astype(np.float32)
c = pcr.PointCloud.from_numpy(pts)
keep, kept, md, st = pcr.sor_mask(c, 10, 1.0, want_mean=True)
n = pcr.normals_array(c, 20)
t = pcr.KdTree(c, 40)
t.knn(pts[:500], 40); t.knn(pts[:500], 7); t.radius_search(pts[:300], 0.7)
off = np.array([0, 1500, 1500, 3340], np.uint64)
pcr.sor_normals_batch(pts, off, 10, 1.0, 20)
tgt = scenes.hemisphere(3000, 1)
src = (tgt @ scenes.rot_z(0.03).T + np.array([0.1, 0.0, -0.05], np.float32)).astype(np.float32)
tc = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt), 15)
print(pcr.icp_point_to_plane(pcr.PointCloud.from_numpy(np.ascontiguousarray(src)), tc, 10, 1e-6))
print(pcr.icp_point_to_point(pcr.PointCloud.from_numpy(np.ascontiguousarray(src)), tc, 10, 1e-6))
v = pcr.voxel_downsample(c, 0.3)
d = pcr.DeviceCloud.from_numpy(pts)
o = d.voxel_downsample(0.05).sor_normals(10, 1.0, 20)
print("device pipeline", o.len(), "voxels", v.len())
print("sanitizer case done", kept)
