import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.415
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tgt_np = scenes.aerial_scene(42, scale)
src_np = np.ascontiguousarray((tgt_np @ scenes.rot_z(0.05).T + np.array([0.3, -0.2, 0.1], np.float32)).astype(np.float32))
tgt = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt_np), 20)
src = pcr.PointCloud.from_numpy(src_np)
r = pcr.icp_point_to_plane(src, tgt, iters, 0.0)
print("done", r, pcr.default_context().launch_count)
