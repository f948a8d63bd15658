"""Developer sweep: per-stage DEVICE times (library cudaEvents) of the batch pipeline on the bench frame
and of normals on the other shapes, for the current PCR_* tuning environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
ctx = pcr.default_context()
TAGS = ["build", "knn", "knn_deferred", "sor_stats", "icp_step", "icp_solve", "knn_normals", "other"]
def stage(fn, reps=6, warm=2):
    for _ in range(warm): fn()
    ctx.set_timing(True); ctx.get_timing()
    for _ in range(reps): fn()
    tm = ctx.get_timing(); ctx.set_timing(False)
    return {t: round(v[0] / reps * 1e3) for t, v in tm.items() if v[0] > 0}
pts = scenes.voxel_downsample_np(scenes.kitti_scene(), 0.05)
off = np.array([0, len(pts)], np.uint64)
env = {k: v for k, v in os.environ.items() if k.startswith("PCR_")}
print(env)
r = stage(lambda: pcr.sor_normals_batch(pts, off, 10, 1.0, 20)); print("  kitti batch us", r, "sum", sum(r.values()))
for name, c, k in (("aerial", scenes.aerial_scene(), 20), ("cube100k", scenes.uniform_cube(100000), 10), ("hemi200k", scenes.hemisphere(200000, 1, 50.0), 20)):
    cl = pcr.PointCloud.from_numpy(c)
    r = stage(lambda: pcr.normals_array(cl, k)); print(f"  {name} normals k={k} us", r, "sum", sum(r.values()))
