import os, sys
os.environ["PCR_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
for name, pts, k in (("kitti", scenes.kitti_scene(), 20), ("aerial", scenes.aerial_scene(), 20), ("cube", scenes.uniform_cube(100000), 10)):
    c = pcr.PointCloud.from_numpy(pts)
    t = pcr.KdTree(c, k)
    print(name, t.info(), flush=True)
    pcr.normals_array(c, k)
