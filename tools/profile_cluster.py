import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
c = pcr.PointCloud.from_numpy(scenes.kitti_scene())
for _ in range(2):
    pcr.cluster_arrays(c, 0.5, 30, 25000)
    pcr.voxel_downsample(c, 0.05)
