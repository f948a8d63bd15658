import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
def timeit(fn, reps=8, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3
kitti = pcr.PointCloud.from_numpy(scenes.kitti_scene())
aer = pcr.PointCloud.from_numpy(scenes.aerial_scene())
cube = pcr.PointCloud.from_numpy(scenes.uniform_cube(100000))
hemi = pcr.PointCloud.from_numpy(scenes.hemisphere(200000, 1, 50.0))
print(os.environ.get("PCR_OCC_SCALE"), "kitti sor10 %.3f nrm20 %.3f | aerial nrm20 %.3f nrm15 %.3f | cube sor10 %.3f nrm10 %.3f | hemi nrm20 %.3f" % (
    timeit(lambda: pcr.sor_mask(kitti, 10, 1.0)), timeit(lambda: pcr.normals_array(kitti, 20)),
    timeit(lambda: pcr.normals_array(aer, 20)), timeit(lambda: pcr.normals_array(aer, 15)),
    timeit(lambda: pcr.sor_mask(cube, 10, 1.0)), timeit(lambda: pcr.normals_array(cube, 10)),
    timeit(lambda: pcr.normals_array(hemi, 20))))
