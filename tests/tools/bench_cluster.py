"""Developer timing: euclidean_cluster on the KITTI-shaped frame (host API wall time + device stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
from oracle import oracle as O
pts = scenes.kitti_scene()
c = pcr.PointCloud.from_numpy(pts)
for thr, mn, mx in ((0.5, 30, 25000), (0.3, 10, 200000)):
    for _ in range(3): pcr.cluster_arrays(c, thr, mn, mx)
    ctx = pcr.default_context(); ctx.set_timing(True); ctx.get_timing()
    t0 = time.perf_counter()
    for _ in range(10): off, idx = pcr.cluster_arrays(c, thr, mn, mx)
    gpu = (time.perf_counter() - t0) / 10
    print('  device ms/call', ctx.get_timing()['other'][0] / 10); ctx.set_timing(False)
    t0 = time.perf_counter(); ref = O.euclidean_cluster(pts, thr, mn, mx); cpu = time.perf_counter() - t0
    print(f"cluster r={thr}: gpu api {gpu*1e3:.3f} ms, oracle (1 thread) {cpu*1e3:.1f} ms, clusters {len(off)-1}, largest {off[1]-off[0] if len(off)>1 else 0}")
