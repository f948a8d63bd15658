"""Developer timing: the KITTI pipeline of configs[1] three ways -- host-pointer calls (one PCIe round trip per step),
device-resident DeviceCloud, and the CPU oracle -- plus voxel / cluster alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
from oracle import oracle as O
def med(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e3
pts = scenes.kitti_scene()
c = pcr.PointCloud.from_numpy(pts)
def host_pipe():
    v = pcr.voxel_downsample(c, 0.05); s = pcr.statistical_outlier_removal(v, 10, 1.0); return pcr.normals_array(s, 20)
def dev_pipe():
    d = pcr.DeviceCloud.from_cloud(c); n = d.voxel_downsample(0.05).statistical_outlier_removal(10, 1.0).estimate_normals(20); return n.normals_to_numpy()
d0 = pcr.DeviceCloud.from_cloud(c)
def dev_resident():
    return d0.voxel_downsample(0.05).statistical_outlier_removal(10, 1.0).estimate_normals(20)
print("voxel 0.05 (host api)       %.3f ms" % med(lambda: pcr.voxel_downsample(c, 0.05)))
print("voxel 0.05 (device cloud)   %.3f ms" % med(lambda: d0.voxel_downsample(0.05)))
print("cluster 0.5 (host api)      %.3f ms" % med(lambda: pcr.cluster_arrays(c, 0.5, 30, 25000)))
print("pipeline host-pointer calls %.3f ms" % med(host_pipe))
print("pipeline upload+device+dl   %.3f ms" % med(dev_pipe))
print("pipeline device resident    %.3f ms" % med(dev_resident))
t0 = time.perf_counter(); v = O.voxel_downsample(pts, 0.05); t1 = time.perf_counter()
k, _, _ = O.sor(v, 10, 1.0, threads=1); t2 = time.perf_counter()
O.normals(v[k.astype(bool)], 20, threads=os.cpu_count()); t3 = time.perf_counter()
print("oracle: voxel %.1f ms, SOR (1 thread) %.1f ms, normals (%d threads) %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, os.cpu_count(), (t3-t2)*1e3))
