"""One warm-up + one measured call of each consumer (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
which = sys.argv[1] if len(sys.argv) > 1 else "sor"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if which == "batch":
    from pointclouds_rs_b200 import scenes as _s
    pts = _s.voxel_downsample_np(_s.kitti_scene(), 0.05)
    off = np.array([0, len(pts)], np.uint64)
    for _ in range(reps):
        pcr.sor_normals_batch(pts, off, 10, 1.0, 20)
    print("done", pcr.default_context().launch_count)
    sys.exit(0)
if which == "frame":  # the bench step on device-resident clouds: voxel -> fused SOR + normals (frame stream on)
    import bench
    ctx = pcr.Context(device=0)
    ctx.set_frame_stream(True)
    d = pcr.DeviceCloud.from_numpy(bench.make_frames(0, 1)[0], ctx)
    for _ in range(reps):
        v = d.voxel_downsample(bench.VOXEL)
        o = v.sor_normals(bench.K_SOR, bench.STD_MUL, bench.K_NORMALS)
        v.free(); o.free()
    print("done", ctx.launch_count)
    sys.exit(0)
if which in ("sor", "normals", "knn"):
    c = pcr.PointCloud.from_numpy(scenes.kitti_scene())
elif which == "aerial":
    c = pcr.PointCloud.from_numpy(scenes.aerial_scene())
for _ in range(reps):
    if which == "sor":
        pcr.sor_mask(c, 10, 1.0)
    elif which in ("normals", "aerial"):
        pcr.normals_array(c, 20)
print("done", pcr.default_context().launch_count)
