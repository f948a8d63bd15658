"""Developer timing: BASELINE configs[4] on one GPU -- 100 KITTI-shaped frames x 80 000 points, SOR k=10 + normals k=20
in ONE pcr_sor_normals_batch call (frames are an extra axis of the cell table), device stages and host-API wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pointclouds_rs_b200 as pcr
from pointclouds_rs_b200 import scenes
F = int(sys.argv[1]) if len(sys.argv) > 1 else 100
frames = [scenes.kitti_scene(seed=s, counts=(70_600, 3_530, 590, 1_750)) for s in range(F)]
pts = np.ascontiguousarray(np.vstack(frames), np.float32)
off = np.cumsum([0] + [len(f) for f in frames]).astype(np.uint64)
ctx = pcr.default_context()
for _ in range(2): pcr.sor_normals_batch(pts, off, 10, 1.0, 20)
ctx.set_timing(True); ctx.get_timing()
t0 = time.perf_counter()
R = 3
for _ in range(R): keep, nrm, kept = pcr.sor_normals_batch(pts, off, 10, 1.0, 20)
wall = (time.perf_counter() - t0) / R
tm = ctx.get_timing(); ctx.set_timing(False)
dev = {k: round(v[0] / R, 3) for k, v in tm.items() if v[0] > 0}
print(f"{F} frames, {len(pts)} points: host-API wall {wall*1e3:.2f} ms = {len(pts)/wall/1e6:.1f} M points/s; device stages ms {dev} sum {sum(dev.values()):.2f} -> {len(pts)/sum(dev.values())/1e3:.1f} M points/s device")
