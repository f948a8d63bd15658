#!/usr/bin/env python3
"""Host-side timeline of the BASELINE configs[1] step (PCR_TRACE marks inside the library + the Python call
boundaries).  Usage: PCR_TRACE=1 python tools/trace_step.py"""
import os
import sys
import time

os.environ.setdefault("PCR_TRACE", "1")
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bench  # noqa: E402
import pointclouds_rs_b200 as pcr  # noqa: E402


def main():
    raw = bench.make_frames(0, 1)[0]
    ctx = pcr.Context(device=0)
    ctx.set_frame_stream(True)
    d = pcr.DeviceCloud.from_numpy(raw, ctx)
    py = []
    for i in range(30):
        t0 = time.perf_counter()
        v = d.voxel_downsample(bench.VOXEL)
        t1 = time.perf_counter()
        o = v.sor_normals(bench.K_SOR, bench.STD_MUL, bench.K_NORMALS)
        t2 = time.perf_counter()
        v.free()
        o.free()
        t3 = time.perf_counter()
        py.append((t1 - t0, t2 - t1, t3 - t2))
    last = py[-10:]
    print("python side, mean of last 10 steps (us): voxel %.1f  sor_normals %.1f  frees %.1f  total %.1f" % (
        1e6 * sum(p[0] for p in last) / 10, 1e6 * sum(p[1] for p in last) / 10, 1e6 * sum(p[2] for p in last) / 10,
        1e6 * sum(sum(p) for p in last) / 10), file=sys.stderr)
    d.free()
    ctx.close()  # prints the last marks


if __name__ == "__main__":
    main()
