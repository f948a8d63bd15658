#!/usr/bin/env python3
"""Repeated device-resident 100-frame batch calls with per-call wall times; with PCR_TRACE=1 PCR_TRACE_SLOW_US=<t> the
library prints the host-side marks of the slow calls.  Usage: PCR_TRACE=1 PCR_TRACE_SLOW_US=16000 python tools/batch_outliers.py [reps]"""
import os
import sys
import time

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import pointclouds_rs_b200 as pcr  # noqa: E402
from pointclouds_rs_b200 import scenes  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
fr = [scenes.kitti_scene(seed=f, counts=scenes.KITTI_COUNTS["frame80k"]) for f in range(100)]
off = np.cumsum([0] + [len(f) for f in fr]).astype(np.uint64)
n = int(off[-1])
pts = np.vstack(fr)
ctx = pcr.Context(device=0)
vp = np.zeros(3, np.float32)
d_in = torch.from_numpy(np.ascontiguousarray(pts.T)).cuda()
d_keep = torch.empty(n, dtype=torch.uint8, device="cuda")
d_nrm = torch.empty((3, n), dtype=torch.float32, device="cuda")
flush = torch.empty(bench.L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
times = []
for i in range(reps + 2):
    flush.fill_(i & 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pcr.sor_normals_batch_raw(ctx, d_in[0].data_ptr(), d_in[1].data_ptr(), d_in[2].data_ptr(), n, off, bench.K_SOR, bench.STD_MUL, bench.K_NORMALS, vp,
                              d_keep.data_ptr(), d_nrm[0].data_ptr(), d_nrm[1].data_ptr(), d_nrm[2].data_ptr(), device=True)
    ctx.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
print("ms per call:", " ".join("%.1f" % t for t in times[2:]))
print("median %.2f  min %.2f  max %.2f" % (float(np.median(times[2:])), min(times[2:]), max(times[2:])))
ctx.close()
