#!/usr/bin/env python3
"""Frames in flight: end-to-end throughput of the BASELINE configs[1] pipeline (upload -> voxel 0.05 -> SOR k=10
-> normals k=20 -> download) with T host threads, each with its own context and stream, working through a
stream of frames.  One frame is latency-bound on a B200 (many short kernels, host round trips for counts), so
independent frames overlap well.  Usage: python tools/bench_inflight.py [steps_per_thread]"""
import json
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import bench  # noqa: E402
import pointclouds_rs_b200 as pcr  # noqa: E402


def worker(ctx, h_raw, h_out, n_raw, steps, start_evt, lens):
    start_evt.wait()
    for _ in range(steps):
        d = pcr.DeviceCloud.upload_block(ctx, h_raw.data_ptr(), n_raw, n_raw)
        v = d.voxel_downsample(bench.VOXEL)
        o = v.sor_normals(bench.K_SOR, bench.STD_MUL, bench.K_NORMALS)
        d.free()
        v.free()
        o.download_block(h_out.data_ptr(), n_raw, with_normals=True)
        lens.append(len(o))
        o.free()


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    raw, _ = bench.make_frame(0)
    n_raw = len(raw)
    out = {}
    for T in (1, 2, 3, 4, 6, 8):
        ctxs, bufs = [], []
        for _ in range(T):
            c = pcr.Context(device=0)
            c.set_frame_stream(True)
            ctxs.append(c)
            bufs.append((torch.from_numpy(np.ascontiguousarray(raw.T)).pin_memory(), torch.empty((6, n_raw), dtype=torch.float32).pin_memory()))
        for phase_steps in (5, steps):  # warm-up, then timed
            evt = threading.Event()
            lens = []
            th = [threading.Thread(target=worker, args=(ctxs[t], bufs[t][0], bufs[t][1], n_raw, phase_steps, evt, lens)) for t in range(T)]
            for t in th:
                t.start()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            evt.set()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        assert len(set(lens)) == 1
        same = all(np.array_equal(bufs[0][1].numpy()[:, :lens[0]], b[1].numpy()[:, :lens[0]]) for b in bufs)
        out[T] = {"points_per_s": n_raw * steps * T / dt, "ms_per_frame": dt / (steps * T) * 1e3, "identical": bool(same)}
        print(T, out[T], flush=True)
        for c in ctxs:
            c.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
