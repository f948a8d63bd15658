#!/usr/bin/env python3
"""bench.py -- the headline benchmark of the B200 KNN engine (driver contract: ONE JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): SOR + normals points/sec, k = 10 / 20.
Workload at every N (weak scaling, one process per GPU, no data-path collective): BASELINE
configs[1] -- a KITTI-shaped synthetic frame of 122 000 points (numpy PCG64, seed 42 + rank); per step

    voxel_downsample(0.05) -> statistical_outlier_removal(k = 10, std_mul = 1.0) -> estimate_normals(k = 20)
    on the kept points

A step is one pass of that pipeline over one frame; points/s counts the RAW input points.
  value : points/s, the raw frame resident in HBM (pcr_cloud_voxel_downsample + pcr_cloud_sor_normals on
          a device-resident pcr_cloud), CUDA events on the stream the kernels run on, max over ranks,
          L2 flushed between steps.
  e2e   : the same through the C ABI from PINNED host buffers: pcr_cloud_upload of x/y/z, the two
          calls, pcr_cloud_download of the kept points and their normals, all inside the timed region
          (wall clock).
  roofline     : the dominant kernel (KNN + fused normals, grid level 0), algorithmic bytes / its
                 device time measured live with cudaEvents inside the library (pcr_ctx_get_timing).
  cpu_baseline : the CPU oracle (C port of the reference path) timed on this box's host cores on a
                 bounded sample, rank 0, N = 1 only.
--impl reference times that CPU port as the reference arm (the Rust reference cannot be built in
this image: no cargo/rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_SOR, STD_MUL, K_NORMALS = 10, 1.0, 20
VOXEL = 0.05
METRIC = "sor_normals_points_per_sec"
UNIT = "points/s"
L2_FLUSH_BYTES = 256 << 20


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_frame(rank: int):
    from pointclouds_rs_b200 import scenes

    raw = np.ascontiguousarray(scenes.kitti_scene(seed=42 + rank), np.float32)
    pts = scenes.voxel_downsample_np(raw, VOXEL)  # what the voxel step hands to SOR (used for sizes and the roofline only)
    return raw, np.ascontiguousarray(pts, np.float32)


def workload_config(n_raw: int, n_in: int, world: int):
    return {
        "workload": "BASELINE configs[1]: KITTI-shaped synthetic frame 122K pts: voxel 0.05 -> SOR k=10 -> normals",
        "points_per_step_per_gpu": n_raw,
        "points_after_voxel": n_in,
        "k_sor": K_SOR,
        "std_mul": STD_MUL,
        "k_normals": K_NORMALS,
        "frames_per_step_per_gpu": 1,
        "sharding": f"frames: one independent frame per GPU per step x {world} GPU(s), no data-path collective",
        "l2": "256 MiB buffer overwritten between timed steps (outside the events); the frame itself is L2-sized",
        "seed": "numpy PCG64, 42 + rank",
        "frame_stream": "pcr_ctx_set_frame_stream(1): the cell size probed on the first frame is reused for the following frames",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2.0)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def cpu_reference_run(raw: np.ndarray, reps: int, threads_normals: int):
    """The CPU port of the reference path, threaded the way the reference is: voxel_downsample and SOR
    are serial loops (voxel_downsample.rs:24, statistical_outlier.rs:19), normals run on all cores
    (rayon par_iter, estimate.rs:42-44)."""
    from oracle import oracle as O  # the checker, used here only as the timed CPU baseline

    O.lib()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        pts = O.voxel_downsample(raw, VOXEL)
        keep, _, _ = O.sor(pts, K_SOR, STD_MUL, threads=1)
        kept = pts[keep.astype(bool)]
        O.normals(kept, K_NORMALS, threads=threads_normals)
        times.append(time.perf_counter() - t0)
    return times


def cpu_reference_all_threads(raw: np.ndarray, threads: int):
    from oracle import oracle as O

    t0 = time.perf_counter()
    pts = O.voxel_downsample(raw, VOXEL)
    keep, _, _ = O.sor(pts, K_SOR, STD_MUL, threads=threads)
    O.normals(pts[keep.astype(bool)], K_NORMALS, threads=threads)
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # the CPU arm runs on rank 0 only
    raw, pts = make_frame(0)
    cores = os.cpu_count() or 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_run(raw, 1, cores)
    steps = max(1, min(args.steps, 8))  # bounded: each step is ~0.3 s of CPU work
    times = cpu_reference_run(raw, steps, cores)
    t = float(np.mean(times))
    value = len(raw) / t
    t_all = cpu_reference_all_threads(raw, cores)
    sample = (f"{steps} step(s) of the full workload (one {len(raw)}-point frame per step); voxel and SOR on 1 thread (the "
              f"reference's loops are serial), normals on {cores} threads (reference: rayon); C port of the reference path (oracle/)")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(len(raw), len(pts), 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "all_threads_value": len(raw) / t_all},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-icp", action="store_true", help="skip the secondary ICP measurement")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import pointclouds_rs_b200 as pcr
    from pointclouds_rs_b200 import dist as pdist

    stream = torch.cuda.current_stream()
    ctx = pcr.Context(device=local_rank, stream=stream.cuda_stream)
    ctx.set_frame_stream(True)  # the steps are consecutive frames of one sensor: the probed cell size is reused (config.frame_stream)

    raw, pts = make_frame(rank)
    n_raw, n = len(raw), len(pts)

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    # the raw frame resident in HBM (device arm) and in pinned host memory (end-to-end arm)
    h_raw = torch.from_numpy(np.ascontiguousarray(raw.T)).pin_memory()  # (3, n_raw): x | y | z
    h_out = torch.empty((6, n_raw), dtype=torch.float32).pin_memory()   # kept x, y, z, nx, ny, nz
    d_raw = pcr.DeviceCloud.upload_raw(ctx, h_raw[0].data_ptr(), h_raw[1].data_ptr(), h_raw[2].data_ptr(), n_raw)
    last = {}

    def pipeline(cloud):
        v = cloud.voxel_downsample(VOXEL)
        o = v.sor_normals(K_SOR, STD_MUL, K_NORMALS)
        v.free()
        return o

    def step_device():
        o = pipeline(d_raw)
        if "dev" in last:
            last["dev"].free()
        last["dev"] = o

    def step_e2e():
        # x | y | z rows of one pinned block that nobody writes: the copy is queued, the voxel step runs right behind it
        d = pcr.DeviceCloud.upload_block(ctx, h_raw.data_ptr(), n_raw, n_raw, wait=False)
        o = pipeline(d)
        d.free()
        o.download_block(h_out.data_ptr(), n_raw, with_normals=True)           # x | y | z | nx | ny | nz rows
        last["e2e_len"] = len(o)
        o.free()

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows every scratch buffer to its final size) -------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()
    torch.cuda.synchronize()
    n_kept = len(last["dev"])

    # ---- timed region 1: device-resident ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    ctx.set_timing(True)
    ctx.get_timing()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    sampler.start()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 (outside the timed events)
        starts[i].record(stream)
        step_device()
        stops[i].record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    stage = ctx.get_timing()
    ctx.set_timing(False)
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))

    # ---- timed region 2: end to end through the host-pointer C ABI --------------------------------
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()

    dev_ms_max = pdist.max_over_ranks(dist if distributed else None, dev_ms, "cuda")
    e2e_s_max = pdist.max_over_ranks(dist if distributed else None, e2e_s, "cuda")
    n_total = pdist.sum_over_ranks(dist if distributed else None, float(n_raw), "cuda")

    # parity spot check of the timed configuration against the e2e arm (same inputs, same results)
    dev_pts, dev_nrm = last["dev"].to_numpy(), last["dev"].normals_to_numpy()
    m = last["e2e_len"]
    same = (m == len(dev_pts) and np.array_equal(dev_pts.T, h_out[0:3, :m].numpy()) and np.array_equal(dev_nrm.T, h_out[3:6, :m].numpy()))

    value = n_total * args.steps / (dev_ms_max * 1e-3)
    e2e_value = n_total * args.steps / e2e_s_max

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peak, peak_src = load_peaks()
    knn_n_ms, knn_n_cnt = stage["knn_normals"]
    knn_s_ms, knn_s_cnt = stage["knn"]
    # The dominant kernel of the fused pipeline is the level-0 KNN of the SOR pass: it searches
    # K = max(k_sor, k_normals) + 1 neighbours once, writes the SOR mean distance and keeps the neighbour
    # lists the normals are later computed from.  Algorithmic bytes per query: read the
    # cell-sorted float4 (16), write the mean distance (4), the K list entries (4 K) and the list length (1).
    K_LIST = max(K_SOR, K_NORMALS) + 1
    bytes_knn = n * (16 + 4 + 4 * K_LIST + 1)
    dur_knn = knn_s_ms / max(knn_s_cnt, 1) * 1e-3
    achieved = bytes_knn / dur_knn / 1e9 if dur_knn > 0 else 0.0
    props = torch.cuda.get_device_properties(local_rank)
    roofline = {
        "bound": "hbm",
        "kernel": f"knn_thread_kernel<{K_LIST},3> (grid KNN K={K_LIST}: SOR mean distance + neighbour lists kept for the normals; level 0, thread per query)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": None,
        "algorithmic_bytes_per_launch": bytes_knn, "avg_launch_ms": dur_knn * 1e3,
        "note": "a 119 K-point frame is L2-resident: the kernel is bound by instruction issue and per-warp latency, not by HBM (DESIGN.md)",
        # (the library's catch-all tag "other" holds only the voxel step in this pipeline)
        "stage_ms_per_step": {("voxel" if k == "other" else k): v[0] / args.steps for k, v in stage.items() if v[1]},
        "normals_from_lists_kernel": {"avg_launch_ms": knn_n_ms / max(knn_n_cnt, 1), "algorithmic_bytes_per_launch": n_kept * (16 + 4 * K_LIST + 1 + 12)},
        "device": {"name": torch.cuda.get_device_name(local_rank), "sm_count": props.multi_processor_count, "l2_bytes": props.L2_cache_size},
    }
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline["traffic"] = json.load(f).get("dominant_kernel_dram_bytes_per_launch")
        except Exception:
            pass

    extra = {}
    if rank == 0 and world == 1 and not args.no_icp:
        try:
            extra["icp"] = icp_measurement(pcr, ctx)
        except Exception as ex:  # the headline must not die on the secondary measurement
            extra["icp"] = {"error": str(ex)}
        try:
            extra["frames_in_flight"] = frames_in_flight_measurement(pcr, local_rank, h_raw, n_raw, h_out, last["e2e_len"])
        except Exception as ex:
            extra["frames_in_flight"] = {"error": str(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times = cpu_reference_run(raw, 3, cores)
        t_all = cpu_reference_all_threads(raw, cores)
        cpu_baseline = {
            "value": n_raw / float(np.mean(times)), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"3 steps of the same frame ({n_raw} points): voxel and SOR on 1 thread (the reference's loops are serial), "
                       f"normals on {cores} threads (reference: rayon); C port of the reference (oracle/), not the Rust build"),
            "all_threads_value": n_raw / t_all,
        }

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(len(raw), n, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * 4 * n_raw, "d2h_bytes_per_step": 6 * 4 * n_kept,
                    "ms_per_step": e2e_s_max / args.steps * 1e3,
                    "timer": "wall clock around pcr_cloud_upload_block_nowait -> voxel -> sor_normals -> pcr_cloud_download_block, pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kept_points": n_kept,
            "device_and_e2e_results_identical": same,
        }
        line.update(extra)
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def frames_in_flight_measurement(pcr, device, h_raw, n_raw, h_ref, m_ref, flights=(2, 4), steps=60):
    """Secondary figure (NOT the headline): the same end-to-end step with several frames in flight.  One frame is
    latency-bound (short kernels, host round trips for counts), and contexts are independent, so T host threads --
    each with its own context, stream and pinned output block -- overlap their frames on the one GPU."""
    import threading

    import torch

    def worker(ctx, out, n_steps, go, lens):
        go.wait()
        for _ in range(n_steps):
            d = pcr.DeviceCloud.upload_block(ctx, h_raw.data_ptr(), n_raw, n_raw, wait=False)
            v = d.voxel_downsample(VOXEL)
            o = v.sor_normals(K_SOR, STD_MUL, K_NORMALS)
            d.free()
            v.free()
            o.download_block(out.data_ptr(), n_raw, with_normals=True)
            lens.append(len(o))
            o.free()

    res = {"note": "T host threads x (context, stream), same frame and calls as e2e; wall clock; results compared with the e2e arm"}
    for T in flights:
        ctxs = [pcr.Context(device=device) for _ in range(T)]
        outs = [torch.empty((6, n_raw), dtype=torch.float32).pin_memory() for _ in range(T)]
        for c in ctxs:
            c.set_frame_stream(True)
        dt, lens = 0.0, []
        for n_steps in (5, steps):  # warm-up, then timed
            go, lens = threading.Event(), []
            th = [threading.Thread(target=worker, args=(ctxs[t], outs[t], n_steps, go, lens)) for t in range(T)]
            for t in th:
                t.start()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            go.set()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        same = all(m == m_ref for m in lens) and all(np.array_equal(o[:, :m_ref].numpy(), h_ref[:, :m_ref].numpy()) for o in outs)
        res[str(T)] = {"value": n_raw * steps * T / dt, "unit": UNIT, "ms_per_frame": dt / (steps * T) * 1e3, "identical_to_e2e": bool(same)}
        for c in ctxs:
            c.close()
    return res


def icp_measurement(pcr, ctx):
    """BASELINE configs[3] at N = 1: point-to-plane ICP, two synthetic scans of 1 M points, 30 iterations."""
    from pointclouds_rs_b200 import scenes

    tgt_np = scenes.aerial_scene(42, 0.415)
    R = scenes.rot_z(0.05)
    src_np = (tgt_np @ R.T + np.array([0.3, -0.2, 0.1], np.float32)).astype(np.float32)
    tgt = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt_np), 20, ctx)
    src = pcr.PointCloud.from_numpy(np.ascontiguousarray(src_np))
    pcr.icp_point_to_plane(src, tgt, 30, 0.0, ctx=ctx)  # warm-up
    ctx.set_timing(True)
    ctx.get_timing()
    t0 = time.perf_counter()
    res = pcr.icp_point_to_plane(src, tgt, 30, 0.0, ctx=ctx)
    wall = time.perf_counter() - t0
    stage = ctx.get_timing()
    ctx.set_timing(False)
    step_ms, step_cnt = stage["icp_step"]
    return {
        "workload": "BASELINE configs[3]: point-to-plane ICP, two synthetic scans, 30 iterations, tolerance 0",
        "points": len(src), "iterations": res.num_iterations,
        "ms_per_iter_e2e": wall * 1e3 / max(res.num_iterations, 1),
        "ms_per_iter_step_kernel": step_ms / max(step_cnt, 1),
        "rmse": res.rmse, "translation": res.translation,
    }


if __name__ == "__main__":
    sys.exit(main())
