"""Multi-GPU check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Sharded point-to-plane / point-to-point ICP (source split over ranks, NCCL all-reduce of the normal
equations inside the library) must give the same result on every rank and agree with the unsharded
run on one GPU; query-sharded SOR / normals / radius outlier removal of one cloud must be bit-identical to the
unsharded call; frame-sharded SOR + normals must equal the single-process batch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import pointclouds_rs_b200 as pcr
    from pointclouds_rs_b200 import dist as pdist
    from pointclouds_rs_b200 import scenes

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pcr.Context(device=local)
    uid = pdist.share_unique_id(dist, pcr.Context.unique_id, device="cuda")
    ctx.comm_init(uid, rank, world)

    tgt_np = scenes.aerial_scene(42, 0.04)
    R = scenes.rot_z(0.01)
    src_np = np.ascontiguousarray((tgt_np @ R.T + np.array([0.3, -0.2, 0.1], np.float32)).astype(np.float32))
    solo = pcr.Context(device=local)  # no communicator: the unsharded reference run
    tgt = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt_np), 20, solo)
    b, e = pdist.split_range(len(src_np), rank, world)
    shard = pcr.PointCloud.from_numpy(np.ascontiguousarray(src_np[b:e]))
    full = pcr.PointCloud.from_numpy(src_np)
    ok = True
    for name, fn in (("p2plane", pcr.icp_point_to_plane), ("p2p", pcr.icp_point_to_point)):
        for iters, tol in ((12, 0.0), (50, 1e-5)):
            r_sh = fn(shard, tgt, iters, tol, ctx=ctx)
            r_one = fn(full, tgt, iters, tol, ctx=solo)
            v = torch.tensor(r_sh.rotation + [r_sh.translation] + [[r_sh.rmse, r_sh.fitness, float(r_sh.num_iterations)]], dtype=torch.float64, device="cuda")
            gathered = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(gathered, v)
            same_everywhere = all(torch.equal(g, gathered[0]) for g in gathered)
            close = (np.allclose(r_sh.rotation, r_one.rotation, atol=1e-5) and np.allclose(r_sh.translation, r_one.translation, atol=1e-5)
                     and abs(r_sh.rmse - r_one.rmse) < 1e-5 and abs(r_sh.fitness - r_one.fitness) < 1e-6
                     and abs(r_sh.num_iterations - r_one.num_iterations) <= (0 if tol == 0.0 else 2))
            if rank == 0:
                print(f"{name} iters={iters} tol={tol}: sharded {r_sh} vs single {r_one}: identical on all ranks={same_everywhere} close={close}", flush=True)
            ok = ok and same_everywhere and close

    # query sharding of ONE cloud (BASELINE configs[2] shape): every rank gets the whole cloud, searches
    # its share, the library merges over NCCL -> every rank holds the full result, bit-identical to the unsharded call
    aer = scenes.aerial_scene(42, 0.03)
    aer[5] = [np.nan, 0, 0]  # a non-finite point (owned by rank 0)
    cloud = pcr.PointCloud.from_numpy(aer)
    ctx.set_query_sharding(True)
    keep_s, kept_s, mean_s, stats_s = pcr.sor_mask(cloud, 10, 1.0, ctx=ctx, want_mean=True)
    nrm_s = pcr.normals_array(cloud, 20, ctx=ctx)
    ror_s, _ = pcr.ror_mask(cloud, 2.0, 5, ctx=ctx)
    ctx.set_query_sharding(False)
    keep_1, kept_1, mean_1, stats_1 = pcr.sor_mask(cloud, 10, 1.0, ctx=solo, want_mean=True)
    nrm_1 = pcr.normals_array(cloud, 20, ctx=solo)
    ror_1, _ = pcr.ror_mask(cloud, 2.0, 5, ctx=solo)
    shard_ok = (np.array_equal(mean_s.view(np.uint32), mean_1.view(np.uint32)) and np.array_equal(stats_s.view(np.uint32), stats_1.view(np.uint32))
                and np.array_equal(keep_s, keep_1) and kept_s == kept_1 and np.array_equal(nrm_s.view(np.uint32), nrm_1.view(np.uint32))
                and np.array_equal(ror_s, ror_1))
    if rank == 0:
        print(f"query-sharded SOR / normals / radius outlier removal of one {len(aer)}-point cloud: bit-identical to one GPU = {shard_ok}", flush=True)
    ok = ok and shard_ok

    # frame sharding: frames dealt round-robin == the single-process batch (at least one frame per rank:
    # with 6 frames on 8 ranks two ranks had nothing to stack and the others waited for them in the all-reduce)
    frames = [scenes.kitti_scene(s, (4000, 200, 40, 80)) for s in range(max(6, 2 * world))]
    off = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    keep_all, nrm_all, kept_all = pcr.sor_normals_batch(np.vstack(frames), off, 10, 1.0, 20, ctx=solo)
    mine = pdist.deal_frames(len(frames), rank, world)
    loc = [frames[f] for f in mine]
    loff = np.concatenate([[0], np.cumsum([len(f) for f in loc])])
    local_pts = np.vstack(loc) if loc else np.zeros((0, 3), np.float32)  # a rank may hold no frame
    keep, nrm, kept = pcr.sor_normals_batch(local_pts, loff, 10, 1.0, 20, ctx=ctx)
    for j, f in enumerate(mine):
        a = slice(loff[j], loff[j + 1])
        g = slice(off[f], off[f + 1])
        fr_ok = np.array_equal(keep[a], keep_all[g]) and np.array_equal(nrm[a], nrm_all[g]) and kept[j] == kept_all[f]
        ok = ok and fr_ok
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1.0 else "FAIL", f"world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if flag.item() == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
