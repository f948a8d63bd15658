"""Regenerates tests/golden/kitti_small.npz from the CPU oracle.

The reference (Rust) cannot be run in this image, so the golden vectors come from the oracle, which
is itself pinned to the reference's known-answer tests (tests/test_oracle_reference_kats.py).  They
freeze today's oracle output so that (a) a later change of the oracle is noticed and (b) the GPU box
can check the CUDA path against committed vectors.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pointclouds_rs_b200 import scenes  # noqa: E402


def main():
    pts = scenes.kitti_scene(seed=123, counts=(1400, 70, 12, 40))  # 1592 points
    pts[17] = [np.nan, 0.0, 0.0]  # one non-finite point
    rng = np.random.default_rng(7)
    q = np.vstack([pts[rng.integers(0, len(pts), 150)], rng.uniform(-40, 40, (50, 3)).astype(np.float32)]).astype(np.float32)
    tree = O.Tree(pts)
    k11 = tree.knn_batch(q, 11)
    k40 = tree.knn_batch(q, 40)
    sor_keep, sor_mean, sor_stats = O.sor(pts, 10, 1.0)
    ror_keep = O.ror(pts, 0.5, 5)
    nrm = O.normals(pts, 20)
    rcount = tree.radius_count_batch(q, 0.75)
    tgt = scenes.hemisphere(400, 5, 5.0)
    src = O.apply_transform(tgt, scenes.rot_z(0.04), [0.2, -0.1, 0.05])
    tn = O.normals(tgt, 15)
    p2p = O.icp_point_to_point(src, tgt, 30, 0.0)
    p2l = O.icp_point_to_plane(src, tgt, tn, 30, 0.0)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kitti_small.npz")
    np.savez_compressed(
        out, pts=pts, q=q, knn11_idx=k11[0], knn11_dist=k11[1], knn11_cnt=k11[2], knn40_idx=k40[0], knn40_dist=k40[1],
        sor_keep=sor_keep, sor_mean=sor_mean, sor_stats=sor_stats, ror_keep=ror_keep, normals=nrm, radius_count=rcount,
        icp_tgt=tgt, icp_src=src, icp_tgt_normals=tn,
        p2p_R=p2p.rotation, p2p_t=p2p.translation, p2p_rmse=np.float32(p2p.rmse), p2p_iters=np.int64(p2p.num_iterations),
        p2l_R=p2l.rotation, p2l_t=p2l.translation, p2l_rmse=np.float32(p2l.rmse), p2l_iters=np.int64(p2l.num_iterations))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
