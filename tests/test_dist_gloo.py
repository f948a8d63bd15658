"""world_size-2 gloo test of the host-side multi-GPU logic (frame dealing, query split, id exchange,
max-over-ranks timing) -- the N > 1 path of bench.py without GPUs."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from pointclouds_rs_b200 import dist as pdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = pdist.share_unique_id(dist, lambda: bytes(range(128)))
        mx = pdist.max_over_ranks(dist, 1.0 + rank)
        sm = pdist.sum_over_ranks(dist, 10.0 * (rank + 1))
        mine = pdist.deal_frames(7, rank, world)
        merged = pdist.gather_frame_results(dist, [f * f for f in mine], 7, rank, world)
        q.put((rank, uid, mx, sm, mine, merged, pdist.split_range(11, rank, world)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, uid0, mx0, sm0, mine0, merged0, sl0), (r1, uid1, mx1, sm1, mine1, merged1, sl1) = res
    assert uid0 == uid1 == bytes(range(128))
    assert mx0 == mx1 == 2.0 and sm0 == sm1 == 30.0
    assert mine0 == [0, 2, 4, 6] and mine1 == [1, 3, 5]
    assert merged0 == [f * f for f in range(7)] and merged1 is None
    assert sl0 == (0, 6) and sl1 == (6, 11)


def test_partition_helpers():
    from pointclouds_rs_b200 import dist as pdist

    for n, w in ((100, 8), (7, 3), (0, 4), (5, 8)):
        owned = sorted(f for r in range(w) for f in pdist.deal_frames(n, r, w))
        assert owned == list(range(n))
        cover = [pdist.split_range(n, r, w) for r in range(w)]
        assert cover[0][0] == 0 and cover[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        sizes = [e - b for b, e in cover]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pdist.deal_frames(4, 2, 2)
