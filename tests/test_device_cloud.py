"""Device-resident clouds: every step must give exactly what the host-pointer call gives (and hence what
the oracle gives), `select` must keep order and carry normals (crates/core/src/cloud.rs:103-140)."""
import numpy as np
import pytest

from pointclouds_rs_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 100_003])
def test_rows_in_and_out_are_the_soa_arrays(pcr, n):
    """(N, 3) rows cross PCIe as they are and are split / interleaved on the device (pcr_cloud_upload_rows /
    _download_rows): same bits as the SoA transfers, NaN / inf / -0.0 included, whatever the tail of the last block."""
    rng = np.random.default_rng(n)
    pts = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    if n > 2:
        pts[1] = [np.nan, -0.0, np.inf]
    d = pcr.DeviceCloud.from_numpy(pts)                                  # rows path (float32, C-contiguous)
    s = pcr.DeviceCloud.from_cloud(pcr.PointCloud.from_numpy(pts))       # SoA path
    assert len(d) == n == len(s)
    x, y, z = (np.zeros(max(n, 1), np.float32) for _ in range(3))
    d.download_raw(x.ctypes.data, y.ctypes.data, z.ctypes.data)          # SoA download of the rows upload
    assert np.array_equal(np.stack([x[:n], y[:n], z[:n]], 1).view(np.uint32), pts.view(np.uint32))
    assert np.array_equal(s.to_numpy().view(np.uint32), pts.view(np.uint32))   # rows download of the SoA upload
    assert pcr.DeviceCloud.from_numpy(pts.astype(np.float64)).to_numpy().shape == (n, 3)   # other dtypes: the host class's path
    if n >= 256:
        dn = d.estimate_normals(8)
        nx, ny, nz = (np.zeros(n, np.float32) for _ in range(3))
        dn.download_raw(x.ctypes.data, y.ctypes.data, z.ctypes.data, nx.ctypes.data, ny.ctypes.data, nz.ctypes.data)
        assert np.array_equal(dn.normals_to_numpy().view(np.uint32), np.stack([nx, ny, nz], 1).view(np.uint32))
    assert d.normals_to_numpy() is None


def test_upload_download_select(pcr):
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5, 5, (5000, 3)).astype(np.float32)
    d = pcr.DeviceCloud.from_numpy(pts)
    assert len(d) == 5000 and not d.has_normals() and not d.is_empty()
    assert np.array_equal(d.to_numpy(), pts)
    idx = rng.integers(0, 5000, 777)
    assert np.array_equal(d.select(idx).to_numpy(), pts[idx])          # arbitrary order and repeats, like the reference
    assert len(d.select([])) == 0
    with pytest.raises(IndexError):                                    # crates/python/src/cloud.rs:56-61
        d.select([5000])
    dn = d.estimate_normals(10)
    sel = dn.select(idx)
    assert sel.has_normals() and np.array_equal(sel.normals_to_numpy(), dn.normals_to_numpy()[idx])
    block = np.ascontiguousarray(pts.T)                                 # x | y | z rows of one block: one strided transfer each way
    b = pcr.DeviceCloud.upload_block(pcr.default_context(), block.ctypes.data, block.shape[1], block.shape[1])
    assert np.array_equal(b.to_numpy(), pts)
    b2 = pcr.DeviceCloud.upload_block(pcr.default_context(), block.ctypes.data, block.shape[1], block.shape[1], wait=False)
    assert np.array_equal(b2.to_numpy(), pts)                            # (the download waits for the queued copy)
    outb = np.zeros((6, 5000 + 24), np.float32)                          # a stride larger than the cloud
    dn.download_block(outb.ctypes.data, outb.shape[1], with_normals=True)
    assert np.array_equal(outb[:3, :5000].T, pts) and np.array_equal(outb[3:, :5000].T, dn.normals_to_numpy())
    with pytest.raises(ValueError):
        d.download_block(outb.ctypes.data, outb.shape[1], with_normals=True)   # no normals on this cloud
    empty = pcr.DeviceCloud.from_numpy(np.zeros((0, 3), np.float32))
    assert len(empty) == 0 and empty.to_numpy().shape == (0, 3)


def test_pipeline_matches_host_api_and_oracle(pcr, oracle):
    """BASELINE configs[1] end to end on the device: voxel 0.05 -> SOR k=10 -> normals k=20, then clustering."""
    pts = scenes.kitti_scene(5, (30_000, 1500, 250, 650))
    d = pcr.DeviceCloud.from_numpy(pts)
    v = d.voxel_downsample(0.05)
    o_v = oracle.voxel_downsample(pts, 0.05)
    assert np.array_equal(v.to_numpy(), o_v)
    s = v.statistical_outlier_removal(10, 1.0)
    o_keep, _, _ = oracle.sor(o_v, 10, 1.0, threads=8)
    o_s = o_v[o_keep.astype(bool)]
    assert np.array_equal(s.to_numpy(), o_s)
    nrm = s.estimate_normals(20)
    assert np.array_equal(nrm.to_numpy(), o_s)
    host_n = pcr.normals_array(pcr.PointCloud.from_numpy(o_s), 20)
    assert np.array_equal(nrm.normals_to_numpy(), host_n)                # same kernels, same bits
    o_n = oracle.normals(o_s, 20, threads=8)
    a, b = nrm.normals_to_numpy().astype(np.float64), o_n.astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), np.abs(np.sum(a * b, axis=1)))
    assert ang.max() < 1e-4
    fused = v.sor_normals(10, 1.0, 20)                                   # one index for both steps: same bits
    assert np.array_equal(fused.to_numpy(), o_s) and np.array_equal(fused.normals_to_numpy(), host_n)
    got = s.euclidean_cluster(0.5, 30, 25000)
    want = [list(map(int, c)) for c in oracle.euclidean_cluster(o_s, 0.5, 30, 25000)]
    assert got == want
    r = s.radius_outlier_removal(0.5, 5)
    keep, _ = pcr.ror_mask(pcr.PointCloud.from_numpy(o_s), 0.5, 5)
    assert np.array_equal(r.to_numpy(), o_s[keep.astype(bool)])


def test_filters_carry_normals_and_edge_cases(pcr):
    pts = scenes.uniform_cube(3000, 2, 0, 5)
    d = pcr.DeviceCloud.from_numpy(pts).estimate_normals(8)
    s = d.statistical_outlier_removal(6, 0.5)                            # select keeps the normals (cloud.rs:117-128)
    keep, _, _, _ = pcr.sor_mask(pcr.PointCloud.from_numpy(pts), 6, 0.5)
    assert s.has_normals() and np.array_equal(s.normals_to_numpy(), d.normals_to_numpy()[keep.astype(bool)])
    assert len(d.statistical_outlier_removal(0, 1.0)) == 0               # statistical_outlier.rs:5-7
    one = pcr.DeviceCloud.from_numpy(np.array([[1, 2, 3]], np.float32))
    assert np.array_equal(one.statistical_outlier_removal(5, 1.0).to_numpy(), [[1, 2, 3]])   # :10-12
    # one point: zero covariance -> (0,0,1) (estimate.rs:174-177), flipped towards the origin (:99-107)
    assert np.array_equal(one.estimate_normals(5).normals_to_numpy(), pcr.normals_array(pcr.PointCloud.from_numpy(np.array([[1, 2, 3]], np.float32)), 5))
    assert one.estimate_normals(5).normals_to_numpy().tolist() == [[-0.0, -0.0, -1.0]]
    with pytest.raises(ValueError):                                       # k == 0: reported (the reference attaches empty normals)
        one.estimate_normals(0)
    with pytest.raises(ValueError):
        d.voxel_downsample(0.0)


def test_device_icp_matches_host_api(pcr, oracle):
    tgt = scenes.hemisphere(6000, seed=3, radius=5.0)
    src = oracle.apply_transform(tgt, scenes.rot_z(0.04), [0.2, -0.1, 0.05])
    d_t = pcr.DeviceCloud.from_numpy(tgt).estimate_normals(15)
    d_s = pcr.DeviceCloud.from_numpy(src)
    h_t = pcr.estimate_normals(pcr.PointCloud.from_numpy(tgt), 15)
    h_s = pcr.PointCloud.from_numpy(src)
    for dev, host in ((d_s.icp_point_to_plane(d_t, 15, 0.0), pcr.icp_point_to_plane(h_s, h_t, 15, 0.0)),
                      (d_s.icp_point_to_point(d_t, 15, 0.0), pcr.icp_point_to_point(h_s, h_t, 15, 0.0))):
        assert dev.num_iterations == host.num_iterations == 15
        assert np.array_equal(np.array(dev.rotation), np.array(host.rotation))
        assert np.array_equal(np.array(dev.translation), np.array(host.translation))
    with pytest.raises(ValueError):                                      # registration.rs:80-86
        d_s.icp_point_to_plane(pcr.DeviceCloud.from_numpy(tgt))
    R, t = scenes.rot_z(0.3), [1.0, 2.0, 3.0]
    moved = d_s.apply_transform(R, t)
    assert np.array_equal(moved.to_numpy(), oracle.apply_transform(src, R, t))


def test_frames_in_flight_on_independent_contexts(pcr):
    """Contexts share nothing but the device: host threads, each with its own context and stream, run whole
    pipelines concurrently and every one gets the bits the single-threaded run gets."""
    import threading

    frames = [scenes.kitti_scene(20 + i, (12_000, 600, 100, 260)) for i in range(3)]
    want = []
    for f in frames:
        o = pcr.DeviceCloud.from_numpy(f).voxel_downsample(0.05).sor_normals(10, 1.0, 20)
        want.append((o.to_numpy(), o.normals_to_numpy(), o.euclidean_cluster(0.5, 30, 25000)))
    errors = []

    def worker(t):
        try:
            ctx = pcr.Context(device=0)
            ctx.set_frame_stream(t % 2 == 0)
            for rep in range(6):
                i = (t + rep) % len(frames)
                o = pcr.DeviceCloud.from_numpy(frames[i], ctx).voxel_downsample(0.05).sor_normals(10, 1.0, 20)
                assert np.array_equal(o.to_numpy(), want[i][0]) and np.array_equal(o.normals_to_numpy(), want[i][1])
                assert o.euclidean_cluster(0.5, 30, 25000) == want[i][2]
            ctx.close()
        except Exception as ex:  # surfaced in the main thread
            errors.append(repr(ex))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_frame_stream_deferred_counts_not_awaited_and_redone(pcr):
    """Frame stream, fused SOR -> normals: after a frame whose first grid level left nothing over, the next frame's
    leftover counts are not awaited (pcr_ctx_set_frame_stream).  Frames that DO leave queries over --
    a handful of points hundreds of metres away need the third grid level -- must then be redone the careful way:
    same bits as a context without the hint, and the miss is counted."""
    base = scenes.kitti_scene(11, (20_000, 1000, 170, 430))
    far = np.array([[900.0, 0, 0], [0, -850.0, 5], [910.0, 3.0, 1], [-700, 600, 40], [905.0, -2.0, 0.5]], np.float32)
    frames = [base, scenes.kitti_scene(12, (20_000, 1000, 170, 430)), np.vstack([scenes.kitti_scene(13, (20_000, 1000, 170, 430)), far]),
              scenes.kitti_scene(14, (20_000, 1000, 170, 430)), scenes.kitti_scene(15, (20_000, 1000, 170, 430))]
    plain = pcr.Context(device=0)
    stream = pcr.Context(device=0)
    stream.set_frame_stream(True)
    for pts in frames:
        want = pcr.DeviceCloud.from_numpy(pts, plain).sor_normals(10, 1.0, 20)
        got = pcr.DeviceCloud.from_numpy(pts, stream).sor_normals(10, 1.0, 20)
        assert np.array_equal(got.to_numpy(), want.to_numpy())
        assert np.array_equal(got.normals_to_numpy(), want.normals_to_numpy())
    h = stream.hint_stats()
    assert h["deferred_counts_not_awaited_zero"] >= 1           # frames 2 and 5 ran without the round trip
    assert h["deferred_counts_not_awaited_nonzero"] >= 1        # frame 3 was redone
    plain.close()
    stream.close()
