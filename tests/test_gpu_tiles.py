"""GPU: dirty-rectangle replication of map writes (one owner GPU writes the map,
replicas receive tiles over NCCL).  Single-GPU part: the rectangle covers every written cell and
pack -> unpack reproduces planes and probabilities on a second handle.  Two-GPU part (skipped with
one GPU): rank 0 runs SLAM and broadcasts tiles after every map write; rank 1's replica must end
bit-identical and must match scans identically."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def test_dirty_rect_pack_unpack(hsb_lib):
    import torch

    from hector_slam_b200 import capi

    g = load_golden("slam3.npz")
    a = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    b = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    assert a.get_dirty_rect(0) is None
    pose = g["first_hint"]
    for k in range(3):
        before = [a.download_level(l) for l in range(3)]
        p, _ = a.matchData(pose, g["scans"][k])
        a.updateByScan(g["scans"][k], p)
        for l in range(3):
            rect = a.get_dirty_rect(l, reset=True)
            assert rect is not None and a.get_dirty_rect(l) is None
            x0, y0, x1, y1 = rect
            after = a.download_level(l)
            changed = np.argwhere(after != before[l])
            assert changed[:, 1].min() >= x0 and changed[:, 1].max() <= x1
            assert changed[:, 0].min() >= y0 and changed[:, 0].max() <= y1
            n = (x1 - x0 + 1) * (y1 - y0 + 1)
            assert n < after.size // 2                      # a tile, not the whole plane
            buf = torch.empty(n, dtype=torch.float32, device="cuda")
            # no synchronisation around pack / unpack: the C-ABI orders the caller's stream against the handles' own
            a.pack_rect_device(l, rect, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
            b.unpack_rect_device(l, rect, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert np.array_equal(b.download_level_rect(l, rect), after[y0:y1 + 1, x0:x1 + 1])
            assert np.array_equal(buf.cpu().numpy().reshape(y1 - y0 + 1, x1 - x0 + 1), after[y0:y1 + 1, x0:x1 + 1])
        pose = p
    for l in range(3):
        assert np.array_equal(a.download_level(l), b.download_level(l))
        assert np.array_equal(a.download_prob(l), b.download_prob(l))
    pa, _ = a.matchData(pose, g["scans"][3])
    pb, _ = b.matchData(pose, g["scans"][3])
    assert np.array_equal(pa, pb)
    a.close()
    b.close()


def test_one_shot_tile_transport(hsb_lib):
    """hsb_pack_dirty_device / hsb_unpack_dirty_device: the self-describing buffer reproduces the owner's planes on a
    second handle without any size passing through the host; an undersized buffer ships nothing, keeps the rectangles
    and is counted on the receiving side."""
    import torch

    from hector_slam_b200 import capi

    g = load_golden("slam3.npz")
    a = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    b = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    cap = 4 << 20
    buf = torch.zeros(cap // 4, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    pose = g["first_hint"]
    for k in range(4):
        pose, _ = a.matchData(pose, g["scans"][k])
        a.updateByScan(g["scans"][k], pose)
        if k == 1:
            continue                                          # two writes accumulate in one rectangle
        rects = a.get_dirty_rects()
        a.pack_dirty_device(buf.data_ptr(), cap, True, st)
        b.unpack_dirty_device(buf.data_ptr(), cap, st)       # no synchronisation in between
        hdr = buf[:64].view(torch.int32).cpu().numpy()
        assert hdr[1] == 3 and hdr[2] == 0 and hdr[3] == sum((r[2] - r[0] + 1) * (r[3] - r[1] + 1) for r in rects)
        assert a.get_dirty_rects() == [None, None, None]
        for l in range(3):
            assert np.array_equal(a.download_level(l), b.download_level(l))
            assert np.array_equal(a.download_prob(l), b.download_prob(l))
            assert b.get_mirror_dirty_rect(l) is not None
    assert b.replication_overflows() == 0
    pa, _ = a.matchData(pose, g["scans"][4])
    pb, _ = b.matchData(pose, g["scans"][4])
    assert np.array_equal(pa, pb)
    # undersized buffer: nothing shipped, rectangles kept, replica counts the event
    a.updateByScan(g["scans"][4], pa)
    small = 64 * 4 + 1000 * 4
    before = a.get_dirty_rects()
    a.pack_dirty_device(buf.data_ptr(), small, True, st)
    b.unpack_dirty_device(buf.data_ptr(), small, st)
    assert a.get_dirty_rects() == before and b.replication_overflows(reset=True) == 1 and b.replication_overflows() == 0
    a.pack_dirty_device(buf.data_ptr(), cap, True, st)       # the full-size buffer then carries it
    b.unpack_dirty_device(buf.data_ptr(), cap, st)
    for l in range(3):
        assert np.array_equal(a.download_level(l), b.download_level(l))
    a.close()
    b.close()


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from hector_slam_b200 import capi, parallel

        g = np.load(os.path.join(ROOT, "tests", "golden", "slam3.npz"))
        rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, device=rank, update_factor_free=0.4,
                              update_factor_occupied=0.9)
        pose = g["first_hint"]
        shipped = 0
        probes = []
        tbuf = parallel.tile_buffer(dev)
        for k in range(g["scans"].shape[0]):
            if rank == 0:                                  # the owner runs SLAM and writes the map
                pose, _ = rep.matchData(pose, g["scans"][k])
                rep.updateByScan(g["scans"][k], pose)
                rep.onMapUpdated()
            if k % 2:                                      # alternate the two protocols: two-step (sizes via the hosts) ...
                shipped += parallel.broadcast_dirty_tiles(rep, dev, src=0)
            else:                                          # ... and one-shot (self-describing buffer, no host in the loop)
                parallel.broadcast_dirty_tiles_async(rep, tbuf, src=0)
                shipped += 1
            # straight after the broadcast, no synchronize: the replica's match must already see the new tiles
            pr, _ = rep.match_batch(g["est"][k:k + 1], g["scans"][k], None)
            probes.append(pr[0])
        pt = torch.from_numpy(np.asarray(probes, np.float32)).to(dev)
        pg = [torch.empty_like(pt) for _ in range(world)]
        dist.all_gather(pg, pt)
        same_probe = all(bool(torch.equal(pg[0], x)) for x in pg)
        torch.cuda.synchronize()
        planes = [rep.download_level(l) for l in range(3)]
        sums = torch.tensor([float(np.abs(p).sum(dtype=np.float64)) for p in planes], dtype=torch.float64, device=dev)
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        same_map = all(bool(torch.equal(allsums[0], s)) for s in allsums)
        K = g["scans"].shape[0]
        pts = g["scans"].reshape(-1, 2)
        offs = (np.arange(K + 1) * g["scans"].shape[1]).astype(np.int32)
        got, _ = rep.match_batch(g["est"], pts, offs)
        t = torch.from_numpy(got).to(dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        same_match = all(bool(torch.equal(gathered[0], x)) for x in gathered)
        q.put((rank, same_map, same_match and same_probe and rep.replication_overflows() == 0, shipped, float(sums.sum().item())))
        rep.close()
    finally:
        dist.destroy_process_group()


def test_two_gpus_tile_broadcast(hsb_lib):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == res[1][3] > 0 and res[0][4] > 0
