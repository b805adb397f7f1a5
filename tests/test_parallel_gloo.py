"""CPU, world_size 2 on gloo: the multi-process plumbing of hector_slam_b200/parallel.py —
contiguous scan sharding, map replication by broadcast, result assembly — with the CPU oracle
standing in for the per-rank matcher (the sharding logic is independent of who computes)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden_planes, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hector_slam_b200 import parallel
        from oracle import pyoracle

        g = np.load(os.path.join(ROOT, "tests", "golden", "match3.npz"))
        size, levels = int(g["size"]), int(g["levels"])
        # rank 0 owns the map; the others start empty and receive it by broadcast
        planes = []
        for l in range(levels):
            p = np.zeros(((size >> l), (size >> l)), np.float32)
            if rank == 0:
                flat = p.reshape(-1)
                flat[g[f"map_idx{l}"]] = g[f"map_val{l}"]
            planes.append(torch.from_numpy(p))
        parallel.broadcast_planes(planes, src=0)
        orc = pyoracle.Oracle("port", float(g["res"]), size, levels)
        orc.set_update_factors(0.4, 0.9)
        for l in range(levels):
            orc.set_logodds(l, planes[l].numpy())

        K = g["scans"].shape[0]
        # ragged batch: drop a different number of beams from every scan, one scan empty
        chunks, offs = [], [0]
        for k in range(K):
            s = g["scans"][k][: (0 if k == 5 else 1081 - 13 * k)]
            chunks.append(s)
            offs.append(offs[-1] + s.shape[0])
        pts = np.concatenate(chunks).astype(np.float32)
        offs = np.asarray(offs, np.int32)
        hints = g["hints"]

        def match_fn(h, p, o):
            P, C, _ = orc.match_batch(h, p, o, nthreads=1)
            return P, C

        poses, cov = parallel.match_sharded(match_fn, hints, pts, offs)
        # every rank must hold the full, correctly ordered result == the unsharded run
        want, want_cov, _ = orc.match_batch(hints, pts, offs, nthreads=1)
        ok = bool(np.array_equal(poses.numpy(), want)) and bool(np.array_equal(cov.numpy(), want_cov))
        # dirty-tile protocol with a host stand-in for the handle (same method names as capi.MapRepB200)
        import ctypes

        MAGIC, HDR = 0x48534254, 64   # one-shot buffer: header words of hsb_pack_dirty_device

        class FakeRep:
            def __init__(self):
                self.planes = [np.zeros((32 >> l, 32 >> l), np.float32) for l in range(2)]
                self.dirty = [None, None]

            def getMapLevels(self):
                return 2

            def get_dirty_rects(self, reset=False):
                return [self.get_dirty_rect(l, reset) for l in range(2)]

            def write(self, l, x0, y0, x1, y1, v):
                self.planes[l][y0:y1 + 1, x0:x1 + 1] = v
                self.dirty[l] = (x0, y0, x1, y1)

            def get_dirty_rect(self, l, reset=False):
                r = self.dirty[l]
                if reset:
                    self.dirty[l] = None
                return r

            def _view(self, ptr, n):
                return np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr))

            def pack_rect_device(self, l, rect, ptr, stream=0):
                x0, y0, x1, y1 = rect
                self._view(ptr, (x1 - x0 + 1) * (y1 - y0 + 1))[:] = self.planes[l][y0:y1 + 1, x0:x1 + 1].reshape(-1)

            def unpack_rect_device(self, l, rect, ptr, stream=0):
                x0, y0, x1, y1 = rect
                n = (x1 - x0 + 1) * (y1 - y0 + 1)
                self.planes[l][y0:y1 + 1, x0:x1 + 1] = self._view(ptr, n).reshape(y1 - y0 + 1, x1 - x0 + 1)

            # one-shot protocol: the self-describing buffer of hsb_pack_dirty_device / hsb_unpack_dirty_device
            # (header of 64 words: magic, levels, overflow flag, cells, then x0 y0 x1 y1 per level; rows back to back)
            def pack_dirty_device(self, ptr, nbytes, reset, stream=0):
                words = self._view(ptr, nbytes // 4)
                hdr = words[:HDR].view(np.int32)
                rects = [self.dirty[l] if self.dirty[l] is not None else (2 ** 31 - 1, 2 ** 31 - 1, -1, -1) for l in range(2)]
                cells = sum((r[2] - r[0] + 1) * (r[3] - r[1] + 1) for r in rects if r[2] >= r[0])
                overflow = cells + HDR > nbytes // 4
                hdr[:4] = [MAGIC, 2, int(overflow), cells]
                hdr[4:12] = np.int32(rects).reshape(-1)
                if overflow:
                    return
                o = HDR
                for l, (x0, y0, x1, y1) in enumerate(rects):
                    if x1 >= x0:
                        n = (x1 - x0 + 1) * (y1 - y0 + 1)
                        words[o:o + n] = self.planes[l][y0:y1 + 1, x0:x1 + 1].reshape(-1)
                        o += n
                if reset:
                    self.dirty = [None, None]

            def unpack_dirty_device(self, ptr, nbytes, stream=0):
                words = self._view(ptr, nbytes // 4)
                hdr = words[:HDR].view(np.int32)
                if hdr[0] != MAGIC or hdr[1] != 2 or hdr[2] != 0:
                    self.overflows = getattr(self, "overflows", 0) + 1
                    return
                o = HDR
                for l in range(2):
                    x0, y0, x1, y1 = (int(v) for v in hdr[4 + 4 * l: 8 + 4 * l])
                    if x1 >= x0:
                        n = (x1 - x0 + 1) * (y1 - y0 + 1)
                        self.planes[l][y0:y1 + 1, x0:x1 + 1] = words[o:o + n].reshape(y1 - y0 + 1, x1 - x0 + 1)
                        o += n

        fr = FakeRep()
        if rank == 0:
            fr.write(0, 3, 4, 10, 9, 1.5)      # level 1 stays clean on purpose
        st = {}
        shipped = parallel.broadcast_dirty_tiles(fr, "cpu", src=0, stats=st)
        ok = ok and shipped == 8 * 6 and float(fr.planes[0].sum()) == 1.5 * 48 and float(fr.planes[1].sum()) == 0.0
        ok = ok and fr.get_dirty_rect(0) is None and st["cells"] == 48
        # both levels dirty: ONE packed buffer carries both rectangles back to back
        if rank == 0:
            fr.write(0, 0, 0, 4, 1, 2.0)
            fr.write(1, 2, 2, 3, 5, -1.0)
        shipped = parallel.broadcast_dirty_tiles(fr, "cpu", src=0)
        ok = ok and shipped == 10 + 8 and float(fr.planes[1].sum()) == -8.0 and float(fr.planes[0][0:2, 0:5].sum()) == 20.0
        ok = ok and parallel.broadcast_dirty_tiles(fr, "cpu", src=0) == 0   # nothing dirty: nothing shipped
        # one-shot protocol over the same transport: one fixed-size broadcast, no sizes through the hosts
        buf = torch.zeros(64 + 200, dtype=torch.float32)
        if rank == 0:
            fr.write(0, 5, 5, 9, 7, 3.0)
            fr.write(1, 0, 1, 1, 2, -2.0)
        parallel.broadcast_dirty_tiles_async(fr, buf, src=0)
        ok = ok and float(fr.planes[0][5:8, 5:10].sum()) == 45.0 and float(fr.planes[1][1:3, 0:2].sum()) == -8.0
        ok = ok and (rank != 0 or fr.dirty == [None, None])
        parallel.broadcast_dirty_tiles_async(fr, buf, src=0)               # nothing dirty: header only, planes unchanged
        ok = ok and float(fr.planes[0][5:8, 5:10].sum()) == 45.0
        # a dirty area larger than the buffer: nothing is shipped, the owner keeps its rectangles, replicas count it
        before = fr.planes[0].copy()
        if rank == 0:
            fr.write(0, 0, 0, 31, 31, 7.0)                                  # 1024 cells > 200
        parallel.broadcast_dirty_tiles_async(fr, buf, src=0)
        if rank == 0:
            ok = ok and fr.dirty[0] == (0, 0, 31, 31)
        else:
            ok = ok and getattr(fr, "overflows", 0) == 1 and bool(np.array_equal(fr.planes[0], before))
        shipped = parallel.broadcast_dirty_tiles(fr, "cpu", src=0)          # ... and the two-step protocol catches up
        ok = ok and shipped == 1024 and float(fr.planes[0].sum()) == 7.0 * 1024
        lo, hi = parallel.shard_range(K, rank, world)
        q.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from hector_slam_b200 import parallel

    for n in (0, 1, 7, 16, 4096, 65536 + 3):
        for world in (1, 2, 3, 8):
            edges = [parallel.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo(pyoracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results), results
    assert sorted(r[2] for r in results) == [(0, 8), (8, 16)]


def test_cpulist_parsing_and_numa_binding_without_gpu():
    """Host plumbing of the e2e path: sysfs cpulist parsing; binding degrades to a no-op (None) where there is no
    CUDA device or no sysfs entry, and never raises."""
    from hector_slam_b200 import parallel

    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist("") == set()
    assert parallel._parse_cpulist("5") == {5}
    import torch

    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert parallel.bind_process_to_gpu_numa_node(0) is None
        assert os.sched_getaffinity(0) == before
