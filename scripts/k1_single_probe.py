"""K1 on ONE scan (the 40 Hz use, config 3 map): where do the ~28 us go?  Per-level stamps (%globaltimer, tuning key
trace) for the launch shapes a single scan can take, plus host latency of hsb_match_data.

  python scripts/k1_single_probe.py > gpurun_out/k1_single_probe.log
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hector_slam_b200 import capi, synth  # noqa: E402

size = 4096
rep = capi.MapRepB200(bench.RES, size, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
world = synth.World.for_map_size(size)
bench.build_map_on_gpu(rep, world)
rng = np.random.default_rng(31)
poses = world.sample_free_poses(32, rng)
scans = [np.ascontiguousarray(synth.make_scan(world, p, rng)) for p in poses]
hints = synth.perturb_hints(poses, seed=32, dxy=0.05, dpsi=0.02)
ref = None
for W, U in ((8, 4), (8, 5), (4, 4), (16, 3), (2, 8)):
    try:
        rep.set_tuning(warps_per_scan=W, scans_per_block=1, unroll=U, trace=1)
        lv, tot, lat = [], [], []
        for i in range(120):
            k = i % 32
            t0 = time.perf_counter()
            p, _ = rep.matchData(hints[k], scans[k])
            lat.append(time.perf_counter() - t0)
            tr = rep.read_trace(1).astype(np.int64)[0]
            lv.append(np.diff(tr[0:5]) / 1e3)
            tot.append((tr[4] - tr[0]) / 1e3)
            if k == 0:
                if ref is None:
                    ref = p.copy()
                same = float(np.abs(p - ref).max())
        lv = np.median(np.asarray(lv)[20:], axis=0)
        print(f"W={W:2d} U={U}: kernel body {np.median(tot[20:]):5.1f} us  per level (coarse -> fine) "
              f"{lv[0]:5.1f} {lv[1]:5.1f} {lv[2]:5.1f}  tail {lv[3]:4.1f}   hsb_match_data p50 {np.median(lat[20:]) * 1e6:5.1f} us   "
              f"max diff to W=8,U=4 on scan 0: {same:.1e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"W={W} U={U}: {e}", flush=True)
rep.set_tuning(warps_per_scan=0, scans_per_block=0, unroll=0, trace=0)
# what the host->device copy of the scan costs the call: the same launch on device-resident inputs, synchronised
import torch  # noqa: E402

dev = torch.device("cuda", 0)
d_pts = [torch.from_numpy(s).to(dev) for s in scans]
d_hints = [torch.from_numpy(np.ascontiguousarray(h)).to(dev) for h in hints]
d_pose = torch.empty(3, dtype=torch.float32, device=dev)
d_cov = torch.empty(9, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
lat = []
for i in range(220):
    k = i % 32
    n = scans[k].shape[0]
    t0 = time.perf_counter()
    rep.match_batch_device(1, d_hints[k].data_ptr(), d_pts[k].data_ptr(), None, n, n, d_pose.data_ptr(), d_cov.data_ptr(), st)
    torch.cuda.current_stream().synchronize()
    lat.append(time.perf_counter() - t0)
print(f"hsb_match_batch_device(B=1) + stream synchronize, inputs resident: p50 {np.median(lat[20:]) * 1e6:5.1f} us "
      f"(hsb_match_data with the copy of the scan: see above)", flush=True)
lat = []
for i in range(220):
    t0 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    lat.append(time.perf_counter() - t0)
print(f"an empty stream synchronize from Python: p50 {np.median(lat[20:]) * 1e6:5.1f} us", flush=True)
rep.close()
