"""One fused SLAM step (single-scan match -> gate -> mark -> apply) inside a cudaProfilerStart/Stop window:
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/slam_step python scripts/ncu_slam_step.py
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hector_slam_b200 import capi

rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
world, poses, pts, offs, hints = bench.make_workload(0, 16)
bench.build_map_on_gpu(rep, world)
rep.setMapUpdateMinDistDiff(0.0); rep.setMapUpdateMinAngleDiff(0.0)
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("="); rep.set_tuning(**{k: int(v)})
scans = [np.ascontiguousarray(pts[offs[i]:offs[i + 1]]) for i in range(16)]
for i in range(8):
    rep.slam_update(hints[i], scans[i])
torch.cuda.synchronize()
torch.cuda.profiler.start()
rep.slam_update(hints[9], scans[9])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
