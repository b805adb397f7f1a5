"""Single-scan latency probe (BASELINE.json configs[0]-style use: one scan at a time, match then map update).

GPU: hsb_match_data / hsb_update_by_scan through the C-ABI with pageable host buffers, wall clock per call.
CPU: the compiled reference (oracle/_ref) on one thread — how the reference runs it.  Diagnostic, not a bench line.
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench
from hector_slam_b200 import capi

N = 300
rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
world, poses, pts, offs, hints = bench.make_workload(0, 64)
bench.build_map_on_gpu(rep, world)
scans = [np.ascontiguousarray(pts[offs[i]:offs[i + 1]]) for i in range(64)]

def wall(fn, n=N):
    for i in range(10): fn(i)
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    return (time.perf_counter() - t0) / n * 1e6

res = {}
out = [None]
def g_match(i):
    out[0] = rep.matchData(hints[i % 64], scans[i % 64])
res["gpu_match_us"] = wall(g_match)
def g_update(i):
    rep.updateByScan(scans[i % 64], poses[i % 64])
res["gpu_update_us"] = wall(g_update)
def g_step(i):
    p, _ = rep.matchData(hints[i % 64], scans[i % 64]); rep.updateByScan(scans[i % 64], p); rep.onMapUpdated()
res["gpu_step_us"] = wall(g_step)
if hasattr(rep, "slam_update"):
    rep.setMapUpdateMinDistDiff(0.0); rep.setMapUpdateMinAngleDiff(0.0)   # every step writes the map
    def g_fused(i):
        rep.slam_update(hints[i % 64], scans[i % 64])
    res["gpu_fused_step_us"] = wall(g_fused)

# the same calls without the numpy wrappers (ctypes only): what a C caller sees
lib, hnd = rep.lib, rep.h
hp = [np.ascontiguousarray(hints[i], np.float32) for i in range(64)]
pp = [np.ascontiguousarray(poses[i], np.float32) for i in range(64)]
o_pose, o_cov = np.zeros(3, np.float32), np.zeros(9, np.float32)
import ctypes as C
upd = C.c_int(0)
def r_match(i):
    k = i % 64
    lib.hsb_match_data(hnd, hp[k].ctypes.data, scans[k].ctypes.data, scans[k].shape[0], None, o_pose.ctypes.data, o_cov.ctypes.data)
res["raw_match_us"] = wall(r_match)
def r_update(i):
    k = i % 64
    lib.hsb_update_by_scan(hnd, scans[k].ctypes.data, scans[k].shape[0], None, pp[k].ctypes.data)
res["raw_update_us"] = wall(r_update)
def r_fused(i):
    k = i % 64
    lib.hsb_slam_update(hnd, hp[k].ctypes.data, scans[k].ctypes.data, scans[k].shape[0], None, 0, o_pose.ctypes.data, o_cov.ctypes.data, C.addressof(upd))
res["raw_fused_us"] = wall(r_fused)

rep.set_tuning(host_out=0)   # results through a device-to-host copy operation instead of mapped host memory
res["raw_match_d2h_us"] = wall(r_match)
res["raw_fused_d2h_us"] = wall(r_fused)
rep.set_tuning(host_out=1)
if "--no-cpu" in sys.argv:
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}); sys.exit(0)
try:
    import pyoracle
    if pyoracle.available("ref"):
        orc = pyoracle.Oracle("ref", bench.RES, bench.MAP_SIZE, 3)
        orc.set_update_factors(0.4, 0.9)
        for l in range(3):
            orc.set_logodds(l, rep.download_level(l))
        def c_match(i):
            orc.match(hints[i % 64], scans[i % 64])
        res["cpu_ref_match_us"] = wall(c_match, 100)
        def c_update(i):
            orc.update_by_scan(scans[i % 64], poses[i % 64])
        res["cpu_ref_update_us"] = wall(c_update, 100)
except Exception as e:  # diagnostic only
    res["cpu_error"] = repr(e)
print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()})
