"""Locate where GPU and oracle split at a given step of the config-3 stream (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hector_slam_b200 import capi, synth
from oracle import pyoracle

STEP = int(sys.argv[1]) if len(sys.argv) > 1 else 12
size = 4096
world = synth.World.for_map_size(size)
orc = pyoracle.Oracle("port", 0.05, size, 3)
orc.set_update_factors(0.4, 0.9); orc.set_map_update_thresholds(0.4, 0.9)
pose = np.array([3.0, 2.0, 0.1]); rng = np.random.default_rng(5)
hint = pose.astype(np.float32)
for k in range(STEP + 1):
    scan = np.ascontiguousarray(synth.make_scan(world, pose, rng))
    if k == STEP:
        break
    hint, _ = orc.update(scan, hint)
    h = pose[2]; pose = pose + np.array([0.0125*np.cos(h), 0.0125*np.sin(h), 0.0075])
planes = [orc.get_logodds(l) for l in range(3)]
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "stream_step.npz"), scan=scan, hint=hint, p0=planes[0][1800:2400,1800:2400])
cur = hint.copy()
for lvl in (2, 1, 0):
    sz = size >> lvl
    pl = (scan * np.float32(2.0 ** -lvl)).astype(np.float32)
    for mode in (1, 2):
        res = []
        for mi in range(0, 6 if lvl == 0 else 4):
            rep = capi.MapRepB200(0.05 * 2 ** lvl, sz, levels=1, max_iterations=[mi if mi > 0 else -1], gather_mode=mode)
            rep.upload_level(0, planes[lvl])
            g, _ = rep.matchData(cur, pl)
            o, _ = orc.match_level(lvl, cur, pl, mi)
            res.append(np.abs(g - o).max())
            rep.close()
        print(f"level {lvl} mode {mode}: |gpu-oracle| after 1..N evaluations:", " ".join(f"{x:.1e}" for x in res))
    cur, _ = orc.match_level(lvl, cur, pl, 5 if lvl == 0 else 3)
print("final oracle", cur)

# --- chains: which hand-off makes the full pipelines split? ------------------------------------
def gpu_level(lvl, start, mi):
    sz = size >> lvl
    pl = (scan * np.float32(2.0 ** -lvl)).astype(np.float32)
    rep = capi.MapRepB200(0.05 * 2 ** lvl, sz, levels=1, max_iterations=[mi], gather_mode=2)
    rep.upload_level(0, planes[lvl])
    g, _ = rep.matchData(start, pl)
    rep.close()
    return g
def orc_level(lvl, start, mi):
    pl = (scan * np.float32(2.0 ** -lvl)).astype(np.float32)
    return orc.match_level(lvl, start, pl, mi)[0]
rep3 = capi.MapRepB200(0.05, size, levels=3, gather_mode=2)
for l in range(3):
    rep3.upload_level(l, planes[l])
full_gpu, _ = rep3.matchData(hint, scan)
full_orc, _ = orc.match(hint, scan)
g2 = gpu_level(2, hint, 3); o2 = orc_level(2, hint, 3)
g1 = gpu_level(1, g2, 3); o1_from_g2 = orc_level(1, g2, 3); o1 = orc_level(1, o2, 3)
g0 = gpu_level(0, g1, 5); o0_from_g1 = orc_level(0, g1, 5); o0 = orc_level(0, o1, 5)
np.set_printoptions(precision=9)
print("hint       ", hint)
print("L2 gpu/orc ", g2, o2, np.abs(g2 - o2).max())
print("L1 gpu(g2) / orc(g2) / orc(o2)", g1, o1_from_g2, o1)
print("L0 gpu(g1) / orc(g1) / orc(o1)", g0, o0_from_g1, o0)
print("full gpu   ", full_gpu, " full orc", full_orc, " diff", np.abs(full_gpu - full_orc))
print("chain gpu vs full gpu", np.abs(g0 - full_gpu))
