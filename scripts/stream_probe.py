"""Config-3 diagnostic: per-step GPU-vs-oracle pose differences in the streaming SLAM loop."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hector_slam_b200 import capi, synth
from oracle import pyoracle

size = 4096
world = synth.World.for_map_size(size)
orc = pyoracle.Oracle("port", 0.05, size, 3)
orc.set_update_factors(0.4, 0.9); orc.set_map_update_thresholds(0.4, 0.9)
rep = capi.MapRepB200(0.05, size, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rep.set_tuning(warps_per_scan=W)
pose = np.array([3.0, 2.0, 0.1]); rng = np.random.default_rng(5)
hint = pose.astype(np.float32)
for k in range(40):
    scan = np.ascontiguousarray(synth.make_scan(world, pose, rng))
    # GPU matches against the ORACLE's current map, from the oracle's hint
    for l in range(3):
        rep.upload_level(l, orc.get_logodds(l))
    g, gc = rep.matchData(hint, scan)
    w, wc = orc.update(scan, hint)
    d = np.abs(g - w)
    # per level detail when it differs noticeably
    flag = ""
    if d.max() > 1e-5:
        flag = "  <<<"
        # one-evaluation comparison at the hint on each level
        for l in range(3):
            pm = orc.map_coords_pose(l, hint)
            pl = (scan * np.float32(2.0 ** -l)).astype(np.float32)
            Ho, do = orc.hessian_derivs(l, pm, pl)
            Hg, dg = rep.hessian_derivs(l, pm, pl)
            print(f"      level {l}: cond(H)={np.linalg.cond(Ho.astype(np.float64)):.2e} relH={np.abs(Ho-Hg).max()/np.abs(Ho).max():.1e} reld={np.abs(do-dg).max()/max(np.abs(do).max(),1e-9):.1e}")
    print(f"step {k:2d}: diff {d[0]:.1e} {d[1]:.1e} {d[2]:.1e}  covdiff {np.abs(gc-wc).max()/max(np.abs(wc).max(),1e-9):.1e}{flag}")
    hint = w
    h = pose[2]; pose = pose + np.array([0.0125*np.cos(h), 0.0125*np.sin(h), 0.0075])
