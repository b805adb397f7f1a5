"""Parity statistics of the CUDA matcher against the CPU oracle (diagnostic, not a test).
Usage: [HSB_LIB_PATH=...] python scripts/parity_probe.py [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hector_slam_b200 import capi, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
print("lib:", capi.LIB_PATH)


def stats(name, got, want):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    d[:, 2] = np.abs((d[:, 2] + np.pi) % (2 * np.pi) - np.pi)
    m = d.max(axis=1)
    q = np.quantile(m, [0.5, 0.9, 0.99, 1.0])
    print(f"{name}: n={len(m)} median={q[0]:.2e} p90={q[1]:.2e} p99={q[2]:.2e} max={q[3]:.2e}  >1e-4: {(m > 1e-4).sum()}")


# 1) SLAM golden sequence
g = np.load(os.path.join(ROOT, "tests", "golden", "slam3.npz"))
for mode in (1, 2):
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9,
                          gather_mode=mode)
    hint = g["first_hint"]
    got = []
    for k in range(g["scans"].shape[0]):
        pose, _ = rep.matchData(hint, g["scans"][k])
        got.append(pose)
        rep.updateByScan(g["scans"][k], pose)
        rep.onMapUpdated()
        hint = pose
    got = np.asarray(got)
    d = np.abs(got - g["est"])
    print("slam mode", mode, "per-step max err:", " ".join(f"{x:.1e}" for x in d.max(axis=1)))
    rep.close()

# 1b) same but each step matched against the ORACLE's map state (isolates the matcher from map drift)
orc = pyoracle.Oracle("port", float(g["res"]), int(g["size"]), 3)
orc.set_update_factors(0.4, 0.9)
orc.set_map_update_thresholds(0.0, 0.0)
rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
hint = g["first_hint"]
errs = []
for k in range(g["scans"].shape[0]):
    for l in range(3):
        rep.upload_level(l, orc.get_logodds(l))
    pose_g, _ = rep.matchData(hint, g["scans"][k])
    pose_o, _ = orc.update(g["scans"][k], hint)
    errs.append(np.abs(pose_g - pose_o).max())
    hint = pose_o
print("slam, oracle map each step:", " ".join(f"{x:.1e}" for x in errs))
rep.close()
orc.close()

# 2) batch statistics on the 2048^2 3-level map and a 1024^2 1-level map
for size, levels, dxy, dpsi in ((2048, 3, 0.1, 0.05), (1024, 1, 0.03, 0.015), (1024, 1, 0.05, 0.025)):
    world = synth.World.for_map_size(size)
    orc = pyoracle.Oracle("port", 0.05, size, levels)
    orc.set_update_factors(0.4, 0.9)
    pyoracle.build_map_known_poses(orc, world)
    rng = np.random.default_rng(42)
    poses = world.sample_free_poses(B, rng)
    pts, offs = synth.make_scan_batch(world, poses, noise_seed=7)
    hints = synth.perturb_hints(poses, seed=1, dxy=dxy, dpsi=dpsi)
    want, _, secs = orc.match_batch(hints, pts, offs, nthreads=min(16, os.cpu_count()))
    ok = np.abs(want[:, :2] - hints[:, :2]).max(axis=1) < 0.5
    for mode in (1, 2):
        rep = capi.MapRepB200(0.05, size, levels=levels, update_factor_free=0.4, update_factor_occupied=0.9, gather_mode=mode)
        for l in range(levels):
            rep.upload_level(l, orc.get_logodds(l))
        got, _ = rep.match_batch(hints, pts, offs)
        stats(f"{size}^2 x{levels} hints +-{dxy} mode {mode} (oracle diverged on {(~ok).sum()})", got[ok], want[ok])
        rep.close()
    orc.close()
