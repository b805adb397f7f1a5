"""Diagnostic: config-3 stream, per-step GPU-vs-oracle difference with (a) the GPU running its own hint chain and
(b) both sides given the oracle's previous pose as the hint.  Usage: python scripts/stream_parity_probe.py [key=val tuning]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle
from hector_slam_b200 import capi, synth

size = 4096
world = synth.World.for_map_size(size)
tune = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
reps = [capi.MapRepB200(0.05, size, levels=3, update_factor_free=0.4, update_factor_occupied=0.9) for _ in range(2)]
for r in reps:
    r.setMapUpdateMinDistDiff(0.4); r.setMapUpdateMinAngleDiff(0.9)
    if tune: r.set_tuning(**tune)
orc = pyoracle.Oracle("port", 0.05, size, 3)
orc.set_update_factors(0.4, 0.9); orc.set_map_update_thresholds(0.4, 0.9)
pose = np.array([3.0, 2.0, 0.1]); rng = np.random.default_rng(5)
hint_o = pose.astype(np.float32); hint_own = hint_o.copy()
worst_own = worst_same = 0.0
for k in range(120):
    scan = np.ascontiguousarray(synth.make_scan(world, pose, rng))
    p_own, _, _ = reps[0].slam_update(hint_own, scan)
    p_same, _, _ = reps[1].slam_update(hint_o, scan)
    want, _ = orc.update(scan, hint_o)
    d_own = np.abs(p_own - want); d_same = np.abs(p_same - want)
    worst_own = max(worst_own, d_own.max()); worst_same = max(worst_same, d_same.max())
    if k < 12 or d_own.max() > 5e-5 or d_same.max() > 5e-5:
        print(k, "own-chain diff %.2e" % d_own.max(), " same-hint diff %.2e" % d_same.max())
    hint_own, hint_o = p_own, want
    h = pose[2]; pose = pose + np.array([0.0125 * np.cos(h), 0.0125 * np.sin(h), 0.0075])
print("worst own-chain %.2e  worst same-hint %.2e" % (worst_own, worst_same))
