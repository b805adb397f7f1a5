"""GPU: the CUDA map writer (K2, two-phase Bresenham with stamp atomics) and the fused probability
refresh (K3) through the C-ABI, against golden planes from the compiled reference and against the
CPU oracle.  Planes are compared with abs tol 1e-5 (a cell freed then hit in one
scan ends as ((l+lf)-lf)+lo in the reference and l+lo here); the number of cells whose VALUE class
differs (a beam end rounding into a neighbouring cell) must be tiny and is reported."""
import numpy as np
import pytest

from conftest import apply_diff, golden_planes, load_golden, pose_err, report

pytestmark = pytest.mark.gpu

PLANE_TOL = 1e-5


# Observed on the B200 (gpurun_out/parity_report.log, round 2): ZERO cells beyond 1e-5 in every comparison of this file
# (66 k - 116 k touched cells each, largest difference 4.8e-7 = the ((l+lf)-lf)+lo vs l+lo ulp).  The bar is the survey's
# own observation for the reference against itself under reordering, 2 cells (2 in 13 M touches) — an
# off-by-one-cell bug in a beam end or a Bresenham carry moves hundreds of cells and cannot hide in it.
MAX_BAD_CELLS = 2


def compare_planes(got, want, what, max_bad=MAX_BAD_CELLS):
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bad = diff > PLANE_TOL
    touched = max(1, int((want != 0).sum()))
    report(f"planes[{what}]: {int(bad.sum())} cells differ by > {PLANE_TOL} of {touched} touched (max |diff| {float(diff.max()):.3e})")
    assert bad.sum() <= max_bad, (what, int(bad.sum()), touched, float(diff.max()))
    return int(bad.sum())


@pytest.mark.parametrize("mode", [1, 2])
def test_update_by_scan_goldens(hsb_lib, mode):
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4,
                          update_factor_occupied=0.9, gather_mode=mode)
    base = golden_planes(g)
    for l, p in enumerate(base):
        rep.upload_level(l, p)
    rep.matchData(g["hints"][0], g["scans"][0])
    rep.updateByScan(g["scans"][0], g["upd1_pose"])
    rep.onMapUpdated()
    want1 = apply_diff(base, g, "upd1")
    for l in range(3):
        compare_planes(rep.download_level(l), want1[l], f"upd1 level {l}")
    # second write: level 0 from the given scan, coarse levels from the last MATCHED scan (Q11)
    rep.matchData(g["hints"][1], g["scans"][1])
    rep.updateByScan(g["scans"][2], g["upd2_pose"])
    rep.onMapUpdated()
    want2 = apply_diff(want1, g, "upd2")
    for l in range(3):
        compare_planes(rep.download_level(l), want2[l], f"upd2 level {l}")
    # the probability planes follow the log-odds (fused refresh): P == f(l) on every cell
    for l in range(3):
        lo = rep.download_level(l).astype(np.float64)
        p = rep.download_prob(l).astype(np.float64)
        assert np.abs(p - np.exp(lo) / (np.exp(lo) + 1.0)).max() < 2e-7
    # and the next match sees the new map
    pose, cov = rep.matchData(g["hints"][3], g["scans"][3])
    ex, ey, ea = pose_err(pose, g["after_upd_pose"])
    assert max(ex, ey) <= 1e-4 and ea <= 1e-4
    rep.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_single_level_update_against_oracle(hsb_lib, pyoracle, oracle_kinds, mode):
    """OccGridMapBase::updateByScan on one level, many scans in a row, incl. beams leaving the map,
    a start cell outside the map and a non-zero origo."""
    from hector_slam_b200 import capi, synth

    kind = "reference" if "reference" in oracle_kinds else "port"
    world = synth.World(1, seed=5)
    orc = pyoracle.Oracle(kind, 0.05, 512, 1)
    orc.set_update_factors(0.4, 0.9)
    rep = capi.MapRepB200(0.05, 512, levels=1, update_factor_free=0.4, update_factor_occupied=0.9, gather_mode=mode)
    rng = np.random.default_rng(8)
    poses = world.sample_free_poses(40, rng)
    for k, p in enumerate(poses):
        scan = synth.make_scan(world, p, rng)
        origo = np.float32([0.0, 0.0]) if k % 3 else np.float32([2.5, -1.5])
        pw = p.astype(np.float32)
        if k == 7:
            pw = np.float32([11.5, 3.0, 0.4])     # most beams leave the 25.6 m map
        if k == 9:
            pw = np.float32([40.0, 0.0, 0.0])     # start cell outside: every beam dropped
        orc.update_level(0, scan, pw, origo)
        rep.update_level_by_scan(0, scan, pw, origo)
    nbad = compare_planes(rep.download_level(0), orc.get_logodds(0), "40 scans")
    print("cells differing after 40 scans:", nbad)
    rep.close()
    orc.close()


def test_update_by_long_scans(hsb_lib, pyoracle, oracle_kinds):
    """Scans with more beams than one wave of the mark grid holds (the grid is capped at the resident CTAs, a team then
    walks several beams): 5405 and 21 620 endpoints per updateByScan on a 3-level map, planes against the oracle."""
    from hector_slam_b200 import capi, synth

    kind = "reference" if "reference" in oracle_kinds else "port"
    world = synth.World(1, seed=6)
    orc = pyoracle.Oracle(kind, 0.05, 1024, 3)
    orc.set_update_factors(0.4, 0.9)
    rep = capi.MapRepB200(0.05, 1024, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    rng = np.random.default_rng(9)
    poses = world.sample_free_poses(6, rng)
    for k, p in enumerate(poses):
        reps = 5 if k % 2 == 0 else 20
        scan = np.concatenate([synth.make_scan(world, p, np.random.default_rng(100 * k + r)) for r in range(reps)])
        pw = p.astype(np.float32)
        orc.match(pw, scan)                    # fills the coarse-level containers on both sides (MapRepMultiMap.h:127)
        rep.matchData(pw, scan)
        orc.update_by_scan(scan, pw)
        rep.updateByScan(scan, pw)
    for l in range(3):
        compare_planes(rep.download_level(l), orc.get_logodds(l), f"long scans level {l}")
    rep.close()
    orc.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_slam_run(hsb_lib, mode):
    """The HectorSlamProcessor::update sequence (match -> gate -> updateByScan -> onMapUpdated)
    starting on an EMPTY map (Q12: first match returns the hint) against the golden run."""
    from hector_slam_b200 import capi

    g = load_golden("slam3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4,
                          update_factor_occupied=0.9, gather_mode=mode)
    hint = g["first_hint"]
    for k in range(g["scans"].shape[0]):
        pose, cov = rep.matchData(hint, g["scans"][k])
        ex, ey, ea = pose_err(pose, g["est"][k])
        assert max(ex, ey) <= 1e-4 and ea <= 1e-4, (k, ex, ey, ea)
        rep.updateByScan(g["scans"][k], pose)   # thresholds 0: every scan writes
        rep.onMapUpdated()
        hint = pose
    final = golden_planes(g, "final")
    for l in range(3):
        compare_planes(rep.download_level(l), final[l], f"final level {l}")
    rep.reset()
    for l in range(3):
        assert np.all(rep.download_level(l) == 0) and np.all(rep.download_prob(l) == 0.5)
    rep.close()
