"""GPU: the CUDA scan matcher through the C-ABI against the golden vectors (generated from the
compiled reference) and against the CPU oracle on fresh seeded inputs.

Parity bar (BASELINE.json north_star): final pose within 1e-4 m / 1e-4 rad of the reference CPU
matcher on identical scans and maps.  Per-evaluation H / dTr: relative 1e-4 of the largest entry
 — summation order differs (lane-strided partial sums + shuffles vs sequential).
"""
import numpy as np
import pytest

from conftest import golden_planes, load_golden, pose_err, report

pytestmark = pytest.mark.gpu

POS_TOL = 1e-4
ANG_TOL = 1e-4

MODES = [1, 2]  # HSB_GATHER_LDG, HSB_GATHER_TEX


def make_rep(capi, g, mode, max_iterations=None, levels=None):
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=int(g["levels"]) if levels is None else levels,
                          max_iterations=max_iterations, update_factor_free=0.4, update_factor_occupied=0.9,
                          gather_mode=mode)
    return rep


def upload(rep, planes):
    for l, p in enumerate(planes):
        rep.upload_level(l, p)


def check_poses(got, want, what=""):
    ex, ey, ea = pose_err(got, want)
    assert ex <= POS_TOL and ey <= POS_TOL and ea <= ANG_TOL, (what, ex, ey, ea)


def check_mat(got, want, rel=1e-4, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1e-6)
    assert np.abs(got - want).max() <= rel * scale, (what, np.abs(got - want).max(), scale)


@pytest.mark.parametrize("mode", MODES)
def test_pose_conversions_and_constants(hsb_lib, mode):
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    rep = make_rep(capi, g, mode)
    assert rep.gather_mode == mode
    assert rep.getMapLevels() == 3 and rep.getScaleToMap() == np.float32(1.0) / np.float32(0.05)
    for k in range(4):
        for l in range(3):
            m = rep.map_coords_pose(l, g["conv_world"][k])
            assert np.array_equal(m, g["conv_map"][k, l])  # host arithmetic: bit-exact
            assert np.array_equal(rep.world_coords_pose(l, m), g["conv_back"][k, l])
    assert np.array_equal(rep.logodds_increments(), g["increments"])
    rep.close()


@pytest.mark.parametrize("mode", MODES)
def test_probability_plane(hsb_lib, mode):
    """K3: P = e^l/(e^l+1) for arbitrary log-odds incl. saturated cells (GridMapLogOdds.h:163-166)."""
    from hector_slam_b200 import capi

    g = load_golden("interp.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=int(g["levels"]), gather_mode=mode)
    for l in range(int(g["levels"])):
        rep.upload_level(l, g[f"plane{l}"])
        assert np.array_equal(rep.download_level(l), g[f"plane{l}"])
        got, want = rep.download_prob(l), g[f"prob{l}"]
        assert np.array_equal(got, want)         # expf as glibc evaluates it (sincosf_glibc.h): bit for bit
    rep.close()


def test_probability_plane_whole_range(hsb_lib, pyoracle, oracle_kinds):
    """Every log-odds value a cell can reach — the free side is not clamped, so long runs go far below expf's underflow
    thresholds (-103.28, -103.97), the occupied side stops at 50 + lo — against getGridProbability of the compiled
    reference, bit for bit."""
    from hector_slam_b200 import capi

    kind = "reference" if "reference" in oracle_kinds else "port"
    size = 512
    rng = np.random.default_rng(11)
    plane = np.empty((size, size), np.float32)
    plane[:128] = rng.uniform(-120.0, 52.2, (128, size))
    plane[128:256] = rng.uniform(-104.2, -103.0, (128, size))          # around the underflow thresholds
    lf, lo = np.float32(np.log(np.float32(0.4) / np.float32(0.6))), np.float32(np.log(np.float32(0.9) / np.float32(0.1)))
    acc = np.zeros(size, np.float32)
    for r in range(256, 384):                                           # accumulated exactly like a cell: k free steps
        acc = (acc + lf).astype(np.float32)
        plane[r] = acc
    plane[384:] = (rng.integers(0, 30, (128, size)).astype(np.float32) * lo + rng.integers(0, 300, (128, size)).astype(np.float32) * lf)
    plane[0, :8] = [0.0, -0.0, 50.0, 52.197224, -103.27893, -103.97208, -1e30, 88.0]
    orc = pyoracle.Oracle(kind, 0.05, size, 1)
    orc.set_logodds(0, plane)
    rep = capi.MapRepB200(0.05, size, levels=1)
    rep.upload_level(0, plane)
    got, want = rep.download_prob(0), orc.get_prob(0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), int((got != want).sum())
    rep.close()
    orc.close()


@pytest.mark.parametrize("mode", MODES)
def test_hessian_derivs_interpolation_bounds(hsb_lib, mode):
    """One evaluation on random planes: bilinear value, the reference's gradient blend (Q1), the
    inclusive [0, S-2] domain (Q6), out-of-map endpoints contribute nothing."""
    from hector_slam_b200 import capi

    g = load_golden("interp.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=int(g["levels"]), gather_mode=mode,
                          max_iterations=[-1, -1])
    for l in range(int(g["levels"])):
        rep.upload_level(l, g[f"plane{l}"])
    for i in range(len(g["ev_level"])):
        l = int(g["ev_level"][i])
        H, d = rep.hessian_derivs(l, g["ev_pose"][i], g["ev_pts"][i])
        both = np.concatenate([g["ev_H"][i], g["ev_dTr"][i]])
        check_mat(np.concatenate([H.reshape(9), d]), both, rel=2e-5, what=f"trial {i}")
    rep.close()


@pytest.mark.parametrize("mode", MODES)
def test_single_gauss_newton_step_and_clamp(hsb_lib, mode):
    """maxIterations = 0 -> exactly one evaluation + step (ScanMatcher.h:74): the 3x3 solve and
    the +-0.2 rad clamp (:209-215)."""
    from hector_slam_b200 import capi

    g = load_golden("interp.npz")
    n = len(g["ev_level"])
    for l in range(int(g["levels"])):
        # a 1-level handle whose only level is the golden's level l
        size = int(g["size"]) >> l
        rep = capi.MapRepB200(float(g["res"]) * (2 ** l), size, levels=1, gather_mode=mode, max_iterations=[-1])
        rep.upload_level(0, g[f"plane{l}"])
        for i in range(n):
            if int(g["ev_level"][i]) != l:
                continue
            pose, cov = rep.matchData(g["step_hint"][i], g["ev_pts"][i])
            want = g["step_pose"][i]
            step = np.abs(want - g["step_hint"][i]).max()
            tol = max(1e-5, 2e-4 * step) * max(1.0, g["step_cond"][i] / 100.0)
            assert np.abs(pose - want).max() <= tol, (i, pose, want, tol)
            if g["step_clamped"][i]:
                assert abs(abs(float(pose[2]) - float(g["step_hint"][i][2])) - 0.2) < 1e-5
            if i >= 2:  # trials 0/1 sit exactly on the bounds; the world<->map round trip moves them
                check_mat(cov.reshape(9), g["ev_H"][i], rel=2e-5)
        rep.close()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["match3.npz", "match1.npz"])
def test_match_goldens(hsb_lib, mode, name):
    from hector_slam_b200 import capi

    g = load_golden(name)
    rep = make_rep(capi, g, mode)
    upload(rep, golden_planes(g))
    K = g["scans"].shape[0]
    # one at a time (hsb_match_data)
    for k in range(K):
        pose, cov = rep.matchData(g["hints"][k], g["scans"][k])
        check_poses(pose, g["ref_poses"][k], f"scan {k}")
        check_mat(cov.reshape(9), g["ref_cov"][k], rel=1e-3)
    # as one batch (hsb_match_batch), several launch shapes
    pts = g["scans"].reshape(-1, 2)
    offs = (np.arange(K + 1) * g["scans"].shape[1]).astype(np.int32)
    for w, s in ((0, 0), (1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (8, 1), (16, 1)):
        rep.set_tuning(warps_per_scan=w, scans_per_block=s)
        P, C = rep.match_batch(g["hints"], pts, offs)
        check_poses(P, g["ref_poses"], f"batch W={w}")
        for k in range(K):
            check_mat(C[k].reshape(9), g["ref_cov"][k], rel=1e-3)
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=0)
    P, _ = rep.match_batch(g["hints"], pts, offs)
    check_poses(P, g["ref_poses"], "no smem staging")
    if mode == 2:
        # experimental f32x2 (FFMA2/FADD2) evaluation path: same per-endpoint operation sequence,
        # so the same launch shape must give bit-identical poses
        for w in (1, 2, 4):
            rep.set_tuning(warps_per_scan=w, scans_per_block=1, stage_smem=1, packed=0)
            A, _ = rep.match_batch(g["hints"], pts, offs)
            rep.set_tuning(packed=1)
            Bp, _ = rep.match_batch(g["hints"], pts, offs)
            assert np.array_equal(A, Bp), w
        rep.set_tuning(packed=0)
    rep.close()


@pytest.mark.parametrize("mode", MODES)
def test_three_evaluation_variant(hsb_lib, mode):
    from hector_slam_b200 import capi

    g = load_golden("match1.npz")
    rep = make_rep(capi, g, mode, max_iterations=[2])
    upload(rep, golden_planes(g))
    for k in range(g["scans"].shape[0]):
        pose, cov = rep.matchData(g["hints"][k], g["scans"][k])
        check_poses(pose, g["ref_poses_3eval"][k])
    rep.close()


@pytest.mark.parametrize("mode", MODES)
def test_edge_cases(hsb_lib, mode):
    from hector_slam_b200 import capi

    g = load_golden("match1.npz")
    rep = make_rep(capi, g, mode)
    upload(rep, golden_planes(g))
    # empty scan: pose = hint exactly, covariance untouched (ScanMatcher.h:68,189)
    cov_in = np.arange(9, dtype=np.float32)
    pose, cov = rep.matchData(g["hints"][0], np.zeros((0, 2), np.float32), cov_inout=cov_in)
    assert np.array_equal(pose, g["hints"][0]) and np.array_equal(cov.reshape(9), cov_in)
    # every endpoint out of the map: H = 0, gate fails, pose = hint after the round trip
    far = (g["scans"][0] + np.float32(1e5)).astype(np.float32)
    pose, cov = rep.matchData(g["hints"][0], far)
    assert np.array_equal(pose, g["edge_far_pose"]) and np.all(cov == 0)
    # non-finite endpoints are treated as out of map (the reference would index with them)
    bad = g["scans"][0].copy()
    bad[::7] = np.nan
    bad[3::11] = np.inf
    pose, cov = rep.matchData(g["hints"][0], bad)
    assert np.all(np.isfinite(pose)) and np.all(np.isfinite(cov))
    # ragged batch with an empty scan in the middle and odd offsets (bulk-copy alignment paths)
    s0, s1, s2 = g["scans"][0], g["scans"][1][:777], g["scans"][2][:1080]
    pts = np.concatenate([s0, s1, np.zeros((0, 2), np.float32), s2]).astype(np.float32)
    offs = np.int32([0, 1081, 1081 + 777, 1081 + 777, 1081 + 777 + 1080])
    hints = g["hints"][[0, 1, 5, 2]]
    covs = np.full((4, 9), 7.0, np.float32)
    P, C = rep.match_batch(hints, pts, offs, out_cov=covs)
    check_poses(P[0], g["ref_poses"][0])
    assert np.array_equal(P[2], hints[2]) and np.all(C[2] == 0.0)  # batch API: zero matrix for an empty scan
    for k, (sc, hk) in enumerate(((s0, 0), (s1, 1), (None, None), (s2, 3))):
        if sc is None:
            continue
        single, _ = rep.matchData(hints[k], sc)
        check_poses(P[k], single, f"ragged {k}")
    # shared-scan / pose-hypothesis mode equals per-item matching
    hyp = np.repeat(g["hints"][4][None], 9, axis=0).copy()
    hyp[:, 0] += np.linspace(-0.05, 0.05, 9, dtype=np.float32)
    P, _ = rep.match_batch(hyp, g["scans"][4], None)
    for k in range(9):
        single, _ = rep.matchData(hyp[k], g["scans"][4])
        check_poses(P[k], single)
    rep.close()


def test_bad_arguments_return_errors(hsb_lib):
    from hector_slam_b200 import capi

    with pytest.raises(capi.HsbError):
        capi.MapRepB200(0.05, 256, levels=9)
    with pytest.raises(capi.HsbError):
        capi.MapRepB200(0.05, 16, levels=4)
    rep = capi.MapRepB200(0.05, 256, levels=2)
    with pytest.raises(capi.HsbError):
        rep.upload_level(5, np.zeros((1, 1), np.float32)) if False else rep._check(
            rep.lib.hsb_upload_level(rep.h, 5, None))
    with pytest.raises(capi.HsbError):
        rep.set_tuning(nonsense=1)
    assert rep.launch_count >= 2
    rep.close()


@pytest.mark.parametrize("mode", MODES)
def test_batch_parity_against_oracle_2048_three_levels(hsb_lib, pyoracle, oracle_kinds, mode):
    """BASELINE.json config 2 at reduced batch: 3-level 2048^2 map built by the oracle, seeded
    scans and hints inside the convergence basin, every pose compared."""
    from hector_slam_b200 import capi, synth

    kind = "reference" if "reference" in oracle_kinds else "port"
    world = synth.World.for_map_size(2048)
    orc = pyoracle.Oracle(kind, 0.05, 2048, 3)
    orc.set_update_factors(0.4, 0.9)
    pyoracle.build_map_known_poses(orc, world)
    rng = np.random.default_rng(123)
    B = 384
    poses = world.sample_free_poses(B, rng)
    pts, offs = synth.make_scan_batch(world, poses, noise_seed=7)
    hints = synth.perturb_hints(poses, seed=1, dxy=0.1, dpsi=0.05)
    want, want_cov, _ = orc.match_batch(hints, pts, offs, nthreads=4)
    ok = np.abs(want[:, :2] - hints[:, :2]).max(axis=1) < 0.5  # drop (and count) oracle divergences (Q4)
    assert ok.mean() > 0.97
    rep = capi.MapRepB200(0.05, 2048, levels=3, update_factor_free=0.4, update_factor_occupied=0.9, gather_mode=mode)
    for l in range(3):
        rep.upload_level(l, orc.get_logodds(l))
    got, got_cov = rep.match_batch(hints, pts, offs)
    check_poses(got[ok], want[ok], "config-2 batch")
    assert np.abs(got[ok][:, :2] - poses[ok][:, :2]).max() < 0.03  # and close to the ground truth
    rep.close()
    orc.close()


def test_long_scans_maximum_sizes(hsb_lib, pyoracle, oracle_kinds):
    """Scans far longer than a lidar's 1081 beams (a projected 3-D cloud): 40 000 endpoints do not fit the 227 KB of
    shared memory of any launch shape, so the staged-prefix / unstaged paths of the single-scan (8 warps) and of the
    batch launch are what runs.  Against the oracle, plus a ragged batch mixing long, short and empty scans."""
    from hector_slam_b200 import capi, synth

    kind = "reference" if "reference" in oracle_kinds else "port"
    world = synth.World.for_map_size(2048)
    orc = pyoracle.Oracle(kind, 0.05, 2048, 3)
    orc.set_update_factors(0.4, 0.9)
    pyoracle.build_map_known_poses(orc, world)
    rep = capi.MapRepB200(0.05, 2048, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    for l in range(3):
        rep.upload_level(l, orc.get_logodds(l))
    rng = np.random.default_rng(77)
    poses = world.sample_free_poses(5, rng)
    hints = synth.perturb_hints(poses, seed=5, dxy=0.05, dpsi=0.02)
    longs = []
    for k, p in enumerate(poses):   # 37 noisy casts from the same pose, concatenated: 39 997 endpoints
        longs.append(np.concatenate([synth.make_scan(world, p, np.random.default_rng(1000 * k + r)) for r in range(37)]))
    assert all(s.shape[0] > 39000 for s in longs)
    # one long scan through hsb_match_data
    want, _ = orc.match(hints[0], longs[0])
    got, cov = rep.matchData(hints[0], longs[0])
    report(f"long scan (39 997 endpoints) single: pose diff {pose_err(got, want)}")
    check_poses(got, want, "long single scan")
    assert np.all(np.isfinite(cov))
    # ragged batch: long, short, empty, long, a 3-point scan, long ...
    short = synth.make_scan(world, poses[1], np.random.default_rng(9))
    items = [longs[0], short, np.zeros((0, 2), np.float32), longs[2], short[:3], longs[3], longs[4]]
    hh = hints[[0, 1, 1, 2, 1, 3, 4]]
    pts = np.ascontiguousarray(np.concatenate(items), dtype=np.float32)
    offs = np.concatenate([[0], np.cumsum([s.shape[0] for s in items])]).astype(np.int32)
    P, C = rep.match_batch(hh, pts, offs)
    W, _, _ = orc.match_batch(hh, pts, offs, nthreads=4)
    ok = [0, 1, 3, 5, 6]
    report(f"long scans ragged batch: pose diff {pose_err(P[ok], W[ok])}")
    check_poses(P[ok], W[ok], "ragged batch of long scans")
    assert np.array_equal(P[2], hh[2])            # empty scan: pose = hint
    assert np.all(np.isfinite(P)) and np.all(np.isfinite(C))
    # the fused converters keep the whole converted scan in shared memory: a cloud that cannot fit is refused, loudly
    rep.set_cloud_format(synth.laser_transform(), **synth.CLOUD_FORMAT)
    cloud = np.zeros((40000, 3), np.float32)
    cloud[:, 0] = np.linspace(1.0, 20.0, 40000, dtype=np.float32)
    with pytest.raises(capi.HsbError):
        rep.match_batch_cloud(hints[:1], cloud, np.int32([0, 40000]))
    got2, _ = rep.matchData(hints[0], longs[0])   # and the handle stays usable
    assert np.array_equal(got2, got)
    rep.close()
    orc.close()


@pytest.mark.parametrize("mode", MODES)
def test_non_square_map_other_resolution_and_start(hsb_lib, pyoracle, oracle_kinds, mode):
    """map_size_x != map_size_y, resolution 0.025 (the node's default), start coords off-centre,
    two levels: SLAM a few scans on both sides, then batch-match; planes and poses must agree."""
    from hector_slam_b200 import capi, synth

    kind = "reference" if "reference" in oracle_kinds else "port"
    res, sx, sy, start = 0.025, 1280, 896, (0.4, 0.55)
    orc = pyoracle.Oracle(kind, res, sx, 2, start=start, size_y=sy)
    orc.set_update_factors(0.4, 0.9)
    rep = capi.MapRepB200(res, sx, sy, levels=2, start=start, update_factor_free=0.4, update_factor_occupied=0.9,
                          gather_mode=mode)
    assert rep.level_info(1)[:2] == (640, 448) and orc.level_size(1) == (640, 448)
    world = synth.World(1, seed=21)
    rng = np.random.default_rng(3)
    scale = rep.getScaleToMap()
    assert scale == np.float32(1.0) / np.float32(res)
    poses = world.mapping_poses()[:10]
    for p in poses:
        scan = synth.make_scan(world, p, rng, scale_to_map=scale)
        p32 = p.astype(np.float32)
        orc.match(p32, scan)
        orc.update_by_scan(scan, p32)
        orc.on_map_updated()
        rep.matchData(p32, scan)
        rep.updateByScan(scan, p32)
        rep.onMapUpdated()
    for l in range(2):
        d = np.abs(rep.download_level(l) - orc.get_logodds(l))
        report(f"planes[non-square level {l}, mode {mode}]: {int((d > 1e-5).sum())} cells differ of {int((orc.get_logodds(l) != 0).sum())} touched")
        assert (d > 1e-5).sum() <= 2, (l, int((d > 1e-5).sum()))   # observed: 0 of 354 001 / 95 400 touched cells
        rep.upload_level(l, orc.get_logodds(l))      # continue from identical planes
    test_poses = world.sample_free_poses(64, rng, margin=0.8)
    pts, offs = synth.make_scan_batch(world, test_poses, noise_seed=5, scale_to_map=scale)
    hints = synth.perturb_hints(test_poses, seed=6, dxy=0.05, dpsi=0.03)
    want, _, _ = orc.match_batch(hints, pts, offs)
    got, _ = rep.match_batch(hints, pts, offs)
    ok = np.abs(want[:, :2] - hints[:, :2]).max(axis=1) < 0.5
    assert ok.mean() > 0.9
    check_poses(got[ok], want[ok], "non-square")
    rep.close()
    orc.close()
