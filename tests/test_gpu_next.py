"""GPU: the rows after / around the path that were widened into.
  N1  log-odds -> nav_msgs/OccupancyGrid int8 (HectorMappingRos::publishMap, src/HectorMappingRos.cpp:448-468)
  N3  pose likelihood (OccGridMapUtil::getLikelihoodForState, map/OccGridMapUtil.h:189-221), batched —
      ranks the relocalisation hypotheses of BASELINE config 4."""
import numpy as np
import pytest

from conftest import golden_planes, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [1, 2])
def test_occupancy_export_bit_exact(hsb_lib, mode):
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, gather_mode=mode)
    planes = golden_planes(g)
    planes[0][5, 7] = np.float32(-0.0)   # isFree / isOccupied are strict comparisons with 0
    planes[0][5, 8] = np.float32(1e-30)
    planes[0][5, 9] = np.float32(-1e-30)
    for l, p in enumerate(planes):
        rep.upload_level(l, p)
    for l, p in enumerate(planes):
        want = np.full(p.shape, -1, np.int8)
        want[p < 0] = 0
        want[p > 0] = 100
        got = rep.download_occupancy(l)
        assert got.dtype == np.int8 and np.array_equal(got, want)
    assert got.min() == -1 and (rep.download_occupancy(0) == 100).sum() > 1000
    rep.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_likelihood_batch(hsb_lib, pyoracle, oracle_kinds, mode):
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    kind = "reference" if "reference" in oracle_kinds else "port"
    orc = pyoracle.Oracle(kind, float(g["res"]), int(g["size"]), 3)
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, gather_mode=mode)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
        orc.set_logodds(l, p)
    K = g["scans"].shape[0]
    pts = g["scans"].reshape(-1, 2)
    offs = (np.arange(K + 1) * g["scans"].shape[1]).astype(np.int32)
    poses = g["hints"].copy()
    poses[3, 0] += 40.0                      # far outside the map: every endpoint counts 1 -> likelihood 0
    for l in range(3):
        got = rep.likelihood_batch(l, poses, pts, offs)
        want = np.float32([orc.likelihood(l, orc.map_coords_pose(l, poses[k]),
                                          (g["scans"][k] * np.float32(2.0 ** -l)).astype(np.float32)) for k in range(K)])
        assert np.abs(got - want).max() <= 2e-6, (l, np.abs(got - want).max())
        assert got[3] == 0.0
    # ranking hypotheses of ONE scan: the matched pose scores higher than its perturbed hint
    hyp = np.stack([g["hints"][0], g["ref_poses"][0], g["truth"][0].astype(np.float32)])
    sc = rep.likelihood_batch(0, hyp, g["scans"][0], None)
    assert sc[1] > sc[0] and sc[2] > sc[0]
    # the device-pointer entry point (what a relocalisation step chains behind hsb_match_batch_device): same bits,
    # per-scan offsets and the shared-scan layout, on a caller's stream
    import torch
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream(device=dev)
    d_poses, d_pts, d_offs = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (poses, pts, offs))
    d_out = torch.full((K,), -7.0, dtype=torch.float32, device=dev)
    for l in range(3):
        with torch.cuda.stream(side):
            rep.likelihood_batch_device(l, K, d_poses.data_ptr(), d_pts.data_ptr(), d_offs.data_ptr(), 0, d_out.data_ptr(),
                                        side.cuda_stream)
        side.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), rep.likelihood_batch(l, poses, pts, offs))
    d_hyp, d_scan = torch.from_numpy(hyp).to(dev), torch.from_numpy(np.ascontiguousarray(g["scans"][0])).to(dev)
    d_sc = torch.empty(3, dtype=torch.float32, device=dev)
    rep.likelihood_batch_device(0, 3, d_hyp.data_ptr(), d_scan.data_ptr(), None, g["scans"].shape[1], d_sc.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(d_sc.cpu().numpy(), sc)
    # score + arg-max in one call (relocalisation): the winner of hsb_best_hypothesis_device is numpy's arg-max of the
    # host call's scores (first index among equals), a non-finite pose never wins, the accumulator re-arms itself
    rng = np.random.default_rng(4)
    many = np.repeat(g["ref_poses"][0][None], 300, axis=0).astype(np.float32)
    many[:, :2] += rng.uniform(-0.3, 0.3, (300, 2)).astype(np.float32)
    many[17] = many[203] = g["ref_poses"][0]          # two equal best candidates: the first one has to win
    many[5] = [np.nan, 0.0, 0.0]
    many[6] = [np.inf, 0.0, 0.0]
    ok = np.all(np.isfinite(many), axis=1)
    sc_all = np.full(300, -1.0, np.float32)
    sc_all[ok] = rep.likelihood_batch(0, many[ok], g["scans"][0], None)
    d_many = torch.from_numpy(many).to(dev)
    d_best = torch.zeros(4, dtype=torch.float32, device=dev)
    d_all = torch.empty(300, dtype=torch.float32, device=dev)
    for rep_i in range(2):
        rep.best_hypothesis_device(0, 300, d_many.data_ptr(), d_scan.data_ptr(), None, g["scans"].shape[1], d_best.data_ptr(),
                                   d_all.data_ptr(), 0)
        torch.cuda.synchronize()
        k = int(np.argmax(sc_all))
        got = d_best.cpu().numpy()
        assert got[0] == sc_all[k] and np.array_equal(got[1:], many[k]), (rep_i, k, got)
        assert np.array_equal(d_all.cpu().numpy()[ok], sc_all[ok])
    rep.best_hypothesis_device(0, 0, d_many.data_ptr(), d_scan.data_ptr(), None, g["scans"].shape[1], d_best.data_ptr(), None, 0)
    torch.cuda.synchronize()
    assert d_best.cpu().numpy()[0] == -1.0           # empty batch
    rep.close()
    orc.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_covariance_batch(hsb_lib, pyoracle, oracle_kinds, mode):
    """N3: getCovarianceForPose (OccGridMapUtil.h:106-160) + getCovMatrixWorldCoords (:162-187) against the compiled
    reference.  The seven likelihoods differ from the sequential sum by summation order only (<= 2e-6 each); the
    covariance is a weighted sum of squared +-1.5-cell / +-0.05-rad deviations, compared to 2e-5 of its largest entry."""
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    kind = "reference" if "reference" in oracle_kinds else "port"
    orc = pyoracle.Oracle(kind, float(g["res"]), int(g["size"]), 3)
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, gather_mode=mode)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
        orc.set_logodds(l, p)
    K = g["scans"].shape[0]
    pts = g["scans"].reshape(-1, 2)
    offs = (np.arange(K + 1) * g["scans"].shape[1]).astype(np.int32)
    poses = g["ref_poses"].copy()
    poses[5] = g["hints"][5]
    for l in range(3):
        cm, cw = rep.covariance_batch(l, poses, pts, offs)
        for k in range(K):
            wm, ww = orc.covariance_for_pose(l, orc.map_coords_pose(l, poses[k]),
                                             (g["scans"][k] * np.float32(2.0 ** -l)).astype(np.float32))
            assert np.abs(cm[k] - wm).max() <= 2e-5 * np.abs(wm).max(), (l, k, cm[k], wm)
            assert np.abs(cw[k] - ww).max() <= 2e-5 * np.abs(ww).max(), (l, k)
            assert np.array_equal(cw[k], cw[k].T)
        assert cm[:, 0, 0].min() > 0 and cm[:, 2, 2].min() > 0
    # shared-scan mode: B poses, one scan
    cm1, _ = rep.covariance_batch(0, poses[:4], g["scans"][2], None)
    wm, _ = orc.covariance_for_pose(0, orc.map_coords_pose(0, poses[1]), g["scans"][2])
    assert np.abs(cm1[1] - wm).max() <= 2e-5 * np.abs(wm).max()
    rep.close()
    orc.close()


def test_config4_best_hypothesis_by_likelihood(hsb_lib, pyoracle):
    """Config 4 end to end on one GPU: match 8192 hypotheses of one scan, score the results with the
    likelihood kernel, pick the best — it must be the in-basin fixed point (== the oracle's answer)."""
    from hector_slam_b200 import capi, synth

    world = synth.World(1, seed=1234)
    g = load_golden("match3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    orc = pyoracle.Oracle("port", float(g["res"]), int(g["size"]), 3)
    orc.set_update_factors(0.4, 0.9)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
        orc.set_logodds(l, p)
    truth, scan = g["truth"][2], g["scans"][2]
    rng = np.random.default_rng(9)
    B = 8192
    hyp = np.tile(truth, (B, 1))
    hyp[:, 0] += rng.uniform(-2.0, 2.0, B)
    hyp[:, 1] += rng.uniform(-2.0, 2.0, B)
    hyp[:, 2] += rng.uniform(-0.5, 0.5, B)
    hyp = hyp.astype(np.float32)
    poses, _ = rep.match_batch(hyp, scan, None)
    finite = np.all(np.isfinite(poses), axis=1)
    score = np.full(B, -1.0, np.float32)
    score[finite] = rep.likelihood_batch(0, poses[finite], scan, None)
    best = poses[int(np.argmax(score))]
    want, _ = orc.match(truth.astype(np.float32), scan)
    assert np.abs(best[:2] - want[:2]).max() < 5e-3 and abs(best[2] - want[2]) < 5e-3, (best, want)
    assert np.abs(best[:2] - truth[:2]).max() < 0.02
    rep.close()
    orc.close()


def test_raycast_batch_bit_exact(hsb_lib, pyoracle):
    """N4: hector_map_tools' checkOccupancyBresenhami, 20 000 rays at once, against the C restatement
    (integer work: distances, hit cells and misses must be identical)."""
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    size = int(g["size"])
    rep = capi.MapRepB200(float(g["res"]), size, levels=3)
    orc = pyoracle.Oracle("port", float(g["res"]), size, 3)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
        orc.set_logodds(l, p)
    rng = np.random.default_rng(4)
    for level in (0, 2):
        s = size >> level
        B = 20000 if level == 0 else 2000
        begin = rng.integers(int(0.3 * s), int(0.7 * s), (B, 2)).astype(np.int32)     # mostly inside the room
        end = rng.integers(-5, s + 5, (B, 2)).astype(np.int32)                       # some outside the map
        begin[:50] = rng.integers(-3, s + 3, (50, 2))                                  # some starts outside too
        end[50:60] = begin[50:60]                                                      # zero-length rays
        dist, hit = rep.raycast_batch(level, begin, end)
        n_hit = 0
        for b in range(0, B, 7 if level == 0 else 1):
            d, h = orc.raycast(level, begin[b], end[b])
            assert dist[b] == d and tuple(hit[b]) == h, (level, b, dist[b], d, hit[b], h)
            n_hit += d >= 0
        assert n_hit > 100
    rep.close()
    orc.close()


def test_raycast_and_get_dist_against_the_compiled_header(hsb_lib, pyoracle, oracle_kinds):
    """N4 pinned on the reference itself: hector_map_tools/HectorMapTools.h compiled unmodified (oracle/maptools_driver.cpp,
    nav_msgs stubbed) over the occupancy grid the device exports — checkOccupancyBresenhami (:148-214) cell for cell and
    getDist (:133-147) with its world <-> map legs (CoordinateTransformer, :41-116), float for float."""
    from hector_slam_b200 import capi

    g = load_golden("match3.npz")
    size = int(g["size"])
    rep = capi.MapRepB200(float(g["res"]), size, levels=3)
    port = pyoracle.Oracle("port", float(g["res"]), size, 3)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
        port.set_logodds(l, p)
    rng = np.random.default_rng(12)
    have_ref = "reference" in oracle_kinds
    for level in (0, 1):
        s = size >> level
        _, _, cell = rep.level_info(level)
        origin = rep.map_origin(level)
        assert np.array_equal(origin, port.map_origin(level))
        B = 6000
        half = 0.5 * s * cell
        bw = rng.uniform(-0.5 * half, 0.5 * half, (B, 2)).astype(np.float32)
        ew = rng.uniform(-1.05 * half, 1.05 * half, (B, 2)).astype(np.float32)      # some end outside the map
        bw[:20] = rng.uniform(-1.2 * half, 1.2 * half, (20, 2))
        ew[20:25] = bw[20:25]
        bw[25] = (np.nan, 0.0)
        dist, hit, found = rep.get_dist_batch(level, bw, ew)
        begin = rng.integers(int(0.3 * s), int(0.7 * s), (B, 2)).astype(np.int32)
        end = rng.integers(-5, s + 5, (B, 2)).astype(np.int32)
        cdist, chit = rep.raycast_batch(level, begin, end)
        mt = pyoracle.RefMapTools(rep.download_occupancy(level), cell, origin) if have_ref else None
        nfound = 0
        for b in range(0, B, 3):
            if b != 25:
                d, hw, f = mt.get_dist(bw[b], ew[b]) if have_ref else port.get_dist(level, bw[b], ew[b])
                assert dist[b] == np.float32(d) and found[b] == f, (level, b, dist[b], d)
                if f:
                    assert np.array_equal(hit[b], hw), (level, b, hit[b], hw)
                    nfound += 1
            d, h = mt.raycast(begin[b], end[b]) if have_ref else port.raycast(level, begin[b], end[b])
            assert cdist[b] == d and tuple(chit[b]) == h, (level, b)
        assert nfound > 200 and not found[25] and dist[25] == np.float32(-cell)
        if mt:
            mt.close()
    rep.close()
    port.close()


def test_dirty_rect_downloads_move_only_the_touched_area(hsb_lib):
    """N1 as SURVEY.md specifies it: dirty-rect int8 / log-odds download.  The mirror rectangle accumulates what K2
    wrote since the host last looked; the rect downloads equal the corresponding window of the full downloads and move
    bytes proportional to the window (hsb_get_d2h_bytes), not to the map."""
    from hector_slam_b200 import capi

    g = load_golden("slam3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    for l in range(3):
        assert rep.get_mirror_dirty_rect(l) is None          # a fresh handle is clean
    pose = g["first_hint"]
    for k in range(4):
        pose, _ = rep.matchData(pose, g["scans"][k])
        rep.updateByScan(g["scans"][k], pose)
    for l in range(3):
        sx, sy, _ = rep.level_info(l)
        rect = rep.get_mirror_dirty_rect(l, reset=True)
        assert rect is not None and rep.get_mirror_dirty_rect(l) is None
        assert rep.get_dirty_rect(l) == rect                  # the replication rectangle is independent: still set
        x0, y0, x1, y1 = rect
        full, occ = rep.download_level(l), rep.download_occupancy(l)
        assert np.all(full[:y0] == 0) and np.all(full[y1 + 1:] == 0) and np.all(full[:, :x0] == 0) and np.all(full[:, x1 + 1:] == 0)
        b0 = rep.d2h_bytes
        win = rep.download_level_rect(l, rect)
        b1 = rep.d2h_bytes
        owin = rep.download_occupancy_rect(l, rect)
        b2 = rep.d2h_bytes
        n = (x1 - x0 + 1) * (y1 - y0 + 1)
        assert b1 - b0 == 4 * n and b2 - b1 == n and n < sx * sy // 2
        assert np.array_equal(win, full[y0:y1 + 1, x0:x1 + 1]) and np.array_equal(owin, occ[y0:y1 + 1, x0:x1 + 1])
    # after a reset / an upload every cell may differ: the rectangles are the whole level
    rep.reset()
    assert rep.get_mirror_dirty_rect(0, reset=True) == (0, 0, int(g["size"]) - 1, int(g["size"]) - 1)
    assert rep.get_dirty_rect(1, reset=True) == (0, 0, (int(g["size"]) >> 1) - 1, (int(g["size"]) >> 1) - 1)
    rep.upload_level(2, np.ones((int(g["size"]) >> 2,) * 2, np.float32))
    assert rep.get_dirty_rect(2) == (0, 0, (int(g["size"]) >> 2) - 1, (int(g["size"]) >> 2) - 1)
    with pytest.raises(capi.HsbError):
        rep.download_level_rect(0, (5, 5, 4, 9))
    rep.close()


def test_gate_state_survives_mixing_fused_and_host_driven_steps(hsb_lib):
    """SlamProcessor::updateUnfused writes the map from the host; the fused step's device-side gate has to
    learn the new lastMapUpdatePose (hsb_set_last_map_update_pose) or the next fused step decides differently."""
    from hector_slam_b200 import capi

    g = load_golden("slam3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    rep.setMapUpdateMinDistDiff(0.4)
    rep.setMapUpdateMinAngleDiff(0.9)
    p0, _, upd = rep.slam_update(g["first_hint"], g["scans"][0])
    assert upd and np.array_equal(rep.last_map_update_pose(), p0)
    far = p0 + np.float32([1.0, 0.0, 0.0])
    rep._check(rep.lib.hsb_set_last_map_update_pose(rep.h, far.ctypes.data))
    assert np.array_equal(rep.last_map_update_pose(), far)
    # a fused step near p0 is now > 0.4 m away from the recorded pose: the gate fires
    _, _, upd2 = rep.slam_update(p0, g["scans"][1])
    assert upd2
    # and right after it does not
    _, _, upd3 = rep.slam_update(rep.last_map_update_pose(), g["scans"][1])
    assert not upd3
    rep.close()


def test_next_rows_against_reference_goldens(hsb_lib):
    """The CUDA path against tests/golden/next.npz — vectors produced by the compiled reference itself (node converter
    source, OccGridMapUtil.h, HectorMapTools.h): conversion, ray cast and getDist bit-exact, covariance to 2e-5."""
    from hector_slam_b200 import capi, synth

    g, m = load_golden("next.npz"), load_golden("match3.npz")
    size = int(m["size"])
    rep = capi.MapRepB200(float(m["res"]), size, levels=3)
    for l, p in enumerate(golden_planes(m)):
        rep.upload_level(l, p)
    co, ko = g["cloud_offsets"], g["cloud_kept_offsets"]
    for k in range(len(co) - 1):
        rep.set_cloud_format(g["cloud_T"][k], **synth.CLOUD_FORMAT)
        pts, og = rep.cloud_to_points(g["cloud_xyz"][co[k]:co[k + 1]])
        assert np.array_equal(pts, g["cloud_kept"][ko[k]:ko[k + 1]]) and np.array_equal(og, g["cloud_origo"][k])
    K = m["scans"].shape[0]
    pts = m["scans"].reshape(-1, 2)
    offs = (np.arange(K + 1) * m["scans"].shape[1]).astype(np.int32)
    for l in range(3):
        cm, cw = rep.covariance_batch(l, m["ref_poses"], pts, offs)
        assert np.abs(cm - g["cov_map"][l]).max(axis=(1, 2)).max() <= 2e-5 * np.abs(g["cov_map"][l]).max()
        assert (np.abs(cw - g["cov_world"][l]).max(axis=(1, 2)) <= 2e-5 * np.abs(g["cov_world"][l]).max(axis=(1, 2))).all()
    assert np.array_equal(rep.map_origin(0), g["map_origin"])
    dist, hit = rep.raycast_batch(0, g["ray_begin"], g["ray_end"])
    assert np.array_equal(dist, g["ray_dist"]) and np.array_equal(hit, g["ray_hit"])
    dist, hw, found = rep.get_dist_batch(0, g["gd_begin"], g["gd_end"])
    assert np.array_equal(dist, g["gd_dist"]) and np.array_equal(found, g["gd_found"].astype(bool))
    assert np.array_equal(hw[found], g["gd_hit"][found])
    rep.close()
