"""GPU: point clouds in (the node's DEFAULT path) — HectorMappingRos::rosPointCloudToDataContainer
(hector_mapping/src/HectorMappingRos.cpp:509-542) fused into the match kernel's staging step.  Apart from double /
float products with fixed rounding the conversion is selection + compaction, so it must be BIT-EXACT against the
oracle (the node's own source text compiled in oracle/_ref where available, else the C port): same endpoints, same
order, same count, same origo."""
import numpy as np
import pytest

from conftest import golden_planes, load_golden, pose_err

pytestmark = pytest.mark.gpu


def conv_kind(oracle_kinds):
    return "reference" if "reference" in oracle_kinds else "port"


def oracle_cloud(pyoracle, cloud, T, fmt, scale, kind):
    return pyoracle.cloud_to_points(cloud, T, fmt["sqr_laser_min_dist"], fmt["sqr_laser_max_dist"], fmt["laser_z_min_value"],
                                    fmt["laser_z_max_value"], scale, kind=kind)


@pytest.mark.parametrize("mode", [1, 2])
def test_cloud_to_points_bit_exact(hsb_lib, pyoracle, oracle_kinds, mode):
    from hector_slam_b200 import capi, synth

    kind = conv_kind(oracle_kinds)
    rep = capi.MapRepB200(0.05, 512, levels=1, gather_mode=mode)
    scale = rep.getScaleToMap()
    rng = np.random.default_rng(0)
    fmt = dict(synth.CLOUD_FORMAT)
    for trial in range(14):
        T = synth.laser_transform(xyz=rng.uniform(-0.4, 0.4, 3), rpy=rng.uniform(-0.08, 0.08, 3))
        n = int(rng.integers(1, 1400))
        r = rng.uniform(0.0, 35.0, n)
        a = rng.uniform(-np.pi, np.pi, n)
        cloud = np.stack([r * np.cos(a), r * np.sin(a), rng.normal(0, 0.4, n)], axis=1).astype(np.float32)
        if trial == 1:
            cloud[:, :2] *= 0.001                         # everything inside laser_min_dist
        if trial == 2:                                    # the function's edges
            cloud[0] = (-0.5, 0.3, 0.0)                   # x < 0 and d^2 < 0.5: rejected          (:528)
            cloud[1] = (-0.7, 0.1, 0.0)                   # x < 0, d^2 = 0.5 exactly (fp32 0.49 + 0.01): not < 0.5 -> kept
            cloud[2] = (0.4, 0.0, 0.0)                    # d^2 = 0.16000001 vs sqr_min 0.16000001: strict >
            cloud[3] = (30.0, 0.0, 0.0)                   # d^2 = 900 vs sqr_max 900: strict <
            cloud[4] = (np.nan, 1.0, 0.0)
            cloud[5] = (1.0, np.inf, 0.0)
            cloud[6] = (2.0, 1.0, np.nan)
        if trial == 3:
            fmt = dict(fmt, laser_z_min_value=-0.05, laser_z_max_value=0.05)   # most points fall outside the z window
        if trial == 4:
            T = synth.laser_transform(xyz=(0, 0, 0), rpy=(0, 0, 0))           # identity
            fmt = dict(synth.CLOUD_FORMAT)
        rep.set_cloud_format(T, **fmt)
        got, origo = rep.cloud_to_points(cloud)
        want, want_origo = oracle_cloud(pyoracle, cloud, T, fmt, scale, kind)
        assert got.shape == want.shape, (trial, got.shape, want.shape)
        assert np.array_equal(got, want), trial
        assert np.array_equal(origo, want_origo), trial
    got, _ = rep.cloud_to_points(np.zeros((0, 3), np.float32))
    assert got.shape == (0, 2)
    rep.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_match_batch_cloud_equals_endpoint_path(hsb_lib, pyoracle, oracle_kinds, mode):
    """cloud -> (fused conversion) -> match == oracle conversion -> hsb_match_batch (bitwise, same launch shape)
    == (1e-4) the CPU oracle matching the converted scans."""
    from hector_slam_b200 import capi, synth

    kind = conv_kind(oracle_kinds)
    g = load_golden("match3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9,
                          gather_mode=mode)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
    scale = rep.getScaleToMap()
    world = synth.World(1, seed=1234)
    rng = np.random.default_rng(78)
    B = 48
    poses = world.sample_free_poses(B, rng, margin=0.8)
    ranges = synth.make_range_batch(world, poses, noise_seed=5)
    ranges[3, ::4] = 31.0
    ranges[11, :] = 0.05                                   # empty cloud
    # a laser mounted 12 cm ahead of / 31 cm above the base, slightly tilted; every third scan has its own transform
    T0 = synth.laser_transform()
    Ts = np.tile(T0, (B, 1))
    for b in range(0, B, 3):
        Ts[b] = synth.laser_transform(rpy=(0.01 * (b % 5), -0.02, 0.03))
    fmt = synth.CLOUD_FORMAT
    rep.set_cloud_format(T0, **fmt)
    clouds = [synth.ranges_to_cloud(ranges[b]) for b in range(B)]
    clouds[5][::7, 2] = 2.0                                # knocked out by the z window
    c_offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    cloud_all = np.concatenate(clouds).astype(np.float32)
    chunks, offs, origos = [], [0], []
    for b in range(B):
        pts, og = oracle_cloud(pyoracle, clouds[b], Ts[b], fmt, scale, kind)
        chunks.append(pts)
        origos.append(og)
        offs.append(offs[-1] + pts.shape[0])
    pts = np.concatenate(chunks).astype(np.float32)
    offs = np.asarray(offs, np.int32)
    assert offs[12] - offs[11] == 0 and offs[6] - offs[5] < c_offs[6] - c_offs[5]
    # the base pose that sees these endpoints: hints are for the BASE frame
    hints = synth.perturb_hints(poses, seed=2)
    for w in (0, 1, 2, 4, 8):
        rep.set_tuning(warps_per_scan=w)
        got_c, cov_c, og_c = rep.match_batch_cloud(hints, cloud_all, c_offs, transforms=Ts)
        got_p, cov_p = rep.match_batch(hints, pts, offs)
        assert np.array_equal(got_c, got_p), w
        assert np.array_equal(cov_c, cov_p), w
        assert np.array_equal(og_c, np.asarray(origos)), w
    assert np.array_equal(got_c[11], hints[11])
    # one transform for all (NULL transforms) == B copies of it
    rep.set_tuning(warps_per_scan=0)
    a, _, oa = rep.match_batch_cloud(hints, cloud_all, c_offs)
    b_, _, ob = rep.match_batch_cloud(hints, cloud_all, c_offs, transforms=np.tile(T0, (B, 1)))
    assert np.array_equal(a, b_) and np.array_equal(oa, ob)
    orc = pyoracle.Oracle(kind, float(g["res"]), int(g["size"]), 3)
    orc.set_update_factors(0.4, 0.9)
    for l, p in enumerate(golden_planes(g)):
        orc.set_logodds(l, p)
    want, _, _ = orc.match_batch(hints, pts, offs)
    ex, ey, ea = pose_err(got_c, want)
    assert max(ex, ey) <= 1e-4 and ea <= 1e-4
    rep.close()
    orc.close()


def test_cloud_path_feeds_the_map_writer(hsb_lib, pyoracle, oracle_kinds):
    """The container's origo (laser position * scaleToMap) is where updateByScan starts its beams
    (OccGridMapBase.h:134-137): cloud -> endpoints + origo -> hsb_update_by_scan equals the oracle's update."""
    from hector_slam_b200 import capi, synth

    kind = conv_kind(oracle_kinds)
    rep = capi.MapRepB200(0.05, 1024, levels=2, update_factor_free=0.4, update_factor_occupied=0.9)
    orc = pyoracle.Oracle(kind, 0.05, 1024, 2)
    orc.set_update_factors(0.4, 0.9)
    world = synth.World(1, seed=9)
    rng = np.random.default_rng(4)
    T = synth.laser_transform(xyz=(0.25, -0.1, 0.3), rpy=(0.0, 0.0, 0.1))
    rep.set_cloud_format(T, **synth.CLOUD_FORMAT)
    scale = rep.getScaleToMap()
    for p in world.sample_free_poses(6, rng):
        cloud = synth.ranges_to_cloud(world.cast(p) + rng.normal(0, 0.01, 1081))
        pts, origo = rep.cloud_to_points(cloud)
        want_pts, want_origo = oracle_cloud(pyoracle, cloud, T, synth.CLOUD_FORMAT, scale, kind)
        assert np.array_equal(pts, want_pts) and np.array_equal(origo, want_origo)
        p32 = p.astype(np.float32)
        rep.matchData(p32, pts, origo=origo)
        rep.updateByScan(pts, p32, origo=origo)
        orc.match(p32, want_pts, origo=want_origo)
        orc.update_by_scan(want_pts, p32, origo=want_origo)
        orc.on_map_updated()
    for l in range(2):
        d = np.abs(rep.download_level(l) - orc.get_logodds(l))
        assert (d > 1e-5).sum() <= 2, (l, int((d > 1e-5).sum()))
    rep.close()
    orc.close()


@pytest.mark.parametrize("nowait", [False, True])
def test_fused_cloud_slam_step_equals_the_two_call_path(hsb_lib, nowait):
    """hsb_slam_update_cloud (conversion in the match kernel's staging step, endpoints and their count handed to the map
    writer on the device) == hsb_cloud_to_points followed by hsb_slam_update with the returned origo: same poses, gate
    decisions, kept counts and planes — incl. map_without_matching steps, an empty cloud and a per-scan transform."""
    from hector_slam_b200 import capi, synth

    world = synth.World(1, seed=9)
    rng = np.random.default_rng(14)
    poses = world.mapping_poses()[:10]
    T0 = synth.laser_transform(xyz=(0.2, -0.05, 0.3), rpy=(0.01, -0.02, 0.04))
    clouds = [synth.ranges_to_cloud(world.cast(p) + rng.normal(0, 0.01, 1081)) for p in poses]
    clouds[3][::5, 2] = 1.7                    # a fifth of scan 3 leaves the z window
    clouds[6] = np.zeros((0, 3), np.float32)   # an empty cloud: pose = hint, nothing written

    def run(fused):
        rep = capi.MapRepB200(0.05, 1024, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
        rep.set_cloud_format(T0, **synth.CLOUD_FORMAT)
        rep.setMapUpdateMinDistDiff(0.0)
        rep.setMapUpdateMinAngleDiff(0.0)
        out = []
        for k, p in enumerate(poses):
            hint = p.astype(np.float32)
            T = synth.laser_transform(rpy=(0.0, 0.01 * k, 0.02)) if k % 4 == 1 else None
            without = k in (2, 7)
            if fused:
                pose, cov, upd, kept = rep.slam_update_cloud(hint, clouds[k], transform=T, map_without_matching=without,
                                                             nowait=nowait)
            else:
                if T is not None:
                    rep.set_cloud_format(T, **synth.CLOUD_FORMAT)
                pts, origo = rep.cloud_to_points(clouds[k])
                if T is not None:
                    rep.set_cloud_format(T0, **synth.CLOUD_FORMAT)
                pose, cov, upd = rep.slam_update(hint, pts, map_without_matching=without, origo=origo)
                kept = pts.shape[0]
            out.append((pose, upd, kept, cov))
        rep.onMapUpdated()
        planes = [rep.download_level(l) for l in range(3)]
        rep.close()
        return out, planes

    (a, pa), (b, pb) = run(True), run(False)
    for k, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x[0], y[0]) and x[1] == y[1] and x[2] == y[2], (k, x[:3], y[:3])
        if k not in (2, 6, 7):
            assert np.array_equal(x[3], y[3]), k
    assert a[6][2] == 0 and 0 < a[3][2] < a[0][2]
    for l in range(3):
        assert np.array_equal(pa[l], pb[l]), l
        assert (pa[l] != 0).sum() > 1000
