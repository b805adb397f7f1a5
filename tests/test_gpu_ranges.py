"""GPU: raw laser ranges in — HectorMappingRos::rosLaserScanToDataContainer
(hector_mapping/src/HectorMappingRos.cpp:483-507) fused into the match kernel's staging step.
The conversion is integer/byte-exact work apart from two fp32 products per endpoint, so it must be
BIT-EXACT against the oracle's restatement: same endpoints, same order, same count."""
import numpy as np
import pytest

from conftest import golden_planes, load_golden, pose_err

pytestmark = pytest.mark.gpu


def fmt():
    from hector_slam_b200 import synth

    return synth.SCAN_FORMAT


@pytest.mark.parametrize("mode", [1, 2])
def test_scan_to_points_bit_exact(hsb_lib, pyoracle, mode):
    from hector_slam_b200 import capi

    rep = capi.MapRepB200(0.05, 512, levels=1, gather_mode=mode)
    f = fmt()
    rep.set_scan_format(**f)
    rng = np.random.default_rng(0)
    scale = rep.getScaleToMap()
    for trial in range(12):
        r = rng.uniform(0.0, 35.0, f["n_beams"]).astype(np.float32)
        if trial == 1:
            r[:] = 50.0                                   # nothing valid
        if trial == 2:
            r[::3] = np.nan
            r[1::5] = np.inf
            r[7] = np.float32(f["range_min"])             # boundary: strictly greater required
            r[8] = np.nextafter(np.float32(f["range_min"]), np.float32(1))
            r[9] = np.float32(np.float32(f["range_max"]) - np.float32(0.1))  # strictly smaller required
            r[10] = np.nextafter(r[9], np.float32(0))
        if trial == 3:
            r[:] = 5.0                                    # everything valid
        got = rep.scan_to_points(r)
        want = pyoracle.scan_to_points(r, f["angle_min"], f["angle_increment"], f["range_min"], f["range_max"], scale)
        assert got.shape == want.shape, (trial, got.shape, want.shape)
        assert np.array_equal(got, want), trial
    # a different format (fewer beams, other limits) re-derives the table
    rep.set_scan_format(n_beams=360, angle_min=-3.14159, angle_increment=0.017453292, range_min=0.2, range_max=12.0)
    r = rng.uniform(0.0, 14.0, 360).astype(np.float32)
    got = rep.scan_to_points(r)
    want = pyoracle.scan_to_points(r, -3.14159, 0.017453292, 0.2, 12.0, scale)
    assert np.array_equal(got, want)
    rep.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_match_batch_ranges_equals_endpoint_path(hsb_lib, pyoracle, mode):
    """ranges -> (fused conversion) -> match  ==  oracle conversion -> hsb_match_batch -> ==(1e-4) CPU oracle."""
    from hector_slam_b200 import capi, synth

    g = load_golden("match3.npz")
    rep = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9,
                          gather_mode=mode)
    for l, p in enumerate(golden_planes(g)):
        rep.upload_level(l, p)
    f = fmt()
    rep.set_scan_format(**f)
    world = synth.World(1, seed=1234)
    rng = np.random.default_rng(77)
    B = 48
    poses = world.sample_free_poses(B, rng, margin=0.8)
    ranges = synth.make_range_batch(world, poses, noise_seed=5)
    # knock some beams out so that scans are ragged after conversion; one scan entirely invalid
    ranges[3, ::4] = 31.0
    ranges[7, 100:400] = np.nan
    ranges[11, :] = 0.05
    hints = synth.perturb_hints(poses, seed=2)
    scale = rep.getScaleToMap()
    chunks, offs = [], [0]
    for b in range(B):
        pts = pyoracle.scan_to_points(ranges[b], f["angle_min"], f["angle_increment"], f["range_min"], f["range_max"], scale)
        chunks.append(pts)
        offs.append(offs[-1] + pts.shape[0])
    pts = np.concatenate(chunks).astype(np.float32)
    offs = np.asarray(offs, np.int32)
    assert offs[12] - offs[11] == 0 and offs[4] - offs[3] < 1081
    for w in (0, 1, 2, 4, 8):
        rep.set_tuning(warps_per_scan=w)
        got_r, cov_r = rep.match_batch_ranges(hints, ranges)
        got_p, cov_p = rep.match_batch(hints, pts, offs)
        assert np.array_equal(got_r, got_p), w        # identical endpoints, identical launch shape
        assert np.array_equal(cov_r, cov_p), w
    assert np.array_equal(got_r[11], hints[11])       # empty container: the hint comes back
    orc = pyoracle.Oracle("port", float(g["res"]), int(g["size"]), 3)
    orc.set_update_factors(0.4, 0.9)
    for l, p in enumerate(golden_planes(g)):
        orc.set_logodds(l, p)
    want, _, _ = orc.match_batch(hints, pts, offs)
    ex, ey, ea = pose_err(got_r, want)
    assert max(ex, ey) <= 1e-4 and ea <= 1e-4
    rep.close()
    orc.close()
