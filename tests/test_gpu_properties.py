"""GPU, BASELINE.json's full size (4096 x 1081-pt scans, 3-level 2048^2 map): size-independent
properties of the matcher, plus an oracle spot check on a random subset.

  determinism      the same launch twice gives bit-identical poses and covariances
  batch order      permuting the scans of a batch permutes the results, nothing else
  shared scan      hypothesis mode (one scan, B hints) == the same scan replicated B times
  fixed point      re-matching from a result returns (almost) the result: the 14 evaluations of a
                   3-level match end on the Gauss-Newton fixed point
  hint independence  two different in-basin hints for the same scan end on the same pose
  ranges == endpoints  the fused scan conversion feeds the matcher the same endpoints
"""
import numpy as np
import pytest

from conftest import pose_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(hsb_lib):
    import bench
    from hector_slam_b200 import capi, synth

    rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    world, poses, pts, offs, hints = bench.make_workload(0, bench.BATCH)
    bench.build_map_on_gpu(rep, world)
    rep.set_scan_format(**synth.SCAN_FORMAT)
    yield rep, world, poses, pts, offs, hints, bench.make_workload.ranges
    rep.close()


def test_full_size_properties(setup, pyoracle):
    rep, world, poses, pts, offs, hints, ranges = setup
    B = hints.shape[0]
    assert B == 4096 and pts.shape[0] == B * 1081
    P1, C1 = rep.match_batch(hints, pts, offs)
    P2, C2 = rep.match_batch(hints, pts, offs)
    assert np.array_equal(P1, P2) and np.array_equal(C1, C2)                      # determinism
    assert np.all(np.isfinite(P1))
    err = np.abs(P1[:, :2] - poses[:, :2]).max(axis=1)
    assert np.quantile(err, 0.99) < 0.02 and (err < 0.05).mean() > 0.995           # lands on the truth

    perm = np.random.default_rng(0).permutation(B)                                  # batch order
    scans = pts.reshape(B, 1081, 2)
    Pp, Cp = rep.match_batch(hints[perm], scans[perm].reshape(-1, 2), offs)
    assert np.array_equal(Pp, P1[perm]) and np.array_equal(Cp, C1[perm])

    k = 17                                                                          # shared scan
    hyp = np.repeat(hints[k][None], 512, axis=0).copy()
    hyp[:, 0] += np.linspace(-0.1, 0.1, 512, dtype=np.float32)
    Ps, _ = rep.match_batch(hyp, scans[k], None)
    Pr, _ = rep.match_batch(hyp, np.tile(scans[k], (512, 1)), (np.arange(513) * 1081).astype(np.int32))
    assert np.array_equal(Ps, Pr)
    assert np.abs(Ps - Ps[256]).max() < 1e-5                                        # hint independence

    P3, _ = rep.match_batch(P1, pts, offs)                                          # fixed point
    ex, ey, ea = pose_err(P3, P1)
    moved = np.abs(P3 - P1).max(axis=1)
    assert np.quantile(moved, 0.99) < 1e-5 and (moved < 1e-4).mean() > 0.995, (ex, ey, ea)

    Pq, Cq = rep.match_batch_ranges(hints, ranges)                                  # ranges == endpoints
    d = np.abs(Pq - P1).max(axis=1)   # synth's endpoints use numpy's cos/sin, the library glibc's: 1-ulp input differences
    assert np.quantile(d, 0.99) < 2e-5 and d.max() < 1e-3

    # oracle spot check on 192 random scans of the full batch
    orc = pyoracle.Oracle("port", 0.05, 2048, 3)
    orc.set_update_factors(0.4, 0.9)
    for l in range(3):
        orc.set_logodds(l, rep.download_level(l))
    idx = np.sort(np.random.default_rng(1).choice(B, 192, replace=False))
    want, _, _ = orc.match_batch(hints[idx], scans[idx].reshape(-1, 2), (np.arange(193) * 1081).astype(np.int32), nthreads=8)
    ok = np.abs(want[:, :2] - hints[idx][:, :2]).max(axis=1) < 0.5
    ex, ey, ea = pose_err(P1[idx][ok], want[ok])
    assert ok.mean() > 0.97 and max(ex, ey) <= 1e-4 and ea <= 1e-4
    orc.close()
