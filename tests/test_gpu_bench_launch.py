"""GPU: the launches bench.py TIMES, compared pose by pose with the CPU oracle (the compiled reference where
oracle/_ref is present) at the benchmark's own size.

VERDICT r01 weak #1: `value` runs hsb_match_batch_device at B = 4096 with the auto-selected shape (one warp per scan,
endpoints partly staged / read through L1), `e2e` runs the pipelined host calls (two halves, shape of a B/2 batch) — every one of
these, and every other shape the auto-launcher can pick (B = 256 .. 4096 -> 8, 4, 2, 1 warps per scan), is checked
here on ALL scans of the batch, not on a golden subset.  The workload is bench.py's (same generator, same seeds).
"""
import numpy as np
import pytest

from conftest import pose_err, report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workload(hsb_lib, pyoracle, oracle_kinds):
    import bench
    from hector_slam_b200 import capi

    kind = "reference" if "reference" in oracle_kinds else "port"
    B = 4096
    world, poses, pts, offs, hints = bench.make_workload(0, B)
    ranges = np.ascontiguousarray(bench.make_workload.ranges)
    orc = pyoracle.Oracle(kind, bench.RES, bench.MAP_SIZE, bench.LEVELS)
    orc.set_update_factors(0.4, 0.9)
    pyoracle.build_map_known_poses(orc, world)
    want, want_cov, _ = orc.match_batch(hints, pts, offs, nthreads=16)
    # the same batch as the NODE would see it: endpoints converted from the raw ranges by the reference's own converter
    # (bench.py's endpoint arrays come from numpy's cos / sin, which differ from glibc's cosf / sinf in the last bit)
    from hector_slam_b200 import synth
    f = synth.SCAN_FORMAT
    conv = [pyoracle.scan_to_points(ranges[b], f["angle_min"], f["angle_increment"], f["range_min"], f["range_max"], 20.0,
                                    kind=kind) for b in range(B)]
    r_offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in conv])]).astype(np.int32)
    want_r, _, _ = orc.match_batch(hints, np.concatenate(conv), r_offs, nthreads=16)
    rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=bench.LEVELS, update_factor_free=0.4,
                          update_factor_occupied=0.9)
    for l in range(bench.LEVELS):
        rep.upload_level(l, orc.get_logodds(l))
    orc.close()
    # scans on which the reference itself runs away are counted, not compared
    ok = np.abs(want[:, :2] - hints[:, :2]).max(axis=1) < 0.5
    assert ok.mean() > 0.98, ok.mean()
    yield dict(kind=kind, B=B, pts=pts, offs=offs, hints=hints, ranges=ranges, want=want, want_cov=want_cov, ok=ok,
               rep=rep, poses=poses, want_ranges=want_r)
    rep.close()


def compare(got, w, n, what, key="want"):
    ok = w["ok"][:n]
    ex, ey, ea = pose_err(got[ok], w[key][:n][ok])
    report(f"{what}: {int(ok.sum())} of {n} scans compared with the {w['kind']} oracle, max diff x {ex:.2e} y {ey:.2e} psi {ea:.2e}")
    assert max(ex, ey) <= 1e-4 and ea <= 1e-4, (what, ex, ey, ea)


def device_match(rep, w, n, **tuning):
    import torch

    dev = torch.device("cuda", 0)
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=1, unroll=0, partial=1, prefetch=0, auto_group=1, pace=0,
                   stagger=0)
    rep.set_tuning(**tuning)
    d_pts = torch.from_numpy(w["pts"][: w["offs"][n]]).to(dev)
    d_hints = torch.from_numpy(w["hints"][:n]).to(dev)
    d_offs = torch.from_numpy(w["offs"][: n + 1]).to(dev)
    d_poses = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d_cov = torch.empty((n, 9), dtype=torch.float32, device=dev)
    rep.match_batch_device(n, d_hints.data_ptr(), d_pts.data_ptr(), d_offs.data_ptr(), 0, 1081, d_poses.data_ptr(),
                           d_cov.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_poses.cpu().numpy(), d_cov.cpu().numpy(), rep.last_launch_shape()


def test_value_launch_all_4096(workload):
    """bench.py `value`: hsb_match_batch_device, B = 4096, auto shape = one CTA of 28 one-warp scans per SM."""
    w = workload
    got, cov, shape = device_match(w["rep"], w, w["B"])
    report(f"value launch shape: {shape}")
    assert shape["warps_per_scan"] == 1 and shape["scans_per_block"] == 28 and shape["grid"] == 147
    compare(got, w, w["B"], "value launch")
    ok = w["ok"]
    scale = np.abs(w["want_cov"][ok]).max(axis=(1, 2))
    assert (np.abs(cov[ok].reshape(-1, 3, 3) - w["want_cov"][ok]).max(axis=(1, 2)) <= 2e-3 * scale).all()
    # one scan per CTA (the shape of larger batches), with nothing / a prefix / everything staged in shared memory, with
    # the L2 prefetch, paced, staggered: which lane sums which endpoint in which order never changes -> the same bits
    got0, _, shape0 = device_match(w["rep"], w, w["B"], auto_group=0, partial=0)
    assert shape0["staged_points"] == 0 and shape0["scans_per_block"] == 1 and shape0["grid"] == 4096
    got1, _, shape1 = device_match(w["rep"], w, w["B"], auto_group=0, partial=1)
    assert 0 < shape1["staged_points"] < 1081
    got2, _, shape2 = device_match(w["rep"], w, w["B"], auto_group=0, stage_smem=2)
    assert shape2["staged_points"] == 1081
    assert np.array_equal(got, got0) and np.array_equal(got, got1) and np.array_equal(got, got2)
    for kw in (dict(prefetch=1), dict(pace=1), dict(stagger=150), dict(stage_smem=0)):
        gotv, _, _ = device_match(w["rep"], w, w["B"], **kw)
        assert np.array_equal(got, gotv), kw


@pytest.mark.parametrize("n,warps,groups", [(256, 8, 1), (512, 4, 1), (1024, 2, 1), (2048, 1, 14), (3000, 1, 21)])
def test_every_auto_shape(workload, n, warps, groups):
    w = workload
    got, _, shape = device_match(w["rep"], w, n)
    assert shape["warps_per_scan"] == warps and shape["scans_per_block"] == groups, shape
    compare(got, w, n, f"B={n} auto shape {shape}")


@pytest.mark.parametrize("unroll", [8])
def test_deeper_gather_batches(workload, unroll):
    w = workload
    got, _, shape = device_match(w["rep"], w, w["B"], unroll=unroll)
    assert shape["unroll"] == unroll
    compare(got, w, w["B"], f"U={unroll}")


def test_e2e_launches_all_4096(workload):
    """bench.py `e2e` (hsb_match_batch_ranges) and `e2e_endpoints` (hsb_match_batch): pinned host buffers, chunked
    copy/compute
    pipeline in two halves, both with the launch shape of a B/2 batch."""
    import torch

    from hector_slam_b200 import synth

    w = workload
    rep = w["rep"]
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=1, unroll=0, partial=1, prefetch=0)
    B = w["B"]
    h_pts = torch.from_numpy(w["pts"]).pin_memory()
    h_hints = torch.from_numpy(w["hints"]).pin_memory()
    got, _ = rep.match_batch(h_hints, h_pts, w["offs"])
    assert rep.last_launch_shape()["warps_per_scan"] == 1   # both halves of a 4096-scan call: shape of a 2048-scan batch
    compare(np.asarray(got), w, B, "e2e_endpoints (hsb_match_batch)")
    rep.set_scan_format(**synth.SCAN_FORMAT)
    h_ranges = torch.from_numpy(w["ranges"]).pin_memory()
    got_r, _ = rep.match_batch_ranges(h_hints, h_ranges)
    compare(np.asarray(got_r), w, B, "e2e (hsb_match_batch_ranges)", key="want_ranges")
    # a second call reuses the staging buffers (double-buffered across calls): same answer
    got_r2, _ = rep.match_batch_ranges(h_hints, h_ranges)
    assert np.array_equal(np.asarray(got_r), np.asarray(got_r2))
    # device-resident ranges, auto shape (one warp per scan converts + matches)
    dev = torch.device("cuda", 0)
    d_r, d_h = torch.from_numpy(w["ranges"]).to(dev), torch.from_numpy(w["hints"]).to(dev)
    d_p = torch.empty((B, 3), dtype=torch.float32, device=dev)
    rep.match_batch_ranges_device(B, d_h.data_ptr(), d_r.data_ptr(), d_p.data_ptr(), None,
                                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rep.last_launch_shape()["warps_per_scan"] == 1
    compare(d_p.cpu().numpy(), w, B, "ranges on the device, auto shape", key="want_ranges")


def test_submit_wait_pipeline_equals_blocking_calls(workload):
    """The submit / wait form (two staging sets, two calls in flight) returns the blocking calls' bits for all three
    wire formats, with page-locked buffers from hsb_alloc_pinned."""
    from hector_slam_b200 import capi, synth

    w = workload
    rep = w["rep"]
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=1, unroll=0, partial=0, prefetch=0)
    rep.set_scan_format(**synth.SCAN_FORMAT)
    B = 2048
    hints, ranges = w["hints"][:B], w["ranges"][:B]
    pts, offs = w["pts"][: w["offs"][B]], w["offs"][: B + 1]
    want_r, cov_r = rep.match_batch_ranges(hints, ranges)
    want_p, cov_p = rep.match_batch(hints, pts, offs)
    hp = [capi.pinned_copy(hints), capi.pinned_copy(np.ascontiguousarray(hints[::-1]))]
    rp = [capi.pinned_copy(ranges), capi.pinned_copy(np.ascontiguousarray(ranges[::-1]))]
    op = [capi.PinnedArray((B, 3)), capi.PinnedArray((B, 3))]
    oc = [capi.PinnedArray((B, 9)), capi.PinnedArray((B, 9))]
    t0 = rep.match_batch_ranges_submit(hp[0].array, rp[0].array, op[0].array, oc[0].array)
    t1 = rep.match_batch_ranges_submit(hp[1].array, rp[1].array, op[1].array, oc[1].array)   # both in flight
    assert {t0, t1} == {0, 1}
    rep.match_batch_wait(t0)
    assert np.array_equal(op[0].array, want_r) and np.array_equal(oc[0].array.reshape(B, 3, 3), cov_r)
    t2 = rep.match_batch_ranges_submit(hp[0].array, rp[0].array, op[0].array, oc[0].array)   # reuses set 0
    rep.match_batch_wait(t1)
    assert np.array_equal(op[1].array, want_r[::-1])          # a batch is a set of independent matches
    rep.match_batch_wait(t2)
    assert np.array_equal(op[0].array, want_r)
    pp = capi.pinned_copy(pts)
    t = rep.match_batch_submit(hp[0].array, pp.array, offs, op[1].array, oc[1].array)
    rep.match_batch_wait(t)
    assert np.array_equal(op[1].array, want_p) and np.array_equal(oc[1].array.reshape(B, 3, 3), cov_p)
    rep.match_batch_wait(-1)                                   # ticket of an empty batch: no-op
    gbs = rep.measure_h2d_gbs(rp[0].ptr, rp[0].array.nbytes)
    report(f"hsb_alloc_pinned buffer of {rp[0].array.nbytes / 1e6:.1f} MB: {gbs:.1f} GB/s host->device")
    assert gbs > 5.0
    for a in hp + rp + op + oc + [pp]:
        a.free()
