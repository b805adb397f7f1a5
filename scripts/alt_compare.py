"""A/B check of instruction-level variants of K1 that must not change a single bit: the product library against a
variant build (default: the round-1 gather addressing and branchy accumulation, -DHSB_TLD4_OFFSET=0 -DHSB_PRED_ACC=0).
Both run the BASELINE batch (every auto shape from 8 warps per scan down to 1) and a fused SLAM run in separate
processes; poses, covariances and map planes are compared bitwise.

  python scripts/alt_compare.py            # builds the variant, runs both, prints the verdict
  python scripts/alt_compare.py --dump F   # (internal) run with the library HSB_LIB_PATH points to and dump results
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def dump(path):
    import torch

    import bench
    from hector_slam_b200 import capi

    dev = torch.device("cuda", 0)
    rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    world, poses, pts, offs, hints = bench.make_workload(0, 4096)
    bench.build_map_on_gpu(rep, world)
    out = {}
    for n in (4096, 1024, 512, 256):
        P, C = rep.match_batch(hints[:n], pts[: offs[n]], offs[: n + 1])
        out[f"pose{n}"], out[f"cov{n}"] = P, C
    d_pts, d_h, d_o = torch.from_numpy(pts).to(dev), torch.from_numpy(hints).to(dev), torch.from_numpy(offs).to(dev)
    d_p = torch.empty((4096, 3), dtype=torch.float32, device=dev)
    for name, kw in (("dev_auto", {}), ("dev_g28", dict(warps_per_scan=1, scans_per_block=28, pace=1))):
        rep.set_tuning(warps_per_scan=0, scans_per_block=0, pace=0)
        rep.set_tuning(**kw)
        rep.match_batch_device(4096, d_h.data_ptr(), d_pts.data_ptr(), d_o.data_ptr(), 0, 1081, d_p.data_ptr(), None,
                               torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out[name] = d_p.cpu().numpy().copy()
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, pace=0)
    like = rep.likelihood_batch(0, hints[:512], pts[: offs[512]], offs[:513])
    out["likelihood"] = like
    # a fused SLAM run on a fresh map (sparse map: the sensitive case)
    g = np.load(os.path.join(ROOT, "tests", "golden", "slam3.npz"))
    r2 = capi.MapRepB200(float(g["res"]), int(g["size"]), levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    r2.setMapUpdateMinDistDiff(0.0)
    r2.setMapUpdateMinAngleDiff(0.0)
    hint, traj = g["first_hint"], []
    for k in range(g["scans"].shape[0]):
        hint, _, _ = r2.slam_update(hint, g["scans"][k])
        traj.append(hint)
    out["slam_traj"] = np.asarray(traj)
    for l in range(3):
        out[f"slam_plane{l}"] = r2.download_level(l)
    np.savez(path, **out)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--dump":
        dump(sys.argv[2])
        return
    from hector_slam_b200 import build

    alt = os.path.join(build.LIBDIR, "libhsb200_alt.so")
    flags = sys.argv[1:] or ["-DHSB_TLD4_OFFSET=0", "-DHSB_PRED_ACC=0"]
    build.build_cuda(force=True, extra=flags, out=alt)
    build.build_cuda()
    a, b = "/tmp/hsb_ab_main.npz", "/tmp/hsb_ab_alt.npz"
    subprocess.run([sys.executable, __file__, "--dump", a], check=True)
    subprocess.run([sys.executable, __file__, "--dump", b], check=True, env=dict(os.environ, HSB_LIB_PATH=alt))
    A, B = np.load(a), np.load(b)
    ok = True
    for k in A.files:
        same = np.array_equal(A[k], B[k], equal_nan=True)
        ok &= same
        print(f"{k:14s} {'bit-identical' if same else 'DIFFERENT: max |diff| %.3e' % np.nanmax(np.abs(A[k].astype(np.float64) - B[k]))}")
    print("variant flags:", " ".join(flags))
    print("VERDICT:", "all outputs bit-identical" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
