"""Time build variants of K1 against each other on the same box in one session (each variant is its own library,
loaded in a subprocess through HSB_LIB_PATH).  Usage:
  python scripts/variant_timing.py                 # build all variants, run them, print the table
  python scripts/variant_timing.py --run NAME      # (internal) time the library HSB_LIB_PATH points to
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "shipped flags, diagnostics compiled in": ["-DHSB_DIAG=1"],
    "shipped flags, diagnostics compiled OUT": ["-DHSB_DIAG=0"],
}
OLD_VARIANTS = {
    "uniform handle only": ["-DHSB_UNIFORM_HANDLE=1", "-DHSB_TLD4_OFFSET=0", "-DHSB_PRED_ACC=0"],
    "uniform + TLD4.AOFFI": ["-DHSB_UNIFORM_HANDLE=1", "-DHSB_TLD4_OFFSET=1", "-DHSB_PRED_ACC=0"],
    "uniform + predicated acc": ["-DHSB_UNIFORM_HANDLE=1", "-DHSB_TLD4_OFFSET=0", "-DHSB_PRED_ACC=1"],
    "uniform + AOFFI + predicated (all)": ["-DHSB_UNIFORM_HANDLE=1", "-DHSB_TLD4_OFFSET=1", "-DHSB_PRED_ACC=1"],
}


def run(name):
    import torch

    import bench
    from hector_slam_b200 import capi

    dev = torch.device("cuda", 0)
    rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
    world, poses, pts, offs, hints = bench.make_workload(0, 4096)
    bench.build_map_on_gpu(rep, world)
    stream = torch.cuda.current_stream().cuda_stream
    out = []
    for B, nbuf in ((4096, 8), (65536, 2)):
        reps = B // 4096
        p = np.tile(pts, (reps, 1))
        h = np.tile(hints, (reps, 1))
        o = (np.arange(B + 1) * bench.N_PTS).astype(np.int32)
        d_pts = [torch.from_numpy(p).to(dev).clone() for _ in range(nbuf)]
        d_h = torch.from_numpy(h).to(dev)
        d_o = torch.from_numpy(o).to(dev)
        d_p = torch.empty((B, 3), dtype=torch.float32, device=dev)
        if B == 4096:
            configs = [("auto (grouped, staged prefix)", dict(auto_group=1, partial=1)),
                       ("G=1 staged prefix", dict(stage_smem=1, partial=1)), ("G=1 unstaged", dict(stage_smem=0))]
        else:
            configs = [("G=1 staged (default)", dict(stage_smem=1, partial=1)), ("G=1 unstaged", dict(stage_smem=0))]
        for cname, kw in configs:
            rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=1, partial=0, pace=0, auto_group=0, stagger=0)
            rep.set_tuning(**kw)

            def go(i):
                rep.match_batch_device(B, d_h.data_ptr(), d_pts[i % nbuf].data_ptr(), d_o.data_ptr(), 0, bench.N_PTS,
                                       d_p.data_ptr(), None, stream)
            for i in range(5):
                go(i)
            torch.cuda.synchronize()
            best = 1e9
            for r in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                iters = 30 if B == 4096 else 5
                for i in range(iters):
                    go(i)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / iters)
            out.append(f"B={B:5d} {cname:28s} {best * 1e3:8.1f} us {B / best / 1e3:6.2f} M/s")
    # single-scan latency of the fused step (sync) on this map
    sc = [np.ascontiguousarray(pts[offs[i]:offs[i + 1]]) for i in range(32)]
    rep.set_tuning(warps_per_scan=0, scans_per_block=0, stage_smem=1, partial=0, stagger=0)
    rep.setMapUpdateMinDistDiff(0.0)
    rep.setMapUpdateMinAngleDiff(0.0)
    import time
    lat = []
    for i in range(150):
        t0 = time.perf_counter()
        rep.matchData(hints[i % 32], sc[i % 32])
        lat.append(time.perf_counter() - t0)
    out.append(f"single-scan hsb_match_data p50 {np.median(lat[30:]) * 1e6:6.1f} us")
    print(f"### {name}")
    for l in out:
        print("   ", l)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        run(sys.argv[2])
        return
    from hector_slam_b200 import build

    for k, (name, flags) in enumerate(VARIANTS.items()):
        lib = os.path.join(build.LIBDIR, f"libhsb200_var{k}.so")
        build.build_cuda(force=False, extra=flags, out=lib)
        subprocess.run([sys.executable, __file__, "--run", name + "  [" + " ".join(flags) + "]"],
                       env=dict(os.environ, HSB_LIB_PATH=lib))


if __name__ == "__main__":
    main()
