"""e2e pipeline probe: host-buffer batch calls with different chunk sizes (diagnostic)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hector_slam_b200 import capi, synth

B = 4096
rep = capi.MapRepB200(bench.RES, bench.MAP_SIZE, levels=3, update_factor_free=0.4, update_factor_occupied=0.9)
world, poses, pts, offs, hints = bench.make_workload(0, B)
bench.build_map_on_gpu(rep, world)
rep.set_scan_format(**synth.SCAN_FORMAT)
from hector_slam_b200 import parallel
if len(sys.argv) > 1 and sys.argv[1] == "bind":
    print("affinity before:", len(os.sched_getaffinity(0)), "bound ->", parallel.bind_process_to_gpu_numa_node(0) is not None, len(os.sched_getaffinity(0)))
h_ranges = parallel.pinned_copy(np.ascontiguousarray(bench.make_workload.ranges))
h_pts = parallel.pinned_copy(pts)
h_hints = parallel.pinned_copy(hints)
h_offs_pin = parallel.pinned_copy(offs)
h_poses = parallel.pinned_empty((B, 3))
h_cov = parallel.pinned_empty((B, 9))

def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

# raw copy bandwidth
d = torch.empty_like(h_pts, device="cuda")
ms = timeit(lambda: d.copy_(h_pts, non_blocking=True)); print(f"H2D {h_pts.numel()*4/1e6:.1f} MB: {ms:.3f} ms  {h_pts.numel()*4/ms/1e6:.1f} GB/s")
d2 = torch.empty_like(h_ranges, device="cuda")
ms = timeit(lambda: d2.copy_(h_ranges, non_blocking=True)); print(f"H2D {h_ranges.numel()*4/1e6:.1f} MB: {ms:.3f} ms  {h_ranges.numel()*4/ms/1e6:.1f} GB/s")
for chunk in (0, 4096, 1024):
    rep.set_tuning(chunk=chunk)
    a = timeit(lambda: rep.match_batch_ranges(h_hints, h_ranges, want_cov=True, out_poses=h_poses, out_cov=h_cov))
    b = timeit(lambda: rep.match_batch(h_hints, h_pts, h_offs_pin, want_cov=True, out_poses=h_poses, out_cov=h_cov))
    c = timeit(lambda: rep.match_batch(h_hints, h_pts, offs, want_cov=True, out_poses=h_poses, out_cov=h_cov))
    print(f"chunk {chunk}: ranges {a:.3f} ms ({B/a/1e3:.2f} M/s)  endpoints(pinned offs) {b:.3f} ms ({B/b/1e3:.2f} M/s)  endpoints(pageable offs) {c:.3f} ms")
