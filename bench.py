#!/usr/bin/env python
"""bench.py — scan-matches/sec of the B200 scan matcher (BASELINE.json metric).

Workload (BASELINE.json configs[1]): batches of 4096 independent 1081-point synthetic scans matched
against a frozen 3-level 2048^2 map (0.05 m), full MapRepMultiMap::matchData each (4+4+6 = 14
H/dTr evaluations).  A "step" = one batch.  With N GPUs every rank matches its own 4096 scans per
step against its replica of the map (weak scaling; the map is built on rank 0 and replicated with
one NCCL broadcast, no collective on the data path).

  value     whole-job matches/s with inputs resident in HBM (one kernel launch per step per rank)
  e2e       the same through the host-buffer C-ABI (hsb_match_batch_ranges_submit / _wait): pinned host sensor
            ranges + hints copied to the device and poses + covariances copied back inside the timed region, every
            step; two staging sets so that step k+1's copies overlap step k's tail
  e2e_endpoints / e2e_cloud   the same for the two other wire formats (DataContainer endpoints — what
            MapRepresentationInterface carries — and the node's default point-cloud input)
  roofline  §8(d)'s algorithmic-bytes figure for the match kernel next to what ncu says binds it (L1TEX texture
            write-back), its L2 and DRAM traffic; roofline_k2 the same for the map writer
  cpu_baseline  the reference's own CPU matcher (oracle/_ref, else the C port) on this box's cores
  extras    the other BASELINE configs measured, not only tested: config3 (fused SLAM step on the 4096^2 map),
            config4 (65 536 pose hypotheses, strong-scaled over the ranks, likelihood + arg-max all-gather),
            config5 (8192^2 map: single-GPU kernel throughput at N = 1, replay with NCCL dirty-tile broadcast at N > 1)

`--impl reference` times only the CPU arm and prints it as the main line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAP_SIZE = 2048
LEVELS = 3
RES = 0.05
BATCH = 4096
N_PTS = 1081
EVALS = 14
BYTES_PER_MATCH = N_PTS * 24 * EVALS  # 363 216 B
METRIC = "scan-matches/sec (1081-pt scans, 3-level map)"
WORKLOAD = "batch of 4096 independent 1081-pt synthetic scans, 3-level 2048^2 map, full matchData"


# ---------------------------------------------------------------------------------------------
def make_workload(seed: int, batch: int, map_size: int = MAP_SIZE):
    """Seeded scans + hints for one rank."""
    from hector_slam_b200 import synth

    world = synth.World.for_map_size(map_size)
    rng = np.random.default_rng(1000 + seed)
    poses = world.sample_free_poses(batch, rng)
    ranges = synth.make_range_batch(world, poses, noise_seed=7 + seed)
    pts = np.concatenate([synth.ranges_to_points(ranges[b], 1.0 / RES) for b in range(batch)])
    offs = (np.arange(batch + 1) * N_PTS).astype(np.int32)
    hints = synth.perturb_hints(poses, seed=1 + seed, dxy=0.1, dpsi=0.05)
    assert pts.shape[0] == batch * N_PTS
    make_workload.ranges = ranges
    return world, poses, np.ascontiguousarray(pts, dtype=np.float32), offs, hints


def build_map_on_gpu(rep, world):
    """Mapping with known poses through the product path (matchData fills the coarse containers,
    updateByScan writes every level) — same recipe as oracle.pyoracle.build_map_known_poses."""
    from hector_slam_b200 import synth

    rng = np.random.default_rng(11)
    for p in world.mapping_poses():
        scan = synth.make_scan(world, p, rng)
        p32 = p.astype(np.float32)
        rep.matchData(p32, scan)
        rep.updateByScan(scan, p32)
        rep.onMapUpdated()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def committed_ncu(name: str):
    """Per-launch figures of a committed `ncu --set full` capture (profiles/<name>.json, written by
    scripts/ncu_summary.py --traffic): DRAM and L2 bytes, utilisation of the binding unit.  {} if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", name + ".json")) as f:
            return json.load(f)
    except Exception:
        return {}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(pts, offs, hints, planes, steps: int, warmup: int, max_seconds: float = 20.0,
                      map_size: int = MAP_SIZE):
    """Time the reference's CPU matcher (oracle/_ref when present, else the C port) with all host
    threads on a bounded sample of the workload. Returns (matches_per_s, info)."""
    from oracle import pyoracle

    if not pyoracle.available("port"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    kind = "reference" if pyoracle.available("reference") else "port"
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 256))
    orc = pyoracle.Oracle(kind, RES, map_size, LEVELS)
    orc.set_update_factors(0.4, 0.9)
    for l in range(LEVELS):
        orc.set_logodds(l, planes[l])
    B = hints.shape[0]
    # calibrate on a small single-thread run, then size a step to ~max_seconds/(steps+warmup)
    n0 = min(B, 64)
    _, _, s0 = orc.match_batch(hints[:n0], pts[: offs[n0]], offs[: n0 + 1], nthreads=1, want_cov=True)
    per_match = s0 / n0
    budget = max_seconds / max(1, steps + warmup)
    n_step = int(min(B, max(threads * 8, budget * threads / per_match)))
    sub_pts, sub_offs, sub_hints = pts[: offs[n_step]], offs[: n_step + 1], hints[:n_step]
    secs = []
    for it in range(warmup + steps):
        _, _, s = orc.match_batch(sub_hints, sub_pts, sub_offs, nthreads=threads)
        if it >= warmup:
            secs.append(s)
    orc.close()
    total = float(np.sum(secs))
    value = n_step * len(secs) / total
    info = {"value": value, "unit": "scan-matches/s", "cores": threads, "kind": kind,
            "sample": f"{len(secs)} x {n_step} of the {B} scans, {threads} threads (one private matcher + map copy "
                      f"each, warm probability cache), single-thread {1.0 / per_match:.0f} matches/s",
            "host_cpus": cores, "ms_per_step": 1e3 * total / len(secs), "n_step": n_step}
    return value, info


def slam_step_latency(rep, scans, hints, planes, map_size, n_gpu: int = 200, n_cpu: int = 40, with_cpu: bool = True):
    """Config 3's step: one scan at a time through HectorSlamProcessor::update — match, gate, map
    write — as the fused hsb_slam_update call with pageable host buffers, thresholds 0 so that every step writes the
    map; next to it the compiled reference doing the same step on one host core (how the reference runs it), and the
    K2 roofline (8 B per unique cell written, cells counted by plane difference on one step).  Mutates rep's map."""
    import ctypes as C

    nscan = len(scans)
    hp = [np.ascontiguousarray(hints[i], np.float32) for i in range(nscan)]
    rep.setMapUpdateMinDistDiff(0.0)
    rep.setMapUpdateMinAngleDiff(0.0)
    o_pose, o_cov, upd = np.zeros(3, np.float32), np.zeros(9, np.float32), C.c_int(0)
    lib, hnd = rep.lib, rep.h
    lat = []
    for i in range(n_gpu + 20):
        k = i % nscan
        t0 = time.perf_counter()
        st = lib.hsb_slam_update(hnd, hp[k].ctypes.data, scans[k].ctypes.data, scans[k].shape[0], None, 0,
                                 o_pose.ctypes.data, o_cov.ctypes.data, C.addressof(upd))
        t1 = time.perf_counter()
        if st != 0:
            raise RuntimeError("hsb_slam_update failed")
        if i >= 20:
            lat.append(t1 - t0)
    lat = np.sort(np.asarray(lat)) * 1e6
    out = {"call": "hsb_slam_update (match + gate + updateByScan + onMapUpdated, one scan, host buffers)",
           "map": f"{map_size}^2 x {LEVELS} levels", "gpu_us_p50": float(lat[len(lat) // 2]),
           "gpu_us_p99": float(lat[int(0.99 * (len(lat) - 1))]), "scans_per_s": float(1e6 / lat.mean()),
           "steps": n_gpu, "budget_us_at_40hz": 25000.0}
    # the same step returning as soon as the pose is known (hsb_slam_update_nowait: the match kernel publishes pose +
    # gate decision to mapped host memory, the host polls; the map write continues on the stream).  "isolated": the
    # stream is idle when the call starts, as at 40 Hz; "back_to_back": sustained rate, each call queues behind the
    # previous step's map write.
    lat_iso, lat_b2b = [], []
    for i in range(n_gpu + 20):
        k = i % nscan
        t0 = time.perf_counter()
        st = lib.hsb_slam_update_nowait(hnd, hp[k].ctypes.data, scans[k].ctypes.data, scans[k].shape[0], None, 0,
                                        o_pose.ctypes.data, o_cov.ctypes.data, C.addressof(upd))
        t1 = time.perf_counter()
        lib.hsb_on_map_updated(hnd)
        if st != 0:
            raise RuntimeError("hsb_slam_update_nowait failed")
        if i >= 20:
            lat_iso.append(t1 - t0)
    t0 = time.perf_counter()
    for i in range(n_gpu):
        k = i % nscan
        lib.hsb_slam_update_nowait(hnd, hp[k].ctypes.data, scans[k].ctypes.data, scans[k].shape[0], None, 0,
                                   o_pose.ctypes.data, o_cov.ctypes.data, C.addressof(upd))
    lib.hsb_on_map_updated(hnd)
    b2b = (time.perf_counter() - t0) / n_gpu
    lat_iso = np.sort(np.asarray(lat_iso)) * 1e6
    out["pose_latency_us_p50"] = float(lat_iso[len(lat_iso) // 2])
    out["pose_latency_us_p99"] = float(lat_iso[int(0.99 * (len(lat_iso) - 1))])
    out["pose_latency_call"] = "hsb_slam_update_nowait, stream idle at call time (40 Hz use); map write completes in the background"
    out["sustained_scans_per_s_nowait"] = float(1.0 / b2b)
    # K2 alone: device time of mark + apply (CUDA events on the handle's stream) and the cells it wrote
    rep.set_tuning(time_update=1)
    before = [rep.download_level(l) for l in range(LEVELS)]
    k2_ms = []
    for i in range(12):
        k = i % nscan
        p, _ = rep.matchData(hp[k], scans[k])
        if i == 0:
            before = [rep.download_level(l) for l in range(LEVELS)]
        rep.updateByScan(scans[k], p)
        if i == 0:
            cells = [int((rep.download_level(l) != before[l]).sum()) for l in range(LEVELS)]
        if i >= 2:
            k2_ms.append(rep.last_update_device_ms())
    rep.set_tuning(time_update=0)
    k2 = float(np.median(k2_ms))
    peak, peak_src = measured_peak_gbs()
    alg = 8.0 * sum(cells)
    out["roofline_k2"] = {"kernel": "hsb::update_mark_kernel + hsb::update_apply_kernel (one scan, all levels)",
                          "bound": "latency", "achieved": alg / (k2 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": alg / (k2 * 1e-3) / 1e9 / peak, "kernel_ms": k2, "unique_cells_per_level": cells,
                          "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                          "note": "8 B per unique cell written (SURVEY.md §8d); two ~10 us launches sized to one scan: "
                                  "bound by launch + dependent-atomic latency, not by bandwidth"}
    if with_cpu:
        from oracle import pyoracle

        kind = "reference" if pyoracle.available("reference") else "port"
        orc = pyoracle.Oracle(kind, RES, map_size, LEVELS)
        orc.set_update_factors(0.4, 0.9)
        orc.set_map_update_thresholds(0.0, 0.0)
        for l in range(LEVELS):
            orc.set_logodds(l, planes[l])
        cl = []
        for i in range(n_cpu + 5):
            k = i % nscan
            t0 = time.perf_counter()
            orc.update(scans[k], hp[k])
            t1 = time.perf_counter()
            if i >= 5:
                cl.append(t1 - t0)
        orc.close()
        out["cpu_reference_us_p50"] = float(np.median(cl) * 1e6)
        out["cpu_kind"] = kind
        out["cpu_cores"] = 1
    return out


# ---------------------------------------------------------------------------------------------
def time_device_steps(torch, dist, world_size, dev, fn, steps, warmup):
    """W warm-ups, then K steps bracketed by barrier + synchronize, CUDA events on torch's current stream (the stream
    the kernels are launched on), max over ranks of the first-start -> last-stop span.  -> (span_ms_max, per-step ms)"""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()                      # ONE event pair around the K steps: per-step events would sit between the launches
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    span_ms = e0.elapsed_time(e1)
    step_ms = [span_ms / steps] * steps
    t = torch.tensor([span_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), step_ms


TLD4_CYCLES_PER_WARP_GATHER = 15.96   # measured: scripts/ubench/tex_rate.cu, profiles/r02_tex_rate.log (scan-like footprints)


def tex_pipe_ceiling(matches_per_s_per_gpu, clocks):
    """The roof that binds K1: an sm_100a SM returns one warp-level TLD4 (4 fp32 x 32 lanes) every ~16 cycles, whatever
    the footprints' spread; no other path delivers the four neighbours cheaper (profiles/r02_tex_rate.log).  A match
    issues EVALS x ceil(N_PTS / 32) warp-gathers."""
    import torch

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    wg = EVALS * ((N_PTS + 31) // 32)
    ceiling = sms * mhz * 1e6 / (wg * TLD4_CYCLES_PER_WARP_GATHER)
    return {"cycles_per_warp_gather": TLD4_CYCLES_PER_WARP_GATHER, "warp_gathers_per_match": wg, "sm_mhz": mhz,
            "ceiling": ceiling, "unit": "scan-matches/s per GPU", "frac": matches_per_s_per_gpu / ceiling,
            "source": "profiles/r02_tex_rate.log (scripts/ubench/tex_rate.cu)"}


def pcie_bound(h2d_bytes, gbs, B, world_size, e2e_value):
    """What the host link alone allows: the step's input bytes at the H2D rate measured in this process (rank 0's link;
    every rank has its own), and the share of it the pipelined call reaches."""
    if not gbs:
        return None
    floor_ms = h2d_bytes / (gbs * 1e9) * 1e3
    bound = B * world_size / (floor_ms * 1e-3)
    return {"h2d_floor_ms_per_step": floor_ms, "value_if_only_the_copy_counted": bound, "frac": e2e_value / bound}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--gather", default="auto", choices=["auto", "ldg", "tex"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / 5 figures")
    ap.add_argument("--only", default="", help="run only this part (for ncu captures): value | config5 | slam")
    ap.add_argument("--nbuf", type=int, default=8, help="distinct input batches cycled through (inputs > L2)")
    ap.add_argument("--tune", default="", help="k=v,k=v passed to hsb_set_tuning (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import pyoracle  # noqa: F401  (bench.py's reference arm may execute oracle/)

        world, poses, pts, offs, hints = make_workload(0, args.batch)
        # the same map recipe on the CPU side
        kind = "reference" if pyoracle.available("reference") else "port"
        if not pyoracle.available("port"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
        orc = pyoracle.Oracle(kind, RES, MAP_SIZE, LEVELS)
        orc.set_update_factors(0.4, 0.9)
        pyoracle.build_map_known_poses(orc, world)
        planes = [orc.get_logodds(l) for l in range(LEVELS)]
        orc.close()
        value, info = cpu_reference_run(pts, offs, hints, planes, args.steps, args.warmup)
        line = {"metric": METRIC, "value": value, "unit": "scan-matches/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD, "map": f"{MAP_SIZE}^2 x {LEVELS} levels @ {RES} m",
                           "evaluations_per_match": EVALS, "step": f"{info['n_step']} scans on the host cores"},
                "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": value, "unit": "scan-matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from hector_slam_b200 import capi, parallel, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left exactly as the launcher set it (the driver reads rank counts from NCCL's INFO lines);
        # rank 0's JSON line is the LAST line this process prints
        dist.init_process_group("nccl", device_id=dev)

    gather = {"auto": capi.GATHER_AUTO, "ldg": capi.GATHER_LDG, "tex": capi.GATHER_TEX}[args.gather]
    tune = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in args.tune.split(",") if kv)

    def new_rep(size):
        r = capi.MapRepB200(RES, size, levels=LEVELS, device=local_rank, update_factor_free=0.4,
                            update_factor_occupied=0.9, gather_mode=gather)
        if tune:
            r.set_tuning(**tune)
        return r

    if args.only in ("config5", "slam"):   # one part alone (ncu captures)
        part = (extra_config5(torch, dist, capi, synth, parallel, new_rep, rank, world_size, dev, args) if args.only == "config5"
                else extra_config3(capi, synth, new_rep, args))
        if rank == 0:
            print(json.dumps(part))
        return

    rep = new_rep(MAP_SIZE)
    world, poses, pts, offs, hints = make_workload(rank, args.batch)
    ranges = np.ascontiguousarray(make_workload.ranges)
    B = args.batch
    # endpoints exactly as the node's converter produces them from these ranges (make_workload's come from numpy's cos / sin,
    # which differ from glibc's cosf / sinf in the last bit): the device-resident path and the host paths then see
    # identical scans and `e2e.max_abs_diff_vs_device_path` is a true consistency check
    rep.set_scan_format(**synth.SCAN_FORMAT)
    conv = [rep.scan_to_points(ranges[b]) for b in range(B)]
    if all(c.shape[0] == N_PTS for c in conv):
        pts = np.ascontiguousarray(np.concatenate(conv), dtype=np.float32)

    # ---- map: built once on rank 0 through the product path, replicated with one NCCL broadcast
    if rank == 0:
        build_map_on_gpu(rep, world)
    planes_dev = parallel.replicate_map(rep, dev, src=0)
    planes_host = [p.cpu().numpy() for p in planes_dev] if rank == 0 else None

    # ---- device-resident inputs: nbuf distinct copies so that a step's inputs are not L2-warm
    nbuf = max(1, args.nbuf)
    d_pts = [torch.from_numpy(pts).to(dev).clone() for _ in range(nbuf)]
    d_hints = [torch.from_numpy(hints).to(dev).clone() for _ in range(nbuf)]
    d_offs = torch.from_numpy(offs).to(dev)
    d_poses = torch.empty((B, 3), dtype=torch.float32, device=dev)
    d_cov = torch.empty((B, 9), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(i):
        k = i % nbuf
        rep.match_batch_device(B, d_hints[k].data_ptr(), d_pts[k].data_ptr(), d_offs.data_ptr(), 0, N_PTS,
                               d_poses.data_ptr(), d_cov.data_ptr(), stream)

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = rep.launch_count
    t_wall0 = time.perf_counter()
    span_ms_max, step_ms = time_device_steps(torch, dist, world_size, dev, step_device, args.steps, args.warmup)
    t_wall1 = time.perf_counter()
    launches = (rep.launch_count - launches0) * args.steps // (args.steps + args.warmup)
    shape = rep.last_launch_shape()
    value = world_size * B * args.steps / (span_ms_max * 1e-3)
    kernel_ms = float(np.sum(step_ms)) / args.steps
    if args.only == "value":
        if rank == 0:
            print(json.dumps({"only": "value", "value": value, "kernel_ms": kernel_ms, "shape": shape}))
        rep.close()
        return

    # ---- timed region 2: end to end through the host-buffer C-ABI, page-locked host memory -----------------
    # Host buffers come from hsb_alloc_pinned (cudaHostAlloc by this thread, bound to the GPU's NUMA node) — not from
    # torch's caching host allocator, whose blocks copied at anything between 12 and 55 GB/s on these hosts.
    rep.set_scan_format(**synth.SCAN_FORMAT)
    T_laser = synth.laser_transform()
    rep.set_cloud_format(T_laser, **synth.CLOUD_FORMAT)
    clouds = [synth.ranges_to_cloud(ranges[b]) for b in range(B)]
    c_offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
    cloud_all = np.ascontiguousarray(np.concatenate(clouds), dtype=np.float32)
    prev_affinity = parallel.bind_process_to_gpu_numa_node(local_rank)
    NSET = 2
    h_ranges = [capi.pinned_copy(ranges) for _ in range(NSET)]
    h_pts = [capi.pinned_copy(pts) for _ in range(NSET)]
    h_cloud = [capi.pinned_copy(cloud_all) for _ in range(NSET)]
    h_hints = [capi.pinned_copy(hints) for _ in range(NSET)]
    h_poses = [capi.PinnedArray((B, 3)) for _ in range(NSET)]
    h_cov = [capi.PinnedArray((B, 9)) for _ in range(NSET)]
    h_origo = [capi.PinnedArray((B, 2)) for _ in range(NSET)]

    def raw_h2d_gbs(parr):
        return rep.measure_h2d_gbs(parr.ptr, parr.array.nbytes)

    def submit_ranges(i):
        k = i % NSET
        return rep.match_batch_ranges_submit(h_hints[k].array, h_ranges[k].array, h_poses[k].array, h_cov[k].array)

    def submit_xy(i):
        k = i % NSET
        return rep.match_batch_submit(h_hints[k].array, h_pts[k].array, offs, h_poses[k].array, h_cov[k].array)

    def submit_cloud(i):
        k = i % NSET
        return rep.match_batch_cloud_submit(h_hints[k].array, h_cloud[k].array, c_offs, h_poses[k].array, h_cov[k].array,
                                            h_origo[k].array)

    def time_host(submit, pipelined=True):
        """K steps through the public host-buffer API, wall clock bracketed by synchronize + barrier on both sides, max
        over ranks.  pipelined: step k+1 is submitted before step k is waited for (two staging sets)."""
        for i in range(max(3, args.warmup)):
            rep.match_batch_wait(submit(i))
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        l0 = rep.launch_count
        t0 = time.perf_counter()
        if pipelined:
            pending = submit(0)
            for i in range(1, args.steps):
                nxt = submit(i)
                rep.match_batch_wait(pending)
                pending = nxt
            rep.match_batch_wait(pending)
        else:
            for i in range(args.steps):
                rep.match_batch_wait(submit(i))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.barrier()
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world_size * B * args.steps / float(te.item()), rep.launch_count - l0

    e2e_xy_value, _ = time_host(submit_xy)
    e2e_cloud_value, _ = time_host(submit_cloud)
    e2e_blocking_value, _ = time_host(submit_ranges, pipelined=False)
    e2e_value, e2e_launches = time_host(submit_ranges)
    h2d_gbs = {"ranges": raw_h2d_gbs(h_ranges[0]), "endpoints": raw_h2d_gbs(h_pts[0])} if rank == 0 else None
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)
    clocks = sampler.stop() if rank == 0 else None

    # parity spot check of the e2e result against the device-resident path (same inputs)
    rep.match_batch_device(B, d_hints[0].data_ptr(), d_pts[0].data_ptr(), d_offs.data_ptr(), 0, N_PTS, d_poses.data_ptr(),
                           d_cov.data_ptr(), stream)
    torch.cuda.synchronize()
    rep.match_batch_wait(submit_ranges(0))
    same = float(np.abs(d_poses.cpu().numpy() - h_poses[0].array).max())

    # ---- extras: the other BASELINE configs, measured ----------------------------------------------------
    extras = {}
    if not args.no_extras:
        extras.update(extra_config4(torch, dist, capi, synth, parallel, new_rep, rank, world_size, dev, args))
        extras.update(extra_config5(torch, dist, capi, synth, parallel, new_rep, rank, world_size, dev, args))

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ncu = committed_ncu("match_kernel_traffic")
        traffic = (ncu.get("dram_read_bytes", 0.0) + ncu.get("dram_write_bytes", 0.0)) if ncu else None
        achieved = BYTES_PER_MATCH * B / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "scan-matches/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": span_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "map": f"{MAP_SIZE}^2 x {LEVELS} levels @ {RES} m", "batch_per_gpu": B,
                       "evaluations_per_match": EVALS, "gather": {1: "ldg", 2: "tex"}[rep.gather_mode],
                       "launch_shape": shape,
                       "cache": f"inputs larger than L2: {nbuf} distinct input batches cycled "
                                f"({nbuf * pts.nbytes / 1e6:.0f} MB of endpoints); the frozen map is reused by design",
                       "map_replication": "rank 0 builds, one NCCL broadcast" if world_size > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "scan-matches/s",
                    "h2d_bytes_per_step": int(ranges.nbytes + hints.nbytes),
                    "d2h_bytes_per_step": int(B * 12 + B * 36), "launches": int(e2e_launches),
                    "call": "hsb_match_batch_ranges_submit / hsb_match_batch_wait: pinned host sensor ranges (4 B/beam) + "
                            "hints in, poses + covariances out; scan->endpoint conversion fused into the match kernel; "
                            "step k+1 submitted before step k is waited for (two staging sets)",
                    "blocking_call_value": e2e_blocking_value,
                    "host_buffers": "hsb_alloc_pinned (cudaHostAlloc), process bound to the GPU's NUMA node"
                                    if prev_affinity is not None else "hsb_alloc_pinned (NUMA node of the GPU unknown)",
                    "raw_h2d_gbs": h2d_gbs,
                    "pcie_bound": pcie_bound(int(ranges.nbytes + hints.nbytes), h2d_gbs["ranges"], B, world_size, e2e_value),
                    "max_abs_diff_vs_device_path": same},
            "e2e_endpoints": {"value": e2e_xy_value, "unit": "scan-matches/s",
                              "h2d_bytes_per_step": int(pts.nbytes + hints.nbytes + offs.nbytes),
                              "d2h_bytes_per_step": int(B * 12 + B * 36),
                              "pcie_bound": pcie_bound(int(pts.nbytes + hints.nbytes + offs.nbytes), h2d_gbs["endpoints"], B,
                                                       world_size, e2e_xy_value),
                              "call": "hsb_match_batch_submit / _wait: DataContainer endpoints (8 B each) in — the format "
                                      "MapRepresentationInterface::matchData carries"},
            "e2e_cloud": {"value": e2e_cloud_value, "unit": "scan-matches/s",
                          "h2d_bytes_per_step": int(cloud_all.nbytes + hints.nbytes + c_offs.nbytes),
                          "d2h_bytes_per_step": int(B * 12 + B * 36 + B * 8),
                          "pcie_bound": pcie_bound(int(cloud_all.nbytes + hints.nbytes + c_offs.nbytes), h2d_gbs["endpoints"],
                                                   B, world_size, e2e_cloud_value),
                          "call": "hsb_match_batch_cloud_submit / _wait: sensor_msgs/PointCloud points (12 B each) in, "
                                  "rosPointCloudToDataContainer fused into the match kernel (the node's default path)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "l1tex", "bound_unit": "l1tex__tex_writeback (texture write-back path of the SM)",
                         "bound_unit_frac": ncu.get("tex_writeback_frac"), "issue_frac": ncu.get("issue_frac"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "lts_bytes": ncu.get("lts_bytes"),
                         "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "traffic_source": ncu.get("source"),
                         "note": "achieved/peak/frac are SURVEY.md §8(d)'s algorithmic bytes (24 B per endpoint-evaluation) "
                                 "over kernel time against the measured HBM copy peak; the gathers are served by L1/L2 "
                                 "(DRAM traffic = the endpoints, once: dram_frac), so HBM is not the roof of this "
                                 "configuration and frac > 1 is expected — the unit that binds is the L1TEX texture "
                                 "write-back path (bound_unit_frac, ncu) together with instruction issue (issue_frac)",
                         "tex_pipe": tex_pipe_ceiling(value / world_size, clocks),
                         "peak_source": peak_src, "kernel": "hsb::match_kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": BYTES_PER_MATCH * B},
            "clocks": clocks,
            "wall_ms_per_step": 1e3 * (t_wall1 - t_wall0) / (args.steps + args.warmup),
        }
        line.update(extras)
        if world_size == 1 and not args.no_extras:
            # (latencies of single calls: taken BEFORE the 128-thread CPU baseline below, whose aftermath on the host
            # cores otherwise adds ~3 us to every call)
            line.update(extra_config3(capi, synth, new_rep, args))
        if not args.no_cpu_baseline and world_size == 1:
            _, info = cpu_reference_run(pts, offs, hints, planes_host, steps=3, warmup=1, max_seconds=16.0)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
    rep.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:   # after NCCL's teardown, so that with NCCL_DEBUG=INFO the JSON line is still the last line on stdout
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def extra_config3(capi, synth, new_rep, args):
    """BASELINE config 3 at its stated size: one scan at a time, 3-level 4096^2 map, match + updateByScan per scan."""
    size = 4096
    rep = new_rep(size)
    world = synth.World.for_map_size(size)
    build_map_on_gpu(rep, world)
    planes = [rep.download_level(l) for l in range(LEVELS)]
    rng = np.random.default_rng(31)
    poses = world.sample_free_poses(64, rng)
    scans = [np.ascontiguousarray(synth.make_scan(world, p, rng)) for p in poses]
    hints = synth.perturb_hints(poses, seed=32, dxy=0.05, dpsi=0.02)
    out = slam_step_latency(rep, scans, hints, planes, size, with_cpu=not args.no_cpu_baseline)
    # the node's DEFAULT scanCallback branch as one call (hsb_slam_update_cloud): laser-frame point cloud + tf in,
    # rosPointCloudToDataContainer fused into the match kernel's staging, gate, map write; nowait = pose latency
    rep.set_cloud_format(synth.laser_transform(), **synth.CLOUD_FORMAT)
    rep.setMapUpdateMinDistDiff(0.0)
    rep.setMapUpdateMinAngleDiff(0.0)
    rr = synth.make_range_batch(world, poses, noise_seed=33)
    clouds = [synth.ranges_to_cloud(rr[k]) for k in range(len(poses))]
    for nowait, key in ((False, "cloud_step_us_p50"), (True, "cloud_step_pose_latency_us_p50")):
        lat = []
        for i in range(220):
            k = i % len(clouds)
            t0 = time.perf_counter()
            rep.slam_update_cloud(hints[k], clouds[k], nowait=nowait)
            t1 = time.perf_counter()
            if nowait:
                rep.onMapUpdated()
            if i >= 20:
                lat.append(t1 - t0)
        out[key] = float(np.median(lat) * 1e6)
    out["cloud_step_call"] = ("hsb_slam_update_cloud through the ctypes wrapper (adds ~5 us of Python argument handling "
                              "to the C call): 1081 Point32 + laser tf in, pose out")
    rep.close()
    k2 = out.pop("roofline_k2")
    return {"config3_slam_step": out, "roofline_k2": k2}


def extra_config4(torch, dist, capi, synth, parallel, new_rep, rank, world_size, dev, args):
    """BASELINE config 4: Monte-Carlo relocalisation — 65 536 pose hypotheses x ONE 1081-pt scan against a fixed 4096^2
    map, STRONG-scaled: every rank matches 65 536 / N hypotheses (shared-scan launch), scores the results with the
    likelihood kernel and the best hypothesis is found with one all-gather of (score, pose).  hypotheses/s whole-job."""
    size, H = 4096, 65536
    rep = new_rep(size)
    world = synth.World.for_map_size(size)
    if rank == 0:
        build_map_on_gpu(rep, world)
    parallel.replicate_map(rep, dev, src=0)
    rng = np.random.default_rng(2)
    truth = world.sample_free_poses(1, rng, margin=1.0)[0]
    scan = np.ascontiguousarray(synth.make_scan(world, truth, np.random.default_rng(7)))
    hyp = np.tile(truth, (H, 1))
    hyp[:, 0] += rng.uniform(-2.0, 2.0, H)
    hyp[:, 1] += rng.uniform(-2.0, 2.0, H)
    hyp[:, 2] += rng.uniform(-0.5, 0.5, H)
    hyp[:1024, :2] = truth[:2] + rng.uniform(-0.2, 0.2, (1024, 2))
    hyp[:1024, 2] = truth[2] + rng.uniform(-0.1, 0.1, 1024)
    lo, hi = parallel.shard_range(H, rank, world_size)
    mine = np.ascontiguousarray(hyp[lo:hi], dtype=np.float32)
    n = hi - lo
    d_scan = torch.from_numpy(scan).to(dev)
    d_hyp = torch.from_numpy(mine).to(dev)
    d_out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    best = {}

    mine_best = torch.empty(4, dtype=torch.float32, device=dev)

    def step(i):
        rep.match_batch_device(n, d_hyp.data_ptr(), d_scan.data_ptr(), None, scan.shape[0], scan.shape[0], d_out.data_ptr(),
                               None, stream)
        if i < 0:
            return
        # score + arg-max across ranks (the exchange step of this config): likelihood of every matched pose on level 0
        # with the arg-max folded in (hsb_best_hypothesis_device: two launches), ONE all-gather of (score, pose) per
        # step and one 16-byte-per-rank read by the host
        rep.best_hypothesis_device(0, n, d_out.data_ptr(), d_scan.data_ptr(), None, scan.shape[0], mine_best.data_ptr(),
                                   None, stream)
        if world_size > 1:
            allb = torch.empty((world_size, 4), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allb, mine_best)
            allb = allb.cpu().numpy()
        else:
            allb = mine_best.cpu().numpy()[None]
        best["pose"] = allb[int(np.argmax(allb[:, 0])), 1:]

    steps, warm = 10, 2
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.barrier()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # match kernel alone (device events), for comparison
    span_ms, _ = time_device_steps(torch, dist, world_size, dev, lambda i: step(-1), 5, 2)
    err = float(np.abs(best["pose"][:2] - truth[:2]).max())
    rep.close()
    return {"config4_relocalisation": {
        "workload": "65 536 pose hypotheses x one 1081-pt scan, fixed 3-level 4096^2 map, sharded over the ranks",
        "scaling": "strong", "n_gpus": world_size, "hypotheses_per_gpu": n,
        "value": H * steps / float(te.item()), "unit": "hypotheses/s (match + likelihood score with fused arg-max on the device + all-gather; the host reads the winner every step)",
        "match_kernel_only": H * 5 / (span_ms * 1e-3), "ms_per_step": 1e3 * float(te.item()) / steps,
        "best_hypothesis_error_m": err, "collective": "all_gather of 4 floats per rank per step" if world_size > 1 else "none"}}


def extra_config5(torch, dist, capi, synth, parallel, new_rep, rank, world_size, dev, args):
    """BASELINE config 5: offline replay on the 3-level 8192^2 map (336 MB of probabilities: the one pyramid that does
    not fit L2).  N = 1: device-resident kernel throughput (the HBM / L2-miss case of the roofline).  N > 1: scans
    sharded over the ranks, rank 0 additionally writes the map with a scan every step and the dirty tiles are broadcast
    with NCCL inside the timed region; replicas are checked bit-identical to the owner afterwards."""
    size = 8192
    rep = new_rep(size)
    world = synth.World.for_map_size(size)
    if rank == 0:
        build_map_on_gpu(rep, world)
    parallel.replicate_map(rep, dev, src=0)
    B = 4096
    rng = np.random.default_rng(500 + rank)
    poses = world.sample_free_poses(B, rng)
    rngs = synth.make_range_batch(world, poses, noise_seed=70 + rank)
    pts = np.ascontiguousarray(np.concatenate([synth.ranges_to_points(rngs[b], 1.0 / RES) for b in range(B)]), dtype=np.float32)
    offs = (np.arange(B + 1) * N_PTS).astype(np.int32)
    hints = synth.perturb_hints(poses, seed=71 + rank, dxy=0.1, dpsi=0.05)
    nb = 4
    d_pts = [torch.from_numpy(pts).to(dev).clone() for _ in range(nb)]
    d_hints = torch.from_numpy(hints).to(dev)
    d_offs = torch.from_numpy(offs).to(dev)
    d_poses = torch.empty((B, 3), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def match(i):
        rep.match_batch_device(B, d_hints.data_ptr(), d_pts[i % nb].data_ptr(), d_offs.data_ptr(), 0, N_PTS, d_poses.data_ptr(),
                               None, stream)

    span_ms, step_ms = time_device_steps(torch, dist, world_size, dev, match, 10, 3)
    kernel_ms = float(np.mean(step_ms))
    peak, peak_src = measured_peak_gbs()
    ncu = committed_ncu("match_kernel_traffic_8192")
    traffic = (ncu.get("dram_read_bytes", 0.0) + ncu.get("dram_write_bytes", 0.0)) if ncu else None
    out = {"workload": "4096 independent 1081-pt scans per GPU per step, 3-level 8192^2 map (1.4 GB of planes, 336 MB of "
                       "probabilities: larger than L2), full matchData", "n_gpus": world_size,
           "value": world_size * B * 10 / (span_ms * 1e-3), "unit": "scan-matches/s", "kernel_ms": kernel_ms,
           "seconds_per_1M_scans_at_this_rate": 1e6 / (world_size * B * 10 / (span_ms * 1e-3)),
           "roofline": {"bound": "l1tex+l2", "achieved": BYTES_PER_MATCH * B / (kernel_ms * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": BYTES_PER_MATCH * B / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                        "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                        "lts_bytes": ncu.get("lts_bytes"), "l2_hit_rate": ncu.get("l2_hit_rate"),
                        "bound_unit_frac": ncu.get("tex_writeback_frac"), "traffic_source": ncu.get("source"),
                        "peak_source": peak_src}}
    if args.only == "config5":
        rep.close()
        return {"config5_replay_8192": out}
    if world_size > 1:
        # replay with map writes: every step all ranks match their shard; rank 0 then integrates one scan into the map
        # (hsb_slam_update) and ships the dirty tiles; the next step's matches see the new map everywhere
        wrng = np.random.default_rng(9)
        wposes = world.sample_free_poses(16, wrng)
        wscans = [np.ascontiguousarray(synth.make_scan(world, p, wrng)) for p in wposes]
        whints = wposes.astype(np.float32)
        rep.setMapUpdateMinDistDiff(0.0)
        rep.setMapUpdateMinAngleDiff(0.0)
        stats, bc_us, cells = {}, [], []
        tbuf = parallel.tile_buffer(dev)
        one_shot = [True]

        def replay_step(i):
            match(i)
            if rank == 0:
                torch.cuda.current_stream().synchronize()   # the owner's map write must not overtake its own match
                # nowait: back as soon as the pose is known; the tile pack below is stream-ordered behind the map write
                rep.slam_update(whints[i % 16], wscans[i % 16], nowait=True)
            t0 = time.perf_counter()
            if one_shot[0]:
                parallel.broadcast_dirty_tiles_async(rep, tbuf, src=0)
            else:
                parallel.broadcast_dirty_tiles(rep, dev, src=0, stats=stats)
            if i >= 3:
                bc_us.append((time.perf_counter() - t0) * 1e6)
                cells.append(stats.get("cells", 0))

        rp_steps, rp_warm = 10, 3

        def timed_replay():
            for i in range(rp_warm):
                replay_step(i)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for i in range(rp_steps):
                replay_step(rp_warm + i)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            dist.barrier()
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return te

        one_shot[0] = False
        te2 = timed_replay()                 # two-step protocol (sizes through the hosts)
        two_step_us, two_step_cells = float(np.median(bc_us)), float(np.median(cells))
        bc_us.clear()
        cells.clear()
        one_shot[0] = True
        te = timed_replay()                  # one-shot protocol (self-describing buffer)
        overflows = rep.replication_overflows()
        # replicas bit-identical to the owner: checksum of every plane
        sums = torch.stack([parallel.level_plane_tensor(rep, l, dev).double().sum() for l in range(LEVELS)])
        allsums = [torch.empty_like(sums) for _ in range(world_size)]
        dist.all_gather(allsums, sums)
        same = all(bool(torch.equal(allsums[0], s)) for s in allsums)
        out["replay_with_tile_broadcast"] = {
            "value": world_size * B * rp_steps / float(te.item()), "unit": "scan-matches/s (wall clock, map write + NCCL tile "
            "broadcast every step inside the timed region)", "steps": rp_steps, "ms_per_step": 1e3 * float(te.item()) / rp_steps,
            "protocol": "one-shot: device-side pack of all levels' dirty rectangles + their descriptions into a fixed "
                        f"{tbuf.numel() * 4 >> 20} MB buffer, ONE ncclBroadcast, device-side unpack; no host round trip",
            "host_us_per_broadcast_p50": float(np.median(bc_us)), "overflows": int(overflows),
            "replicas_bit_identical": bool(same),
            "two_step_protocol": {"value": world_size * B * rp_steps / float(te2.item()),
                                  "ms_per_step": 1e3 * float(te2.item()) / rp_steps, "broadcast_us_p50": two_step_us,
                                  "cells_per_broadcast_p50": two_step_cells,
                                  "bytes_per_broadcast_p50": two_step_cells * 4 + 16 * LEVELS,
                                  "collectives_per_step": "2 x ncclBroadcast (rectangles, then one packed buffer of all "
                                                          "levels' rows sized from them on every host)"}}
        assert same and overflows == 0, "replica planes differ from the owner's after tile broadcasts"
    rep.close()
    return {"config5_replay_8192": out}


if __name__ == "__main__":
    main()
