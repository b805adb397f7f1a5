"""Summarise an ncu report (.ncu-rep) as markdown: per captured launch, the metrics the roofline
discussion needs.  Usage: python scripts/ncu_summary.py report.ncu-rep [out.md]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("launch__occupancy_limit_blocks", "occ. limit: blocks"),
    ("launch__occupancy_limit_registers", "occ. limit: registers"),
    ("launch__occupancy_limit_shared_mem", "occ. limit: smem"),
    ("launch__occupancy_limit_warps", "occ. limit: warps"),
    ("launch__occupancy_limit_barriers", "occ. limit: barriers"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (per SM)"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible / scheduler"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe FP64 %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "pipe TEX %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1TEX throughput %"),
    ("l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed", "L1TEX tex write-back %"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "L1TEX lsu write-back %"),
    ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1TEX tex wavefronts %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("l1tex__t_bytes.sum", "L1TEX bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram__bytes_read.sum"),
    ("dram__bytes_write.sum", "dram__bytes_write.sum"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{rep.split('/')[-1]}`", "",
             "(`ncu --set full --clock-control none --import-source on`; one column per captured launch)", ""]
    names = [r[idx["Kernel Name"]] for r in data]
    lines.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    lines.append("|---|---|" + "---|" * len(data))
    lines.append("| kernel | | " + " | ".join(n.split("(")[0].replace("void ", "") for n in names) + " |")
    for key, label in WANT:
        if key not in idx:
            continue
        vals = []
        for r in data:
            v = r[idx[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.4e}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {label} (`{key}`) | {units[idx[key]]} | " + " | ".join(vals) + " |")
    if "--traffic" in sys.argv:
        # per-launch DRAM traffic of the last captured launch, for bench.py's roofline.traffic
        import json

        def val(key, r):
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        r = data[-1]

        def opt(key, scale=1.0):
            try:
                return val(key, r) * scale
            except (KeyError, ValueError):
                return None

        out = {"dram_read_bytes": val("dram__bytes_read.sum", r), "dram_write_bytes": val("dram__bytes_write.sum", r),
               "lts_bytes": opt("lts__t_bytes.sum"), "l1tex_bytes": opt("l1tex__t_bytes.sum"),
               "l2_hit_rate": opt("lts__t_sector_hit_rate.pct", 0.01), "l1_hit_rate": opt("l1tex__t_sector_hit_rate.pct", 0.01),
               "tex_writeback_frac": opt("l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed", 0.01),
               "issue_frac": opt("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.01),
               "achieved_occupancy": opt("sm__warps_active.avg.pct_of_peak_sustained_active", 0.01),
               "duration_us_under_ncu": opt("gpu__time_duration.sum"),
               "kernel": r[idx["Kernel Name"]], "source": "ncu --set full capture " + rep.split("/")[-1]}
        open(sys.argv[sys.argv.index("--traffic") + 1], "w").write(json.dumps(out, indent=1) + "\n")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2 and not sys.argv[2].startswith("--"):
        open(sys.argv[2], "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
