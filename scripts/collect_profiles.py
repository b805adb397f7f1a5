"""Copy what a GPU session (scripts/gpu_session.sh / gpu_session_ncu.sh) left in gpurun_out/ into profiles/ under
round-tagged names, and turn the ncu reports into markdown summaries + the traffic JSON bench.py reads.

  python scripts/collect_profiles.py <session tag> [<ncu tag>]
"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
ntag = sys.argv[2] if len(sys.argv) > 2 else None


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copy(s, os.path.join(P, dst))
        print("copied", src, "->", dst)


cp(f"{tag}_bench.json", "r02_bench_n1.json")
cp(f"{tag}_bench_ref.json", "r02_bench_reference.json")
cp(f"{tag}_k1_probe.log", "r02_k1_probe.log")
cp(f"{tag}_alt_compare.log", "r02_ab_bit_identity.log")
cp("parity_report.log", "r02_parity_report.log")
if ntag:
    cp(f"{ntag}_launches.csv", "r02_launches.csv")
    for rep, md, traffic in ((f"{ntag}_match_full.ncu-rep", "r02_match_kernel_ncu.md", "match_kernel_traffic.json"),
                             (f"{ntag}_match_b65536.ncu-rep", "r02_match_kernel_ncu_b65536.md", None),
                             (f"{ntag}_match_8192.ncu-rep", "r02_match_kernel_8192_ncu.md", "match_kernel_traffic_8192.json"),
                             (f"{ntag}_slam_step.ncu-rep", "r02_slam_step_ncu.md", None)):
        src = os.path.join(G, rep)
        if not os.path.exists(src):
            print("missing", rep)
            continue
        cmd = [sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), src, os.path.join(P, md)]
        if traffic:
            cmd += ["--traffic", os.path.join(P, traffic)]
        subprocess.run(cmd, check=True)
        print("summarised", rep, "->", md)
