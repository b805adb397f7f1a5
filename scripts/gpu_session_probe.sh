#!/bin/bash
# Short K1-only session: the probe at one wave and at many waves.  Usage: bash scripts/gpu_session_probe.sh <tag>
tag=${1:-p}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 300 python scripts/k1_probe.py 4096 > $out/${tag}_k1_probe.log 2>&1; grep -v warpid $out/${tag}_k1_probe.log | head -60
timeout 300 python scripts/k1_probe.py 65536 > $out/${tag}_k1_probe_b65536.log 2>&1; head -6 $out/${tag}_k1_probe_b65536.log
timeout 300 python -m pytest tests/test_gpu_bench_launch.py -q 2>&1 | tail -3
