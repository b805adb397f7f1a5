#!/bin/bash
# One GPU-box session: everything this round needs measured, most important first, each step with its own timeout and log
# under gpurun_out/ (merged back by gpurun).  Usage (from the repo root on the box): bash scripts/gpu_session.sh <tag>
tag=${1:-s}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== pytest (full gpu suite)"; timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
echo "== k1 probe"; timeout 300 python scripts/k1_probe.py 4096 > $out/${tag}_k1_probe.log 2>&1; head -60 $out/${tag}_k1_probe.log
timeout 300 python scripts/k1_probe.py 65536 > $out/${tag}_k1_probe_b65536.log 2>&1; head -4 $out/${tag}_k1_probe_b65536.log
echo "== single-scan step, gather batch 4 vs 5"; for u in 4 5; do timeout 300 python bench.py --only slam --no-cpu-baseline --tune unroll=$u > $out/${tag}_slam_u$u.json 2>> $out/${tag}_bench.err; python -c "
import json,sys; d=json.load(open('$out/${tag}_slam_u$u.json')); c=d['config3_slam_step']; print('unroll $u:', {k:round(v,1) for k,v in c.items() if isinstance(v,float)}, 'k2_ms', d['roofline_k2']['kernel_ms'])"; done
echo "== bench N=1"; timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.err; head -c 1500 $out/${tag}_bench.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
