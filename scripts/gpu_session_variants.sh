#!/bin/bash
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== variant timing"; timeout 900 python scripts/variant_timing.py > $out/${tag}_variants.log 2>&1; cat $out/${tag}_variants.log | grep -v "^$" | head -80
echo "== slam step launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'match_kernel|update_mark|update_apply|slam_gate' -s 1400 -c 60 --csv --log-file $out/${tag}_slam_launches.csv python bench.py --only slam --no-cpu-baseline > $out/${tag}_slam_under_ncu.log 2>&1; tail -30 $out/${tag}_slam_launches.csv | cut -c1-200
echo "== pytest slam step"; timeout 300 python -m pytest tests/test_gpu_slam_step.py -q 2>&1 | tail -3
