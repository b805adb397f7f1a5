#!/bin/bash
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
EXTRA="--metrics lts__t_bytes.sum,lts__t_sectors.sum,l1tex__t_bytes.sum,lts__t_sector_hit_rate.pct"
echo "== match kernel at the BASELINE batch (2048^2 map)"
timeout 900 ncu --set full $EXTRA --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:match_kernel<\(int\)1,' -s 3 -c 2 -o $out/${tag}_match_full python bench.py --only value --steps 3 --warmup 3 > $out/${tag}_ncu_match.log 2>&1
tail -3 $out/${tag}_ncu_match.log
echo "== match kernel at 65536 scans per launch"
timeout 900 ncu --set full $EXTRA --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:match_kernel<\(int\)1,' -s 2 -c 1 -o $out/${tag}_match_b65536 python bench.py --only value --batch 65536 --nbuf 2 --steps 2 --warmup 3 > $out/${tag}_ncu_match65536.log 2>&1
tail -3 $out/${tag}_ncu_match65536.log
echo "== match kernel on the 8192^2 map (config 5)"
timeout 1200 ncu --set full $EXTRA --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:match_kernel<\(int\)1,' -s 3 -c 2 -o $out/${tag}_match_8192 python bench.py --only config5 > $out/${tag}_ncu_match8192.log 2>&1
tail -3 $out/${tag}_ncu_match8192.log
echo "== fused SLAM step (single-scan match + mark + apply)"
timeout 900 ncu --set full $EXTRA --clock-control none --import-source on -k regex:'match_kernel|update_mark|update_apply' \
  -s 1200 -c 6 -o $out/${tag}_slam_step python bench.py --only slam --no-cpu-baseline > $out/${tag}_ncu_slam.log 2>&1
echo "== launch list of the default bench command (value + e2e, no extras)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
ls -la $out/${tag}_*.ncu-rep
