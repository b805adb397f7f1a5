#!/bin/bash
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
echo "== pytest bench-launch + properties"; timeout 600 python -m pytest tests/test_gpu_bench_launch.py tests/test_gpu_properties.py tests/test_gpu_ranges.py tests/test_gpu_cloud.py -q 2>&1 | tail -3
echo "== bench N=1"; timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 300 $out/${tag}_bench.err; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print('value', d['value']); print('e2e', d['e2e']['value'], 'blocking', d['e2e']['blocking_call_value']); print('e2e_endpoints', d['e2e_endpoints']['value'], 'cloud', d['e2e_cloud']['value'])"
bash scripts/gpu_session_ncu.sh $tag
