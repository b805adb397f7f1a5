#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the NCCL tile-broadcast test and bench.py at N ranks (config 4 strong-scaled,
# config 5 replay with tile broadcasts inside the timed region).  Usage: bash scripts/gpu_session_multi.sh <N> <tag>
N=${1:-2}
tag=${2:-r02m}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi -L > $out/${tag}_gpus.txt
echo "== 2-GPU NCCL tile broadcast test"
timeout 600 python -m pytest tests/test_gpu_tiles.py -q > $out/${tag}_pytest_tiles.log 2>&1; tail -3 $out/${tag}_pytest_tiles.log
for n in ${NLIST:-1 $N}; do
  echo "== bench N=$n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
  fi
  tail -c 400 $out/${tag}_bench_n$n.err; tail -n 1 $out/${tag}_bench_n$n.json | head -c 1200; echo
done
