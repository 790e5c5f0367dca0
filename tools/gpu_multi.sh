#!/bin/bash
# Multi-GPU session (gpurun --gpus N): NCCL sharded parity test + bench at N ranks.
N=${1:-2}
TAG=${2:-r1m}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 800 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
for n in ${RANKS:-1 $N}; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29411 \
      bench.py --gpus $n --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err
  fi
  echo "bench n=$n exit $?"; tail -c 1500 $OUT/${TAG}_bench_n$n.json; tail -3 $OUT/${TAG}_bench_n$n.err
done
