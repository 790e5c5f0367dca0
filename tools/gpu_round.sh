#!/bin/bash
# One GPU-box session: smoke -> parity tests -> bench -> ncu launch list -> ncu full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
# Everything lands in gpurun_out/<tag>_*.  ncu only runs after the same command exited 0 without it.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
nproc > $OUT/${TAG}_host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/${TAG}_host.txt

echo "== smoke" | tee $OUT/${TAG}_smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 ${PYTEST_ARGS:-} > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
fi

echo "== bench"
for wl in ${WORKLOADS:-zipf uniform}; do
  timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --workload $wl > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err
  echo "bench $wl exit $?"; tail -c 3000 $OUT/${TAG}_bench_$wl.json; tail -5 $OUT/${TAG}_bench_$wl.err
done

if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload zipf"
  $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  echo "== ncu full"
  [ "${SKIP_NCU_FULL:-0}" != "1" ] && $CMD > $OUT/${TAG}_ncu_plain2.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"hist_kernel|encode_kernel|dec_.*speculate_kernel|dec_sync_kernel|dec_.*write_kernel" -s ${NCU_SKIP:-24} -c ${NCU_COUNT:-8} -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la $OUT | tail -20
