#!/bin/bash
# Record session for profiles/: default bench line, short benches of the other workloads, ncu launch list, ncu --set full
# captures of the hot kernels (each only after the same command exited 0 without ncu), SASS opcode histograms.
# Usage (under gpurun, from the repo root):  bash tools/gpu_record.sh <tag>
TAG=${1:-rec}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
grep -m1 "model name" /proc/cpuinfo > $OUT/${TAG}_host.txt; nproc >> $OUT/${TAG}_host.txt
python bench.py > $OUT/${TAG}_bench_zipf.json 2> $OUT/${TAG}_bench_zipf.err; echo "bench zipf exit $?"
for wl in uniform text skewed; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err; echo "bench $wl exit $?"
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload zipf"
$CMD > $OUT/${TAG}_plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launch list exit $?"
for spec in "zipf:hist_kernel|encode_kernel|dec_speculate_kernel|dec_sync_kernel|dec_write_kernel:z:20:5" "uniform:dec_phase_walk_kernel|encode_kernel:u:6:2" "text:encode_kernel:t:3:1"; do
  IFS=: read wl kre sfx skip cnt <<< "$spec"
  C2="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl"
  $C2 > $OUT/${TAG}${sfx}_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o $OUT/${TAG}${sfx}_prof $C2 > $OUT/${TAG}${sfx}_ncu.log 2>&1
  echo "ncu full $wl exit $?"
  ncu -i $OUT/${TAG}${sfx}_prof.ncu-rep --page raw --csv > $OUT/${TAG}${sfx}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}${sfx}_prof.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}${sfx}_source.csv 2>/dev/null
done
ls -la $OUT | grep ${TAG} | head -40
