#!/bin/bash
# ncu --set full capture of selected kernels of one bench command (after the same command exited 0 without ncu).
# Usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> <workload> [size-mib] [skip] [count]
TAG=$1; KRE=$2; WL=${3:-zipf}; MIB=${4:-1024}; SKIP=${5:-3}; CNT=${6:-1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL --size-mib $MIB"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o $OUT/${TAG}_prof $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/${TAG}_ncu.log
ncu -i $OUT/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_source.csv 2>/dev/null
ls -la $OUT/${TAG}_*
