#!/bin/bash
# One GPU-box session (round 2): smoke -> parity tests -> short benches of the four workloads.
# Usage (under gpurun, from the repo root):  bash tools/gpu_session.sh <tag> [pytest args]
# Everything lands in gpurun_out/<tag>_*.
TAG=${1:-s}
shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
if [ "${SKIP_SMOKE:-0}" != "1" ]; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
  echo "smoke exit $?"; tail -3 $OUT/${TAG}_smoke.log
fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest ${PYTEST_TARGET:-tests} -m gpu -x -q --timeout 600 "$@" > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?"; tail -12 $OUT/${TAG}_pytest.log
fi
for wl in ${WORKLOADS:-zipf uniform text skewed}; do
  timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --workload $wl ${BENCH_ARGS:---no-e2e --no-cpu-baseline} > $OUT/${TAG}_bench_$wl.json 2> $OUT/${TAG}_bench_$wl.err
  echo "bench $wl exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_$wl.json").read().strip().splitlines()[-1])
    print("  value %.1f GB/s  enc %.3f ms  dec %.3f ms  C=%d" % (d["value"], d["encode_ms"], d["decode_ms"], d["config"]["compressed_bytes_total"]))
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        print("  %-28s %5.1f x %8.4f ms  %s" % (k, v["launches_per_step"], v["avg_ms"], ("%.0f GB/s" % v["GBps"]) if v["GBps"] else ""))
except Exception as e:
    print("  (no json)", e)
PY
  tail -3 $OUT/${TAG}_bench_$wl.err
done
