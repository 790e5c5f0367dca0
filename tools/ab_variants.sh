#!/bin/bash
# A/B of tuning builds (build.py build_variant): one short Zipf bench per library, the decode kernels' times side by side.
# Usage (under gpurun):  bash tools/ab_variants.sh <tag> <variant> [<variant> ...]     ("main" = the product library)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  lib=golden-huffman_b200/lib/libgh_b200_$v.so
  [ "$v" = main ] && lib=golden-huffman_b200/lib/libgh_b200.so
  GH_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps ${STEPS:-8} --warmup 3 --workload ${WL:-zipf} --no-e2e --no-cpu-baseline > $OUT/${TAG}_$v.json 2> $OUT/${TAG}_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$v.json").read().strip().splitlines()[-1])
    k = d["kernels"]
    print("%-10s value %.1f enc %.3f dec %.3f | " % ("$v", d["value"], d["encode_ms"], d["decode_ms"]) + "  ".join("%s %.4f" % (n.split("_kernel")[0], k[n]["avg_ms"]) for n in sorted(k, key=lambda n: -k[n]["ms_per_step"])[:4]))
except Exception as e:
    print("$v (no json)", e)
PY
done
