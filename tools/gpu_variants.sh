#!/bin/bash
# Tuning session: bench each prebuilt lib/libgh_b200_<variant>.so (build.py build_variant) on the default workload.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-tune}
for lib in golden-huffman_b200/lib/libgh_b200.so golden-huffman_b200/lib/libgh_b200_*.so; do
  v=$(basename $lib .so)
  GH_LIB_PATH=$PWD/$lib timeout ${VARIANT_TIMEOUT:-90} python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --workload ${WORKLOAD:-zipf} > $OUT/${TAG}_$v.json 2> $OUT/${TAG}_$v.err
  echo "$v exit $?"
done
