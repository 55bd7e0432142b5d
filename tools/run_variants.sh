#!/bin/bash
# On the GPU box: for every build/variants/libb200pose_<name>.so a parity subset and one bench line.
# Usage: bash tools/run_variants.sh [name ...]      (results: gpurun_out/variants/<name>.{json,log})
mkdir -p gpurun_out/variants
names="$@"
[ -z "$names" ] && names=$(ls build/variants/libb200pose_*.so | sed 's/.*libb200pose_\(.*\)\.so/\1/')
for v in $names; do
  lib=$PWD/build/variants/libb200pose_$v.so
  [ -f "$lib" ] || { echo "$v: missing $lib"; continue; }
  B200POSE_LIB=$lib timeout 300 python -m pytest tests/test_gpu.py -x -q -m gpu \
      -k "post_kernels or net_368 or batch_rows or fused_engine" > gpurun_out/variants/$v.log 2>&1
  echo "$v parity rc=$? $(tail -1 gpurun_out/variants/$v.log)"
  B200POSE_LIB=$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline \
      > gpurun_out/variants/$v.json 2>> gpurun_out/variants/$v.log
done
python tools/variants.py report
