#!/usr/bin/env bash
# r2n (1 GPU): after moving the superset kernels to their own translation units: reduce unroll 4 (default) vs 1 / 2,
# producer warpgroup (p4); new GPU tests; ncu captures of the default build at n = 117 and n = 234
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
for v in default u1 u2 p4; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32" "fused 117 3 f32" "arap 117 3 f32" "snh 117 4 f32" "snh 117 3 f64" "fused 58 3 f32"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4"
    APL_LIB=$lib timeout 90 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== GPU suite"
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -4 $out/pytest_${tag}.log
