#!/usr/bin/env bash
# r2p (1 GPU): L2 bulk prefetch of the static stream two tiles ahead (default) vs without (nopf) vs scalar math (nopk)
tag=${1:-r2p}
out=gpurun_out
mkdir -p $out
for v in default nopf nopk; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32 11" "fused 117 3 f32 11" "arap 117 3 f32 11" "snh 117 4 f32 11" "snh 117 3 f64 11" "fused 58 3 f32 11" "snh 117 3 f32 8" "snh 234 3 f32 11" "fused 234 3 f32 11"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4 ops=$5"
    APL_LIB=$lib timeout 120 python tools/prof_one.py --kind $1 --ops $5 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt | paste - -
echo "== parity (operators, full size, superset)"
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_zz_fullsize.py tests/test_gpu_zz_superset.py -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -4 $out/pytest_${tag}.log
