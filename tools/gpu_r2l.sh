#!/usr/bin/env bash
# r2l (1 GPU): config 4 record (forward to equilibrium, adjoint), kernel variants p4 / sleep vs default
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
for v in default p4 sleep; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32" "fused 117 3 f32" "arap 117 3 f32" "snh 117 4 f32" "fused 58 3 f32"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4"
    APL_LIB=$lib timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== p4 parity (operators suite on the variant library)"
APL_LIB=$PWD/apple_b200/libapple_b200_p4.so timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_zz_fullsize.py -m gpu -x -q > $out/pytest_p4_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_p4_${tag}.log; tail -5 $out/pytest_p4_${tag}.log
echo "== config 4 (1 GPU)"
timeout 900 python tools/bench_configs.py --config 4 > $out/config4_${tag}.json 2> $out/config4_${tag}.err; tail -c 3000 $out/config4_${tag}.json; tail -5 $out/config4_${tag}.err
