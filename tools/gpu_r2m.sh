#!/usr/bin/env bash
# r2m (1 GPU): reduce-loop unroll (default = 4) vs 2 / 6, producer warpgroup + setmaxnreg (p4); config 4 record.
# Every command under a SHORT timeout (a hung variant must not burn GPU minutes).
tag=${1:-r2m}
out=gpurun_out
mkdir -p $out
for v in default u2 u6 p4; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32" "fused 117 3 f32" "arap 117 3 f32" "snh 117 4 f32" "snh 117 3 f64" "fused 58 3 f32"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4"
    APL_LIB=$lib timeout 90 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== parity of the default build (operators + full size)"
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_zz_fullsize.py -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -4 $out/pytest_${tag}.log
echo "== config 4 (1 GPU)"
timeout 600 python tools/bench_configs.py --config 4 > $out/config4_${tag}.json 2> $out/config4_${tag}.err; tail -c 3000 $out/config4_${tag}.json; tail -5 $out/config4_${tag}.err
