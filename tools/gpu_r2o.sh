#!/usr/bin/env bash
# r2o (1 GPU): packed fp32 pairs (FFMA2: {F,dF} and {g,Hp} share dhdX) + producer warpgroup for the fused kernels (default)
# vs the scalar form (nopk) vs spin back-off (s32); parity suite; ncu captures of the default build
tag=${1:-r2o}
out=gpurun_out
mkdir -p $out
echo "== parity first (operators, full size)"
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_zz_fullsize.py tests/test_gpu_zz_hostbuf.py -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -4 $out/pytest_${tag}.log
for v in default nopk s32; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32" "fused 117 3 f32" "arap 117 3 f32" "snh 117 4 f32" "fused 117 4 f32" "fused 58 3 f32"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4"
    APL_LIB=$lib timeout 90 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt | paste - -
for cfg in "snh 117" "fused 117"; do
  set -- $cfg; kind=$1; n=$2
  echo "== full capture: $kind n=$n"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fem_pipe_kernel -s 1 -c 1 -f -o $out/prof_${tag}_${kind}_${n} \
      python tools/prof_one.py --kind $kind --ops 11 --n $n --reps 3 --setup device > $out/prof_${tag}_${kind}_${n}.log 2>&1
  if [ -f $out/prof_${tag}_${kind}_${n}.ncu-rep ]; then
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page raw --csv > $out/prof_${tag}_${kind}_${n}_raw.csv 2>/dev/null
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page source --csv > $out/prof_${tag}_${kind}_${n}_src.csv 2>/dev/null
    rm -f $out/prof_${tag}_${kind}_${n}.ncu-rep
  fi
done
