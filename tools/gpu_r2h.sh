#!/usr/bin/env bash
# r2h (1 GPU): scalar REDs as `red` instead of ATOMG (default build) vs try_wait suspend hint vs 3 CTAs/SM; the config-4
# test alone; full ncu captures of the default build
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
for v in default hint c3; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3 f32" "fused 117 3 f32" "arap 117 3 f32" "snh 117 4 f32" "snh 117 3 f64" "fused 58 3 f32"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3 $4"
    APL_LIB=$lib timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --dtype $4 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== config-4 test alone"
timeout 600 python -X faulthandler -m pytest tests/test_gpu_zz_fullsize.py -m gpu -x -q -k config4 > $out/pytest_c4_${tag}.log 2>&1; echo "c4 rc=$?" >> $out/pytest_c4_${tag}.log; tail -15 $out/pytest_c4_${tag}.log
dmesg 2>/dev/null | tail -5 > $out/dmesg_${tag}.txt
for cfg in "snh 117" "fused 117"; do
  set -- $cfg; kind=$1; n=$2
  echo "== full capture: $kind n=$n"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fem_pipe_kernel -s 1 -c 1 -f -o $out/prof_${tag}_${kind}_${n} \
      python tools/prof_one.py --kind $kind --ops 11 --n $n --reps 3 --setup device > $out/prof_${tag}_${kind}_${n}.log 2>&1
  if [ -f $out/prof_${tag}_${kind}_${n}.ncu-rep ]; then
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page raw --csv > $out/prof_${tag}_${kind}_${n}_raw.csv 2>/dev/null
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page source --csv > $out/prof_${tag}_${kind}_${n}_src.csv 2>/dev/null
    rm -f $out/prof_${tag}_${kind}_${n}.ncu-rep
  fi
done
ls -la $out | tail -8
