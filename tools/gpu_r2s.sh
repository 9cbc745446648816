#!/usr/bin/env bash
# r2s (1 GPU): producer warps parked at a hardware barrier (default) vs all four polling (nopark); parity of the fused paths
tag=${1:-r2s}
out=gpurun_out
mkdir -p $out
echo "== parity (full size + operators + pncg)"
timeout 400 python -m pytest tests/test_gpu_zz_fullsize.py tests/test_gpu_operators.py tests/test_gpu_pncg.py -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -3 $out/pytest_${tag}.log
for round in 1 2; do
for cfg in "fused 117 3 f32 11" "fused 234 3 f32 11" "fused 58 3 f32 11" "fused 117 3 f32 7" "fused 117 3 f32 16"; do
  for v in default nopark; do
    set -- $cfg
    lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
    echo "== round$round $v $1 n=$2 ld=$3 $4 ops=$5"
    APL_LIB=$lib timeout 120 python tools/prof_one.py --kind $1 --ops $5 --n $2 --ld $3 --dtype $4 --reps 8 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt | paste - - | awk '{print $2,$3,$4,$5,$8,$10}'
