#!/usr/bin/env bash
# Interleaved same-box A/B of library variants (the only trustworthy comparison: box-to-box variation is +-5 %).
#   APL_BUILD_TAG=x APL_NVCC_FLAGS="-D..." python -m apple_b200.build      # builds apple_b200/libapple_b200_x.so
#   gpurun --timeout 900 -- bash tools/gpu_ab.sh <tag> default x y        # "default" = the product library
# Every run under a SHORT timeout: a hung variant must not burn GPU minutes.
tag=${1:-ab}; shift
variants=("$@")
out=gpurun_out
mkdir -p $out
for round in 1 2; do
for cfg in "snh 117 3 f32 11" "fused 117 3 f32 11" "snh 234 3 f32 11" "fused 234 3 f32 11" "snh 117 3 f32 8" "fused 58 3 f32 11"; do
  read -r kind n ld dt ops <<< "$cfg"
  for v in "${variants[@]}"; do
    lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
    echo "== round$round $v $kind n=$n ld=$ld $dt ops=$ops"
    APL_LIB=$lib timeout 120 python tools/prof_one.py --kind $kind --ops $ops --n $n --ld $ld --dtype $dt --reps 8 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done
done > $out/variants_${tag}.txt 2>&1
paste - - < $out/variants_${tag}.txt | awk '{print $2,$3,$4,$5,$8,$10}'
