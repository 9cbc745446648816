#!/usr/bin/env bash
# r2q (1 GPU): interleaved A/B, two rounds: base (scalar math, no L2 prefetch), pf0 (packed pairs, 16+8-byte rows, no
# prefetch), pf1 (+ static-stream prefetch), pf2 (+ vertex-row prefetch)
tag=${1:-r2q}
out=gpurun_out
mkdir -p $out
for round in 1 2; do
for cfg in "snh 117 3 f32 11" "fused 117 3 f32 11" "snh 234 3 f32 11" "fused 234 3 f32 11" "snh 117 3 f32 8" "fused 58 3 f32 11"; do
  for v in base pf0 pf1 pf2; do
    set -- $cfg
    echo "== round$round $v $1 n=$2 ld=$3 $4 ops=$5"
    APL_LIB=$PWD/apple_b200/libapple_b200_$v.so timeout 120 python tools/prof_one.py --kind $1 --ops $5 --n $2 --ld $3 --dtype $4 --reps 8 --setup device 2>&1 | tail -1 | grep -o "Gtets.*" || echo "FAILED/TIMEOUT"
  done
done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt | paste - - | awk '{print $2,$3,$4,$5,$6,$7,$8,$10}'
