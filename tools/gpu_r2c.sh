#!/usr/bin/env bash
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
for v in default nb1 ldg; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3" "fused 117 3" "fused 58 3" "snh 117 4" "fused 117 4"; do
    set -- $cfg
    [ $v = ldg ] && [ $3 = 4 ] && continue
    echo "== $v $1 n=$2 ld=$3"
    APL_LIB=$lib timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --reps 6 --setup device 2>&1 | tail -2
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== bench n=58 (new bench.py, all sections)"; timeout 900 python bench.py --n 58 > $out/bench_${tag}_n58.json 2> $out/bench_${tag}.err; tail -c 1500 $out/bench_${tag}_n58.json; tail -5 $out/bench_${tag}.err
echo "== bench n=234"; ( time timeout 1200 python bench.py > $out/bench_${tag}_n234.json 2>> $out/bench_${tag}.err ) 2> $out/time_${tag}_n234.txt; tail -c 2500 $out/bench_${tag}_n234.json; tail -5 $out/bench_${tag}.err; cat $out/time_${tag}_n234.txt
