#!/usr/bin/env bash
# r2b: parity of the mbarrier-pipelined element kernel + device-side setup; timing of three builds of the kernel.
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
python -m apple_b200.build > $out/build_${tag}.log 2>&1
echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_${tag}.log 2>&1 ; echo "pytest rc=$?" >> $out/pytest_${tag}.log ; tail -15 $out/pytest_${tag}.log
for v in default nb1 cpasync; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3" "fused 117 3" "fused 58 3" "snh 117 4" "fused 117 4"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3"
    APL_LIB=$lib timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --reps 6 2>&1 | tail -1
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== bench (config 2)" ; timeout 600 python bench.py --no-cpu-baseline > $out/bench_${tag}_1m.json 2> $out/bench_${tag}.err ; tail -c 300 $out/bench_${tag}_1m.json; tail -5 $out/bench_${tag}.err
