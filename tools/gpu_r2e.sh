#!/usr/bin/env bash
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
for cfg in "snh 117 3" "fused 117 3" "fused 58 3" "snh 117 4" "fused 117 4"; do
  set -- $cfg
  echo "== default $1 n=$2 ld=$3"
  timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*"
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== pytest (new tests + operators)"; timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> $out/pytest_${tag}.log; tail -6 $out/pytest_${tag}.log
