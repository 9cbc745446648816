#!/usr/bin/env bash
tag=${1:-r2j}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_zz_superset.py tests/test_gpu_pncg.py -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_${tag}.log; tail -40 $out/pytest_${tag}.log
