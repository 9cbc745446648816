#!/usr/bin/env bash
# r2i (1 GPU): FFMA2 micro-benchmark, new superset tests, config-4 test, whole GPU suite
tag=${1:-r2i}
out=gpurun_out
mkdir -p $out
echo "== f32x2 ubench"; timeout 120 tools/ubench/f32x2 > $out/ubench_f32x2_${tag}.txt 2>&1; cat $out/ubench_f32x2_${tag}.txt
echo "== superset tests"; timeout 900 python -m pytest tests/test_gpu_zz_superset.py -m gpu -x -q > $out/pytest_sup_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_sup_${tag}.log; tail -25 $out/pytest_sup_${tag}.log
echo "== config-4 test alone"
timeout 600 python -m pytest tests/test_gpu_zz_fullsize.py -m gpu -x -q -k config4 > $out/pytest_c4_${tag}.log 2>&1; echo "c4 rc=$?" >> $out/pytest_c4_${tag}.log; tail -8 $out/pytest_c4_${tag}.log
echo "== pytest (rest)"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zz_fullsize.py::test_config4_muscle_model_full_size_matches_c_oracle --ignore tests/test_gpu_zz_superset.py > $out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> $out/pytest_${tag}.log; tail -6 $out/pytest_${tag}.log
