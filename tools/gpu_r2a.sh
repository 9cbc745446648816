#!/usr/bin/env bash
# First GPU call of round 2: parity suite, bench, launch lists and full ncu captures of the SHIPPED kernels.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/gpu.txt 2>&1
nproc >> $out/gpu.txt; free -g >> $out/gpu.txt
python -m apple_b200.build > $out/build.log 2>&1
echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > $out/pytest_${tag}.log 2>&1 ; echo "pytest rc=$?" >> $out/pytest_${tag}.log ; tail -5 $out/pytest_${tag}.log
echo "== pair" ; APL_TEST_PAIR=1 timeout 600 python -m pytest tests/test_gpu_zz_pair.py -m gpu -q -rxX > $out/pytest_pair_${tag}.log 2>&1 ; echo "pair rc=$?" >> $out/pytest_pair_${tag}.log ; tail -3 $out/pytest_pair_${tag}.log
echo "== bench (config 2)" ; timeout 600 python bench.py --no-probe > $out/bench_${tag}_1m.json 2> $out/bench_${tag}.err ; tail -c 300 $out/bench_${tag}_1m.json
echo "== launch list (operators)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-pncg --no-cpu-baseline --no-probe > $out/ncu_bench_${tag}.log 2>&1
echo "== launch list (pncg eager)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_pncg_${tag}.csv \
    python tools/prof_pncg.py --iters 3 > $out/ncu_pncg_${tag}.log 2>&1
echo "== full capture pncg vector kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pncg|halo|ext_force' -c 12 -f -o $out/prof_${tag}_pncg \
    python tools/prof_pncg.py --iters 2 > $out/prof_${tag}_pncg.log 2>&1
[ -f $out/prof_${tag}_pncg.ncu-rep ] && ncu -i $out/prof_${tag}_pncg.ncu-rep --page raw --csv > $out/prof_${tag}_pncg_raw.csv 2>/dev/null
for cfg in "fused 58" "fused 117" "snh 117"; do
  set -- $cfg; kind=$1; n=$2
  echo "== full capture: $kind n=$n"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fem_pipe_kernel -s 1 -c 1 -f -o $out/prof_${tag}_${kind}_${n} \
      python tools/prof_one.py --kind $kind --ops 11 --n $n --reps 3 > $out/prof_${tag}_${kind}_${n}.log 2>&1
  if [ -f $out/prof_${tag}_${kind}_${n}.ncu-rep ]; then
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page raw --csv > $out/prof_${tag}_${kind}_${n}_raw.csv 2>/dev/null
    ncu -i $out/prof_${tag}_${kind}_${n}.ncu-rep --page source --csv > $out/prof_${tag}_${kind}_${n}_src.csv 2>/dev/null
  fi
done
echo "== 8M benches (timing of setup too)"
( time timeout 600 python bench.py --n 117 --steps 10 --no-pncg --no-cpu-baseline --no-probe > $out/bench_${tag}_8m.json 2>> $out/bench_${tag}.err ) 2> $out/time_8m.txt
ls -la $out | tail -40
