#!/usr/bin/env bash
# Final 1-GPU round: GPU suite, smoke, bench (ours + reference arm), launch lists, full ncu captures of the shipped kernels
# at the headline size, config-4 record in fp64.  Every command under its own timeout.
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/gpu_${tag}.txt 2>&1; nproc >> $out/gpu_${tag}.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> $out/pytest_${tag}.log; tail -14 $out/pytest_${tag}.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_${tag}.log 2>&1; tail -4 $out/smoke_${tag}.log
echo "== bench"; ( time timeout 800 python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err ) 2> $out/time_bench_${tag}.txt; tail -c 400 $out/bench_${tag}.json; tail -3 $out/bench_${tag}.err; cat $out/time_bench_${tag}.txt
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference > $out/bench_ref_${tag}.json 2>> $out/bench_${tag}.err; tail -c 600 $out/bench_ref_${tag}.json
echo "== launch list (bench step at 64 M)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-pncg --no-cpu-baseline --no-config2 --no-hvp --no-e2e --no-parity > $out/ncu_bench_${tag}.log 2>&1
echo "== launch list (pncg eager, config 2)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches_pncg_${tag}.csv \
    python tools/prof_pncg.py --iters 3 > $out/ncu_pncg_${tag}.log 2>&1
echo "== full capture pncg / pcg vector kernels"
timeout 400 ncu --set full --clock-control none -k regex:'pncg|pcg_|ext_force' -c 14 -f -o $out/prof_${tag}_pncg \
    python tools/prof_pncg.py --iters 2 > $out/prof_${tag}_pncg.log 2>&1
[ -f $out/prof_${tag}_pncg.ncu-rep ] && ncu -i $out/prof_${tag}_pncg.ncu-rep --page raw --csv > $out/prof_${tag}_pncg_raw.csv 2>/dev/null && rm -f $out/prof_${tag}_pncg.ncu-rep
for cfg in "fused 234 11" "snh 234 11" "fused 58 11" "snh 234 8"; do
  set -- $cfg; kind=$1; n=$2; ops=$3
  echo "== full capture: $kind n=$n ops=$ops"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fem_pipe_kernel -s 1 -c 1 -f -o $out/prof_${tag}_${kind}_${n}_${ops} \
      python tools/prof_one.py --kind $kind --ops $ops --n $n --reps 3 --setup device > $out/prof_${tag}_${kind}_${n}_${ops}.log 2>&1
  if [ -f $out/prof_${tag}_${kind}_${n}_${ops}.ncu-rep ]; then
    ncu -i $out/prof_${tag}_${kind}_${n}_${ops}.ncu-rep --page raw --csv > $out/prof_${tag}_${kind}_${n}_${ops}_raw.csv 2>/dev/null
    ncu -i $out/prof_${tag}_${kind}_${n}_${ops}.ncu-rep --page source --csv > $out/prof_${tag}_${kind}_${n}_${ops}_src.csv 2>/dev/null
    rm -f $out/prof_${tag}_${kind}_${n}_${ops}.ncu-rep
  fi
done
echo "== config 4 fp64 / fp32"
timeout 400 python tools/bench_configs.py --config 4 --dtype f64 > $out/config4_f64_${tag}.json 2> $out/config4_${tag}.err; tail -c 1500 $out/config4_f64_${tag}.json
timeout 300 python tools/bench_configs.py --config 4 --dtype f32 > $out/config4_f32_${tag}.json 2>> $out/config4_${tag}.err; tail -c 300 $out/config4_f32_${tag}.json
echo "== config 3 (1 GPU)"
timeout 300 python tools/bench_configs.py --config 3 > $out/config3_g1_${tag}.json 2> $out/config3_${tag}.err; tail -c 600 $out/config3_g1_${tag}.json
ls -la $out | tail -5
