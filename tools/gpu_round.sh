#!/usr/bin/env bash
# One GPU-box call that produces everything a round needs (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r2a'
# Every step has its own timeout; everything lands in gpurun_out/ (merged back by gpurun) and
# tools/ncu_summary.py turns the raw ncu pages into the transposed CSVs kept under profiles/.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/gpu.txt 2>&1
python -m apple_b200.build > $out/build.log 2>&1

echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_${tag}.log 2>&1 ; echo "pytest rc=$?" >> $out/pytest_${tag}.log ; tail -3 $out/pytest_${tag}.log
echo "== pair layout (experimental, opt-in)" ; APL_TEST_PAIR=1 timeout 600 python -m pytest tests/test_gpu_zz_pair.py -m gpu -q -rxX > $out/pytest_pair_${tag}.log 2>&1 ; echo "pair rc=$?" >> $out/pytest_pair_${tag}.log ; tail -5 $out/pytest_pair_${tag}.log
echo "== smoke"  ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_${tag}.log 2>&1 ; echo "smoke rc=$?" >> $out/smoke_${tag}.log
echo "== bench (config 2)" ; timeout 900 python bench.py > $out/bench_${tag}_1m.json 2> $out/bench_${tag}.err ; tail -c 600 $out/bench_${tag}_1m.json
echo "== bench (8 M tets, operators only)" ; timeout 600 python bench.py --n 117 --steps 10 --no-pncg --no-cpu-baseline > $out/bench_${tag}_8m.json 2>> $out/bench_${tag}.err
echo "== bench (SNH alone, 8 M)" ; timeout 600 python bench.py --n 117 --steps 10 --potentials snh --no-pncg --no-cpu-baseline > $out/bench_${tag}_8m_snh.json 2>> $out/bench_${tag}.err
echo "== bench (pair layout, config 2 and SNH 8 M)" ; timeout 600 python bench.py --layout pair --no-cpu-baseline > $out/bench_${tag}_1m_pair.json 2>> $out/bench_${tag}.err
timeout 600 python bench.py --layout pair --n 117 --steps 10 --potentials snh --no-pncg --no-cpu-baseline > $out/bench_${tag}_8m_snh_pair.json 2>> $out/bench_${tag}.err
echo "== reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_${tag}_reference.json 2>> $out/bench_${tag}.err

echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-pncg --no-cpu-baseline > $out/ncu_bench_${tag}.log 2>&1

for kind in snh fused; do
  echo "== full capture: $kind"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fem_pipe_kernel -s 1 -c 1 -f -o $out/prof_${tag}_${kind} \
      python tools/prof_one.py --kind $kind --ops 11 --n 117 --reps 3 > $out/prof_${tag}_${kind}.log 2>&1
  if [ -f $out/prof_${tag}_${kind}.ncu-rep ]; then
    ncu -i $out/prof_${tag}_${kind}.ncu-rep --page raw --csv > $out/prof_${tag}_${kind}_raw.csv 2>/dev/null
    python tools/ncu_summary.py $out/prof_${tag}_${kind}_raw.csv $out/${tag}_ncu_full_fem_pipe_kernel_${kind}_8m.csv 8008065
  fi
done
echo "== sweep" ; timeout 900 python bench.py --sweep --steps 3 --warmup 1 --no-pncg --no-cpu-baseline > /dev/null 2> $out/sweep_${tag}.txt
ls -la $out | tail -30
