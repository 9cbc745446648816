#!/usr/bin/env bash
# Final multi-GPU round on N GPUs of one box: strong-scaling bench line at N (and at 1 on the same box when N == 2), 2-GPU
# hardware parity, config 3 sharded.  usage: bash tools/gpu_final_multi.sh <tag> <N>
tag=${1:-r2z}; N=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/gpus_${tag}_g${N}.txt 2>&1
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
if [ "$N" = "2" ]; then
  echo "== pytest 2-GPU test"; timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $out/pytest_dist_${tag}.log 2>&1; echo "rc=$?" >> $out/pytest_dist_${tag}.log; tail -3 $out/pytest_dist_${tag}.log
  echo "== dist_check"; run 400 29533 tools/dist_check.py > $out/dist_${tag}.log 2>&1; echo "dist rc=$?" >> $out/dist_${tag}.log; grep -v "^W\|^\[W\|Warning" $out/dist_${tag}.log | tail -12
fi
echo "== bench $N GPUs (64 M tets, strong scaling)"
( time run 240 29535 bench.py --gpus $N > $out/bench_${tag}_g${N}.json 2> $out/bench_${tag}_g${N}.err ) 2> $out/time_${tag}_g${N}.txt; tail -c 500 $out/bench_${tag}_g${N}.json; tail -3 $out/bench_${tag}_g${N}.err; cat $out/time_${tag}_g${N}.txt
echo "== config 3 on $N GPUs"
run 150 29541 tools/bench_configs.py --config 3 > $out/config3_g${N}_${tag}.json 2> $out/config3_${tag}_g${N}.err; tail -c 400 $out/config3_g${N}_${tag}.json
if [ "$N" = "2" ]; then
  echo "== launch list of one sharded step (rank 0 under ncu is not supported for multi-rank: skipped)"
fi
