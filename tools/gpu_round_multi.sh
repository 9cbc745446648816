#!/usr/bin/env bash
# Multi-GPU part of a round (costs N x the wall time: keep every step under its own short timeout):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round_multi.sh 2 r2a'
N=${1:-2}
tag=${2:-rX}
out=gpurun_out
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
python -m apple_b200.build > $out/build_multi.log 2>&1
echo "== parity: sharded vs single GPU, split vs plain evaluation"
timeout 300 bash -c "$(declare -f run); N=$N run tools/dist_check.py" > $out/dist_${tag}_n$N.log 2>&1 ; tail -4 $out/dist_${tag}_n$N.log
echo "== weak scaling (~975k tets per GPU), halo exchange overlapped / not overlapped"
timeout 240 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --steps 20 --warmup 5" > $out/bench_${tag}_weak_n$N.json 2> $out/bench_${tag}_multi.err
timeout 240 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --steps 20 --warmup 5 --no-overlap" > $out/bench_${tag}_weak_n${N}_serial.json 2>> $out/bench_${tag}_multi.err
echo "== strong scaling, 8 M tets, per-rank slab generation (run the same line with 1 GPU for the denominator)"
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --slab --n 117 --steps 20 --warmup 5" > $out/bench_${tag}_strong8m_n$N.json 2>> $out/bench_${tag}_multi.err
tail -c 400 $out/bench_${tag}_weak_n$N.json; echo; tail -c 400 $out/bench_${tag}_strong8m_n$N.json; echo
