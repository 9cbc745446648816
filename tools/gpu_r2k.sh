#!/usr/bin/env bash
# r2k (2 GPUs): recorded runs of BASELINE configs 3 (1 and 2 GPUs) and 4 (1 GPU)
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
echo "== config 4 (1 GPU)"
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/bench_configs.py --config 4 > $out/config4_${tag}.json 2> $out/config4_${tag}.err; tail -c 2500 $out/config4_${tag}.json; tail -5 $out/config4_${tag}.err
echo "== config 3 (1 GPU)"
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/bench_configs.py --config 3 > $out/config3_g1_${tag}.json 2> $out/config3_${tag}.err; tail -c 2500 $out/config3_g1_${tag}.json; tail -5 $out/config3_${tag}.err
echo "== config 3 (2 GPUs)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_configs.py --config 3 > $out/config3_g2_${tag}.json 2>> $out/config3_${tag}.err; tail -c 2500 $out/config3_g2_${tag}.json; tail -5 $out/config3_${tag}.err
