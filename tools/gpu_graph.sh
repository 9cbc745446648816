#!/usr/bin/env bash
# N-GPU step with and without the CUDA-graph replay of (memset, element kernel, push, pull)
tag=${1:-r2z}; N=${2:-2}
out=gpurun_out
mkdir -p $out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for mode in ""; do
  echo "== bench $N GPUs $mode"
  run 200 29561 bench.py --gpus $N --steps 40 --warmup 5 --no-hvp --no-pncg --no-e2e $mode > $out/bench_graph_${tag}_g${N}${mode}.json 2> $out/bench_graph_${tag}.err
  python -c "
import json,sys
d=json.loads(open('$out/bench_graph_${tag}_g${N}${mode}.json').read().strip().split('\n')[-1])
print('value %.2f G tets/s, %.4f ms/step, kernel %.4f ms, parity %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['per_kernel']['snh+arap']['ms'], d['parity']))
print(d['config']['parallelism'][-80:])
" || tail -5 $out/bench_graph_${tag}.err
done
