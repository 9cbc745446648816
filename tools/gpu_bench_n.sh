#!/usr/bin/env bash
# bench.py on N GPUs exactly as the driver launches it.  usage: bash tools/gpu_bench_n.sh <tag> <N>
tag=${1:-bn}; N=${2:-8}
out=gpurun_out; mkdir -p $out
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N > $out/bench_${tag}_g${N}.json 2> $out/bench_${tag}_g${N}.err ) 2> $out/time_${tag}_g${N}.txt
python -c "
import json
d=json.loads(open('$out/bench_${tag}_g${N}.json').read().strip().split('\n')[-1])
print('N=%d value %.2f G tets/s, %.4f ms/step, kernel %.4f ms, e2e %.2f, hvp32 %.1f hvp64 %.1f, pncg %.1f it/s, parity %s' % (d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['roofline']['per_kernel']['snh+arap']['ms'], d['e2e']['value']/1e9, d['hvp']['f32']['value']/1e9, d['hvp']['f64']['value']/1e9, d['pncg']['headline_mesh']['iters_per_s'], d['parity']['within_tolerance']))
" || tail -5 $out/bench_${tag}_g${N}.err
cat $out/time_${tag}_g${N}.txt
