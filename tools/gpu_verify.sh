#!/usr/bin/env bash
# What the driver runs at round end, on one GPU: the GPU suite, smoke(), the bench line (both arms)
tag=${1:-verify}
out=gpurun_out
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> $out/pytest_${tag}.log; tail -4 $out/pytest_${tag}.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_${tag}.log 2>&1; tail -2 $out/smoke_${tag}.log
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference > $out/bench_ref_${tag}.json 2> $out/bench_${tag}.err; tail -c 200 $out/bench_ref_${tag}.json
echo "== bench"; ( time timeout 600 python bench.py > $out/bench_${tag}.json 2>> $out/bench_${tag}.err ) 2> $out/time_bench_${tag}.txt; python -c "
import json
d=json.loads(open('$out/bench_${tag}.json').read().strip().split('\n')[-1])
print('value %.2f G tets/s, e2e %.2f, frac %.3f, traffic %s, snh alone %.3f, config2 %.2f, pncg %.0f' % (d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['per_kernel']['snh (alone, same mesh)']['frac'], d['config2']['value']/1e9, d['config2']['pncg']['iters_per_s']))
print('parity', d['parity']['within_tolerance'], 'clocks', d['clocks'])
"; tail -2 $out/bench_${tag}.err; cat $out/time_bench_${tag}.txt
