#!/usr/bin/env bash
out=gpurun_out; mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/xchg_timing.py > $out/xchg_timing_g2.txt 2> $out/xchg_timing.err; grep -v "^W\|Warning" $out/xchg_timing_g2.txt | tail -4; tail -3 $out/xchg_timing.err | cut -c1-200
