#!/usr/bin/env bash
# 4 and 8 ranks sharing ONE GPU: the peer-memory exchange with middle ranks (two neighbours) on a 1-GPU box
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
for w in 4 8; do
  echo "== peer_check world=$w"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2955$w tools/peer_check.py > $out/peer_check_w${w}_${tag}.log 2>&1; echo "rc=$?" >> $out/peer_check_w${w}_${tag}.log
  grep -v "^W\|Warning\|OMP_NUM\|\*\*\*\*" $out/peer_check_w${w}_${tag}.log | tail -14
done
