#!/usr/bin/env bash
out=gpurun_out; mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29554 tools/peer_check.py > $out/peer_check_w4_v.log 2>&1; echo "rc=$?" >> $out/peer_check_w4_v.log; grep -E "PEER_CHECK|rc=" $out/peer_check_w4_v.log
timeout 300 python -m pytest tests/test_gpu_zz_superset.py -m gpu -x -q -k "fused or psd_passes" > $out/pytest_sup3.log 2>&1; echo "rc=$?" >> $out/pytest_sup3.log; tail -3 $out/pytest_sup3.log
