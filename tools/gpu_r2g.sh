#!/usr/bin/env bash
# r2g (2 GPUs): box diagnostics, the config-4 test alone (it was SIGKILLed in r2e), 3-CTA variant timing, whole GPU suite
# (incl. the 2-GPU test), dist_check, 2-GPU strong-scaling bench at 8 M and 64 M tets, 1-GPU bench at 64 M on the same box
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
{ nvidia-smi -L; free -g; nproc; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null; ulimit -a; df -h /dev/shm /tmp; } > $out/box_${tag}.txt 2>&1
echo "== config-4 test alone"
/usr/bin/time -v timeout 600 python -m pytest tests/test_gpu_zz_fullsize.py -m gpu -x -q -k config4 > $out/pytest_c4_${tag}.log 2>&1; echo "c4 rc=$?" >> $out/pytest_c4_${tag}.log; grep -E "passed|failed|rc=|Maximum resident|Elapsed" $out/pytest_c4_${tag}.log
dmesg 2>/dev/null | tail -5 > $out/dmesg_${tag}.txt
for v in default c3; do
  lib=""; [ $v != default ] && lib=$PWD/apple_b200/libapple_b200_$v.so
  for cfg in "snh 117 3" "arap 117 3" "fused 117 3" "snh 117 4"; do
    set -- $cfg
    echo "== $v $1 n=$2 ld=$3"
    CUDA_VISIBLE_DEVICES=0 APL_LIB=$lib timeout 300 python tools/prof_one.py --kind $1 --ops 11 --n $2 --ld $3 --reps 6 --setup device 2>&1 | tail -1 | grep -o "Gtets.*"
  done
done > $out/variants_${tag}.txt 2>&1
cat $out/variants_${tag}.txt
echo "== dist_check (2 GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py > $out/dist_${tag}.log 2>&1; echo "dist rc=$?" >> $out/dist_${tag}.log; grep -v "^W\|^\[W" $out/dist_${tag}.log | tail -40
echo "== pytest (without config4)"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zz_fullsize.py::test_config4_muscle_model_full_size_matches_c_oracle > $out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> $out/pytest_${tag}.log; tail -6 $out/pytest_${tag}.log
echo "== bench 2 GPUs n=117"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --n 117 --steps 20 --warmup 5 > $out/bench_${tag}_n117_g2.json 2> $out/bench_${tag}.err; tail -c 1200 $out/bench_${tag}_n117_g2.json; tail -3 $out/bench_${tag}.err
echo "== bench 2 GPUs n=234"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_${tag}_n234_g2.json 2>> $out/bench_${tag}.err ) 2> $out/time_${tag}.txt; tail -c 600 $out/bench_${tag}_n234_g2.json; tail -3 $out/bench_${tag}.err; cat $out/time_${tag}.txt
echo "== bench 1 GPU n=234 (same box)"
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-config2 --no-cpu-baseline > $out/bench_${tag}_n234_g1.json 2>> $out/bench_${tag}.err; tail -c 300 $out/bench_${tag}_n234_g1.json
