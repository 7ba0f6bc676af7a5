#!/usr/bin/env bash
# 8 GPUs: the driver's N = 8 commands (reference arm, then ours with parity + config 5) and the multi-rank parity check
# on 8 ranks.
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nproc; free -g | head -2 | tail -1
t0=$(date +%s)
timeout 900 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
t1=$(date +%s); echo "our arm N=$N: $((t1-t0)) s"
tail -c 7000 gpurun_out/bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -5
timeout 600 $TR --master-port 29552 tests/mgpu_check.py > gpurun_out/mgpu_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n$N.log
t2=$(date +%s); echo "mgpu_check N=$N: $((t2-t1)) s"
grep "\[mgpu\]\|rc=" gpurun_out/mgpu_n$N.log | tail -32
timeout 600 $TR --master-port 29553 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
t3=$(date +%s); echo "reference arm N=$N: $((t3-t2)) s"
tail -c 1200 gpurun_out/bench_ref_n$N.json
