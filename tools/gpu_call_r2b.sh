#!/usr/bin/env bash
# Round 2, 2-GPU call: multi-rank parity (p2p and nccl), bench at N=2 with the full-size oracle parity object.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu_n2_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n2_p2p.log
grep "\[mgpu\]" gpurun_out/mgpu_n2_p2p.log | tail -20
FC_NO_P2P=1 timeout 600 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/mgpu_n2_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n2_nccl.log
grep "\[mgpu\]" gpurun_out/mgpu_n2_nccl.log | tail -20
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
