#!/usr/bin/env bash
# 8 GPUs: the pressure-correction section of the multi-rank parity check with BiCGStab at a tolerance it reaches
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MGPU_SECTIONS=core timeout 300 $TR --master-port 29571 tests/mgpu_check.py > gpurun_out/mgpu_n8_final.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n8_final.log
grep "\[mgpu\]\|rc=" gpurun_out/mgpu_n8_final.log | tail -14
