#!/usr/bin/env bash
# round 2, matrix-in-L2 experiment: share of the matrix chunks marked evict_last in the persistent DPCG kernel
set -u
mkdir -p gpurun_out
{
timeout 200 python tools/sweep_bench.py knobs 108 mat_keep=0 mat_keep=20 mat_keep=35 mat_keep=50 mat_keep=65 mat_keep=80 mat_keep=100 mat_keep=50+l2_keep=1 mat_keep=65+l2_keep=1 mat_keep=0+l2_keep=2
timeout 200 python tools/sweep_bench.py knobs 136 mat_keep=0 mat_keep=10 mat_keep=20 mat_keep=30 mat_keep=20+l2_keep=1 mat_keep=0+l2_keep=2
timeout 300 python tools/sweep_bench.py knobs 216 mat_keep=0 mat_keep=4 mat_keep=8 mat_keep=12 mat_keep=8+l2_keep=1 mat_keep=0+l2_keep=2
} > gpurun_out/mat_keep.jsonl 2> gpurun_out/mat_keep.err
cut -c1-330 gpurun_out/mat_keep.jsonl; tail -5 gpurun_out/mat_keep.err
