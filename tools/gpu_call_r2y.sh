#!/usr/bin/env bash
# round 2: persistent DPCG kernel -- eager x update behind the beta reduction (FC_TUNE_DPCG_EAGER), column codes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/pytest_coded.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_coded.log
tail -3 gpurun_out/pytest_coded.log
{
timeout 200 python tools/sweep_bench.py knobs 108 ja_coded=0+dpcg_eager=0 dpcg_eager=1 ja_coded=1+dpcg_eager=0 dpcg_eager=1 ja_coded=0+dpcg_eager=1
timeout 200 python tools/sweep_bench.py knobs 136 ja_coded=0+dpcg_eager=0 dpcg_eager=1 ja_coded=1+dpcg_eager=0 dpcg_eager=1
timeout 300 python tools/sweep_bench.py knobs 216 ja_coded=0+dpcg_eager=0 dpcg_eager=1 ja_coded=1+dpcg_eager=0 dpcg_eager=1
} > gpurun_out/eager.jsonl 2> gpurun_out/eager.err
cut -c1-400 gpurun_out/eager.jsonl; tail -5 gpurun_out/eager.err
