#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
FCAPP_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu > gpurun_out/pytest_experimental.log 2>&1
tail -3 gpurun_out/pytest_experimental.log
timeout 400 python tools/sweep_bench.py sweeps 216 > gpurun_out/sweep_bench_216.jsonl 2> gpurun_out/sweep_bench.err
cut -c1-200 gpurun_out/sweep_bench_216.jsonl; tail -3 gpurun_out/sweep_bench.err
timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216.json 2>&1
cut -c1-700 gpurun_out/simple_iter_216.json
timeout 200 python tools/config3_bench.py 100 > gpurun_out/config3.json 2> gpurun_out/config3.err
cut -c1-1300 gpurun_out/config3.json; tail -3 gpurun_out/config3.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
