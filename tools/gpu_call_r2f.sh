#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
FCAPP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu -s -k "tiled-walk-flags" > gpurun_out/pytest_walk5.log 2>&1
grep "sweeps\]\|passed\|failed\|Error\|error" gpurun_out/pytest_walk5.log | tail -12
timeout 400 python tools/sweep_bench.py sweeps 216 > gpurun_out/sweep_bench_216_walk5.jsonl 2> gpurun_out/sweep_bench.err
cut -c1-250 gpurun_out/sweep_bench_216_walk5.jsonl; tail -3 gpurun_out/sweep_bench.err
FCAPP_TUNE="sweep_tiled=5" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_walk5.json 2>&1
cut -c1-700 gpurun_out/simple_iter_216_walk5.json
