#!/usr/bin/env bash
# Round 2, call d (1 GPU): the narrow-CTA tile walk (FC_TUNE_SWEEP_TILED=4)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
FCAPP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu -s -k "tiled-walk" > gpurun_out/pytest_walk.log 2>&1
grep "sweeps\]\|passed\|failed\|Error\|error" gpurun_out/pytest_walk.log | tail -12
timeout 400 python tools/sweep_bench.py sweeps 216 > gpurun_out/sweep_bench_216_walk.jsonl 2> gpurun_out/sweep_bench.err
cut -c1-330 gpurun_out/sweep_bench_216_walk.jsonl; tail -3 gpurun_out/sweep_bench.err
FCAPP_TUNE="sweep_tiled=4" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_walk.json 2>&1
cut -c1-900 gpurun_out/simple_iter_216_walk.json
FCAPP_TUNE="sweep_tiled=4" timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile_walk' -c 6 \
    -o gpurun_out/prof_tile_walk python tools/simple_iter_bench.py 216 0 1 > gpurun_out/ncu_walk.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_tile_walk.ncu-rep > gpurun_out/prof_tile_walk.txt 2>/dev/null
grep "^kernel\|gpu__time_duration\|registers_per\|warps_active\|occupancy_limit\|inst_executed.sum\|dram__bytes" gpurun_out/prof_tile_walk.txt | cut -c1-150 | head -60
timeout 200 python tools/config3_bench.py 100 > gpurun_out/config3.json 2> gpurun_out/config3.err
cut -c1-1500 gpurun_out/config3.json; tail -3 gpurun_out/config3.err
