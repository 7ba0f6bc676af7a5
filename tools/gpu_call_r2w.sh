#!/usr/bin/env bash
# 1 GPU: the new full-size parity tests, the ncu launch list of the bench command, full captures of the dominant kernel
# (refreshes profiles/ncu_traffic.json) and of the final sweep kernel.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -s -k "config3 or config4 or fused" 2>&1 | grep "config\|passed\|failed" | tail -8
export FC_BENCH_NO_PARITY=1 FC_BENCH_NO_CONFIGS=1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-simple --no-cpu > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_bench.csv > gpurun_out/launches_bench.txt 2>/dev/null; head -30 gpurun_out/launches_bench.txt
timeout 500 ncu --set full --clock-control none -k regex:k_dpcg_persist -c 1 -o gpurun_out/prof_persist216 \
    python bench.py --steps 1 --warmup 0 --no-simple --no-cpu > gpurun_out/ncu_persist216.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_persist216.ncu-rep > gpurun_out/prof_persist216.txt 2>/dev/null
grep "dram__bytes\|gpu__time_duration\|registers" gpurun_out/prof_persist216.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile_walk' -c 6 \
    -o gpurun_out/prof_tile_walk_final python tools/simple_iter_bench.py 216 0 1 > gpurun_out/ncu_walk.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_tile_walk_final.ncu-rep > gpurun_out/prof_tile_walk_final.txt 2>/dev/null
grep "^kernel\|gpu__time_duration\|dram__bytes\|registers_per\|warps_active\|inst_executed.sum" gpurun_out/prof_tile_walk_final.txt | cut -c1-140 | head -24
