#!/usr/bin/env bash
# Round 2, first GPU call (1 GPU): the never-run experimental kernels, the per-operation roofline table and the
# full ncu captures of the assembly / sweep kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
FCAPP_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu -s > gpurun_out/pytest_experimental.log 2>&1
grep "sweeps\]\|passed\|failed\|config 3" gpurun_out/pytest_experimental.log | tail -40
timeout 400 python tools/kernel_bench.py 216 10 > gpurun_out/kernel_bench_216.jsonl 2> gpurun_out/kernel_bench.err
tail -3 gpurun_out/kernel_bench.err
timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216.json 2>&1
FCAPP_TUNE="sweep_tiled=1,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled.json 2>&1
FCAPP_TUNE="sweep_tiled=2,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled_p2p.json 2>&1
FCAPP_TUNE="sweep_tiled=2,tile_ctas=3,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled_p2p_occ3.json 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_uvw_faces|k_uvw_rows|k_calcp_faces|k_grad_pass|k_rows_gather|k_tri_sweep' \
    -c 14 -o gpurun_out/prof_assembly python tools/simple_iter_bench.py 128 0 1 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_assembly.ncu-rep > gpurun_out/prof_assembly.txt 2>/dev/null
FCAPP_TUNE="sweep_tiled=1" timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile_sweep' -c 6 \
    -o gpurun_out/prof_tile_sweep python tools/simple_iter_bench.py 128 0 1 > gpurun_out/ncu_tile.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_tile_sweep.ncu-rep > gpurun_out/prof_tile_sweep.txt 2>/dev/null
FCAPP_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu \
    -k "tiled and poly and iccg" > gpurun_out/sanitizer_race_tile.log 2>&1
grep -h "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_*.log
ls -la gpurun_out | tail -30
