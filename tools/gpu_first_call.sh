#!/usr/bin/env bash
# First gpurun call of a round (1 GPU, about 20-25 minutes):
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_first_call.sh'
# Runs the GPU test suite, the per-operation roofline table, the bench line, the ncu launch list of the bench and one
# full ncu capture of the assembly / momentum / sweep kernels; everything lands in gpurun_out/ (copy what should be
# judged into profiles/).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
FCAPP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu -s > gpurun_out/pytest_experimental.log 2>&1
grep "sweeps\]\|passed\|failed" gpurun_out/pytest_experimental.log | tail -24
timeout 300 python tools/kernel_bench.py 216 10 > gpurun_out/kernel_bench_216.jsonl 2> gpurun_out/kernel_bench.err
timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216.json 2>&1
FCAPP_TUNE="sweep_tiled=1,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled.json 2>&1
FCAPP_TUNE="sweep_tiled=2,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled_p2p.json 2>&1
FCAPP_TUNE="sweep_tiled=2,tile_ctas=3,fused_grad=1" timeout 120 python tools/simple_iter_bench.py 216 2 3 > gpurun_out/simple_iter_216_tiled_p2p_occ3.json 2>&1
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_simple.csv \
    python tools/simple_iter_bench.py 128 1 1 > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_simple.csv > gpurun_out/launches_simple.txt 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_uvw_faces|k_uvw_rows|k_calcp_faces|k_grad_pass|k_rows_gather|k_tri_sweep|k_tile_sweep' \
    -c 12 -o gpurun_out/prof_assembly python tools/simple_iter_bench.py 128 0 1 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_assembly.ncu-rep > gpurun_out/prof_assembly.txt 2>/dev/null
ls -la gpurun_out | tail -20
# the tiled sweep kernel under ncu (it only runs when the tuning key is set)
FCAPP_TUNE="sweep_tiled=1" timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile_sweep' -c 6 \
    -o gpurun_out/prof_tile_sweep python tools/simple_iter_bench.py 128 0 1 > gpurun_out/ncu_tile.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_tile_sweep.ncu-rep > gpurun_out/prof_tile_sweep.txt 2>/dev/null
ls -la gpurun_out | tail -8
# memory / race checks of the kernels that have never run before (small meshes; a few minutes)
FCAPP_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu \
    -k "tiled and poly and iccg" > gpurun_out/sanitizer_race_tile.log 2>&1
FCAPP_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu \
    -k "(tiled and slab and bicgstab) or (fused and poly)" > gpurun_out/sanitizer_mem_experimental.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_zz2_piso.py tests/test_gpu_zz3_gradients.py -q -m gpu \
    -k "(without_krylov and skew) or (dispatcher and poly) or configured_gradients" > gpurun_out/sanitizer_mem_widened.log 2>&1
grep -h "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_*.log
