#!/usr/bin/env bash
# Round 2, third call (1 GPU): value-as-flag sweeps, L2 keep, ncu of the persistent kernel at 108^3 and of the calcp
# assembly kernels, the N=1 bench line with the full-size oracle parity.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
FCAPP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu -s -k "tiled-vf" > gpurun_out/pytest_vf.log 2>&1
grep "sweeps\]\|passed\|failed\|Error\|error" gpurun_out/pytest_vf.log | tail -12
timeout 400 python tools/sweep_bench.py sweeps 216 > gpurun_out/sweep_bench_216.jsonl 2> gpurun_out/sweep_bench.err
cut -c1-330 gpurun_out/sweep_bench_216.jsonl; tail -3 gpurun_out/sweep_bench.err
timeout 200 python tools/sweep_bench.py l2 108 136 > gpurun_out/l2_keep.jsonl 2> gpurun_out/l2_keep.err
cat gpurun_out/l2_keep.jsonl; tail -3 gpurun_out/l2_keep.err
for keep in 0 1; do
FCAPP_TUNE="l2_keep=$keep" timeout 200 ncu --set full --clock-control none -k regex:k_dpcg_persist -c 1 -o gpurun_out/prof_persist108_keep$keep \
    python bench.py --n 108 --steps 1 --warmup 1 --no-parity --no-cpu --no-simple > gpurun_out/ncu_persist108_keep$keep.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_persist108_keep$keep.ncu-rep > gpurun_out/prof_persist108_keep$keep.txt 2>/dev/null
grep "dram__bytes\|gpu__time_duration\|lts__t_sector_hit" gpurun_out/prof_persist108_keep$keep.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_calcp_faces|k_rows_gather|k_grad_passn|k_flux_correct|k_cell_correct' -c 8 \
    -o gpurun_out/prof_calcp_asm python bench.py --n 128 --steps 1 --warmup 0 --no-parity --no-cpu --no-simple > gpurun_out/ncu_calcp.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_calcp_asm.ncu-rep > gpurun_out/prof_calcp_asm.txt 2>/dev/null
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
ls -la gpurun_out | tail -25
