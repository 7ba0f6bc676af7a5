#!/usr/bin/env bash
# the driver's N = 1 commands: reference arm, then ours (with parity, config 3, config 5 at full size)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
t0=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
t1=$(date +%s); echo "reference arm: $((t1-t0)) s"
tail -c 1500 gpurun_out/bench_ref_n1.json; tail -3 gpurun_out/bench_ref_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
t2=$(date +%s); echo "our arm: $((t2-t1)) s"
tail -c 6500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
