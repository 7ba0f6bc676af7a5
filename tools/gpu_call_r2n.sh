#!/usr/bin/env bash
# 2 GPUs: multi-rank parity with the new defaults (k_tile_walk sweeps, fused-p DPCG), fused-p A/B at the per-rank size of
# the 8-GPU run (136^3 on 2 ranks = 1.26 M cells per rank)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for f in 0 1; do
FC_BENCH_N=136 FC_BENCH_NO_PARITY=1 FC_BENCH_NO_CONFIGS=1 FCAPP_TUNE="dpcg_fused=$f" timeout 300 $TR --master-port 2952$f bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_136_fused$f.json 2> gpurun_out/bench_n2_136_fused$f.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_136_fused$f.json").read().strip().splitlines()[-1])
print("fused=$f", d["value"], d["ms_per_step"], d["dpcg_iterations_each_step"], json.dumps(d["roofline"]["phases_us_per_iteration"]))
PY
done
FC_BENCH_N=64 FC_BENCH_POLY_N=32 timeout 300 $TR --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_small.json 2> gpurun_out/bench_n2_small.err
tail -c 1800 gpurun_out/bench_n2_small.json; tail -3 gpurun_out/bench_n2_small.err
