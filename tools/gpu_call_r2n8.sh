#!/usr/bin/env bash
# 8 GPUs, closing run: the bench command without the full-size oracle solve / config 5 (those ran in the earlier 8-GPU
# call, profiles/r02_bench_n8.json), eager x update on (default at 1.26 M rows per rank) and off, and the 8-rank parity check
set -u
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export FC_BENCH_NO_PARITY=1 FC_BENCH_NO_CONFIGS=1
for e in 2 0; do
FCAPP_TUNE="dpcg_eager=$e" timeout 400 $TR --master-port 2956$e bench.py --gpus $N --steps 20 --warmup 5 --no-simple --no-cpu > gpurun_out/bench_n8_eager$e.json 2> gpurun_out/bench_n8_eager$e.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n8_eager$e.json").read().strip().splitlines()[-1])
print("eager=$e", d["value"], d["ms_per_step"], d["dpcg_iterations_each_step"][:3], d["e2e"]["value"], json.dumps(d["roofline"]["phases_us_per_iteration"]), d["clocks"])
PY
done
unset FC_BENCH_NO_PARITY FC_BENCH_NO_CONFIGS
MGPU_SECTIONS=core timeout 500 $TR --master-port 29571 tests/mgpu_check.py > gpurun_out/mgpu_n8_final.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n8_final.log
grep "\[mgpu\]\|rc=" gpurun_out/mgpu_n8_final.log | tail -8
