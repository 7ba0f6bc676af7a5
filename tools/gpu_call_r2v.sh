#!/usr/bin/env bash
# 2 GPUs: multi-rank parity with the eager x update (small meshes select it), and its A/B at the per-rank size of the
# 8-GPU run (136^3 on 2 ranks = 1.26 M cells per rank)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu_n2_eager.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu_n2_eager.log
grep "\[mgpu\]\|rc=" gpurun_out/mgpu_n2_eager.log | tail -12
for f in 0 1; do
FC_BENCH_N=136 FC_BENCH_NO_PARITY=1 FC_BENCH_NO_CONFIGS=1 FCAPP_TUNE="dpcg_eager=$f" timeout 300 $TR --master-port 2952$f bench.py --gpus 2 --steps 5 --warmup 3 --no-simple --no-cpu > gpurun_out/bench_n2_136_eager$f.json 2> gpurun_out/bench_n2_136_eager$f.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_136_eager$f.json").read().strip().splitlines()[-1])
print("eager=$f", d["value"], d["ms_per_step"], d["dpcg_iterations_each_step"], json.dumps(d["roofline"]["phases_us_per_iteration"]))
PY
done
