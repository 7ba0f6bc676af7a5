#!/usr/bin/env bash
# round 2, closing run on one GPU: the whole GPU suite (default + experimental), the driver's bench command, the ncu
# launch list of the bench command and a full capture of the dominant kernel (refreshes profiles/ncu_traffic.json)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
FCAPP_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_zz9_experimental.py -q -m gpu > gpurun_out/pytest_experimental.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_experimental.log
tail -3 gpurun_out/pytest_experimental.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench: $(( $(date +%s) - t0 )) s"; tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","dpcg_iterations_per_step")}, d["e2e"]["value"], d["clocks"])
print(json.dumps(d["roofline"])[:900])
print(json.dumps(d.get("parity"))[:400])
print(json.dumps(d.get("simple_iteration"))[:700])
print(json.dumps(d.get("other_solvers"))[:600])
PY
export FC_BENCH_NO_PARITY=1 FC_BENCH_NO_CONFIGS=1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-simple --no-cpu > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_bench.csv > gpurun_out/launches_bench.txt 2>/dev/null; head -12 gpurun_out/launches_bench.txt
timeout 500 ncu --set full --clock-control none -k regex:k_dpcg_persist -c 1 -o gpurun_out/prof_persist216_final -f \
    python bench.py --steps 1 --warmup 0 --no-simple --no-cpu > gpurun_out/ncu_persist216.log 2>&1
python tools/ncu_summary.py full gpurun_out/prof_persist216_final.ncu-rep > gpurun_out/prof_persist216_final.txt 2>/dev/null
grep "^kernel\|dram__bytes\|gpu__time_duration\|registers" gpurun_out/prof_persist216_final.txt
