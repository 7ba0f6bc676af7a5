#!/usr/bin/env bash
# round 2: ncu of the SpMV pipeline (stand-alone launch) and of the persistent kernel with one-byte column codes
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 3 -c 1 -o gpurun_out/prof_spmv_tma216 -f python tools/spmv_probe.py 216 4 > gpurun_out/ncu_spmv_tma.log 2>&1
tail -3 gpurun_out/ncu_spmv_tma.log
FCAPP_TUNE="ja_coded=1" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dpcg_persist -s 1 -c 1 -o gpurun_out/prof_persist216_coded -f python tools/spmv_probe.py 216 40 > gpurun_out/ncu_persist_coded.log 2>&1
tail -3 gpurun_out/ncu_persist_coded.log
FCAPP_TUNE="ja_coded=0" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dpcg_persist -s 1 -c 1 -o gpurun_out/prof_persist216_ja -f python tools/spmv_probe.py 216 40 > gpurun_out/ncu_persist_ja.log 2>&1
tail -3 gpurun_out/ncu_persist_ja.log
ls -la gpurun_out/*.ncu-rep | tail -5
