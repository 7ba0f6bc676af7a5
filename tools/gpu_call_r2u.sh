#!/usr/bin/env bash
# round 2: polyhedral (BCC-Voronoi) sweeps: tile bin width (FC_TILE_MIN_SHRINK) and hand-over mode, 2 x 128^3 cells
set -u
mkdir -p gpurun_out
for sh in 0 1 2 3; do
echo "== FC_TILE_MIN_SHRINK=$sh"
MESH=poly FC_TILE_MIN_SHRINK=$sh SWEEP_MODES="4:2,5:3" timeout 300 python tools/sweep_bench.py sweeps 128 2> gpurun_out/sweep_bench_poly.err | grep iccg | cut -c1-120,215-420
done
tail -3 gpurun_out/sweep_bench_poly.err
