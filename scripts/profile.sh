#!/bin/bash
# ncu evidence for the bench numbers (run under gpurun, 1 GPU). Outputs under gpurun_out/.
# usage: scripts/profile.sh <workload> <tag>
set -x
WL=${1:-C2}; TAG=${2:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $BENCH > gpurun_out/launches_${WL}_${TAG}.log 2>&1
# the reassignment kernels: early bulk rounds and early exact rounds
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_bulk_evaluate|k_tile_filter" -s 9 -c 6 -f -o gpurun_out/prof_bulk_${WL}_${TAG} $BENCH > gpurun_out/prof_bulk_${WL}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_evaluate|k_commit" -s 130 -c 8 -f -o gpurun_out/prof_exact_${WL}_${TAG} $BENCH > gpurun_out/prof_exact_${WL}_${TAG}.log 2>&1
ls -la gpurun_out
