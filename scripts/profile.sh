#!/bin/bash
# ncu evidence for the bench numbers (run under gpurun, 1 GPU). Outputs under gpurun_out/.
# usage: scripts/profile.sh <workload> <tag>
set -x
WL=${1:-C4}; TAG=${2:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $BENCH > gpurun_out/launches_${WL}_${TAG}.log 2>&1
# the dominant kernel: the TMA-staged dense bulk scan (two launches well inside the first phase)
ncu --set full --clock-control none --import-source on -k regex:"k_scan_bulk_dense" -s 4 -c 2 -f -o gpurun_out/prof_bulkdense_${WL}_${TAG} python scripts/prof_run.py $WL max_loops=8 > gpurun_out/prof_bulkdense_${WL}_${TAG}.log 2>&1
# exact rounds of the constrained QEM phases: frontier scan (list kernel), candidate evaluation, commit
ncu --set full --clock-control none --import-source on -k regex:"k_evaluate|k_scan<|k_commit" -s 0 -c 6 -f -o gpurun_out/prof_exact_${WL}_${TAG} python scripts/prof_run.py $WL unconstrained_init=0 max_loops=2 > gpurun_out/prof_exact_${WL}_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
