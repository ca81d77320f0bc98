#!/bin/bash
# ncu evidence for the bench numbers (run under gpurun, 1 GPU). Outputs under gpurun_out/.
# usage: scripts/profile_r02.sh <workload> <tag>
WL=${1:-C4}; TAG=${2:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $BENCH > gpurun_out/launches_${WL}_${TAG}.log 2>&1
# the dominant kernel: the TMA-staged dense bulk scan (two launches well inside the first phase)
ncu --set full --clock-control none --import-source on -k regex:"k_scan_bulk_dense" -s 4 -c 2 -f -o gpurun_out/prof_bulkdense_${WL}_${TAG} python scripts/prof_run.py $WL max_loops=8 > gpurun_out/prof_bulkdense_${WL}_${TAG}.log 2>&1
# the cluster pass (statistics + connectivity) and the persistent sparse-round kernel
ncu --set full --clock-control none --import-source on -k regex:"k_cluster_pass|k_sparse|k_evaluate|k_bulk_commit" -s 0 -c 8 -f -o gpurun_out/prof_rest_${WL}_${TAG} python scripts/prof_run.py $WL > gpurun_out/prof_rest_${WL}_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
