#!/bin/bash
# ncu evidence for the bench numbers (run under gpurun, 1 GPU). Outputs under gpurun_out/.
# usage: scripts/profile.sh <workload> <tag>
set -x
WL=${1:-C2}; TAG=${2:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $BENCH > gpurun_out/launches_${WL}_${TAG}.log 2>&1
# dense bulk rounds: the fused scan (k_scan<6,true>); first exact rounds: k_scan<6,false>, k_evaluate, k_commit
ncu --set full --clock-control none --import-source on -k regex:"k_scan" -s 4 -c 2 -f -o gpurun_out/prof_bulkscan_${WL}_${TAG} $BENCH > gpurun_out/prof_bulkscan_${WL}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_evaluate|k_scan<\(int\)6, \(bool\)0>" -s 2 -c 4 -f -o gpurun_out/prof_exact_${WL}_${TAG} $BENCH > gpurun_out/prof_exact_${WL}_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
