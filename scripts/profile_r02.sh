#!/bin/bash
# ncu evidence for the bench numbers (run under gpurun, 1 GPU).  Outputs under gpurun_out/: the launch list (csv) and, for
# every --set full capture, the markdown summary (scripts/ncu_summary.py) and the per-SASS-instruction source page (csv);
# the .ncu-rep files themselves are removed on the box (gpurun brings back at most 64 MiB).
# usage: scripts/profile_r02.sh <workload> <tag>
WL=${1:-C4}; TAG=${2:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0"
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv $BENCH > gpurun_out/launches_${WL}_${TAG}.log 2>&1
capture() {   # name, kernel regex, skip, count, prof_run args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local rep=gpurun_out/prof_${name}_${WL}_${TAG}
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $rep python scripts/prof_run.py $WL "$@" > $rep.log 2>&1
  python scripts/ncu_summary.py raw $rep.ncu-rep "ncu --set full: $name, $WL ($TAG); prof_run.py $WL $*" > $rep.md 2>> $rep.log
  ncu -i $rep.ncu-rep --page source --csv > $rep.source.csv 2>> $rep.log
  rm -f $rep.ncu-rep
}
# the dominant kernels: the split dense bulk scan (k_scan_classify + k_bulk_decide), two rounds early in the first phase and one late in it
capture split "k_scan_classify|k_bulk_decide" 8 4 max_loops=8
capture splitlate "k_scan_classify|k_bulk_decide" 100 2 max_loops=60
# the cluster pass (statistics + connectivity), the bulk commit, the opening-round evaluation
capture rest "k_cluster_pass|k_evaluate|k_bulk_commit|k_bulk_evaluate" 0 8
ls -la gpurun_out | tail -12
