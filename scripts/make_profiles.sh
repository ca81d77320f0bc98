#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (tracked): markdown summaries of the ncu captures of a round.
# usage: scripts/make_profiles.sh <tag of the gpurun_out files> <name prefix under profiles/>     e.g. r02p r02
TAG=${1:?tag}; OUT=${2:-r02}
cd "$(dirname "$0")/.."
[ -f gpurun_out/launches_C4_${TAG}.csv ] && python scripts/ncu_summary.py launches gpurun_out/launches_C4_${TAG}.csv "Launch list of one bench.py step, C4 (${TAG}; includes the mesh set-up and the initial fill)" > profiles/${OUT}_launches_C4.md
for k in bulkdense split splitlate rest; do
  [ -f gpurun_out/prof_${k}_C4_${TAG}.ncu-rep ] && python scripts/ncu_summary.py raw gpurun_out/prof_${k}_C4_${TAG}.ncu-rep "ncu --set full: ${k} kernels, C4 (${TAG})" > profiles/${OUT}_full_${k}_C4.md
done
# SASS evidence: the TMA bulk-copy engine and its mbarrier in the streaming kernels
for fn in k_scan_classify k_scan_bulk_dense; do
  cuobjdump -sass acvd_b200/libacvd_b200.so 2>/dev/null | awk -v fn="$fn" '/Function : /{f = index($0, fn) > 0; if (f) print} f && /UBLKCP|SYNCS|CCTL|ATOMS|LDG|LDS|STG|RED/ {print}' | awk '{ $1=""; print }' | sed 's#/\*[0-9a-f]*\*/##g' | sort | uniq -c | sort -rn | awk '$1 > 0' > /tmp/sass_$fn.txt
done
{
  echo "# SASS of the streaming kernels (cuobjdump -sass acvd_b200/libacvd_b200.so, memory / TMA / mbarrier instructions, counts over all instantiations)"
  echo; echo '`UBLKCP.S.G` = cp.async.bulk global -> shared (TMA engine, non-tensor form), `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` = mbarrier expect_tx / try_wait, `CCTL.E.PF1` = prefetch.global.L1'; echo
  for fn in k_scan_classify k_scan_bulk_dense; do echo "## $fn"; echo '```'; grep -E "UBLKCP|SYNCS|CCTL" /tmp/sass_$fn.txt | head -30; echo '```'; done
} > profiles/${OUT}_sass_streaming_kernels.md
ls -la profiles | tail -8
