"""Summarise ncu outputs (launch-list csv or .ncu-rep raw page) into markdown for profiles/."""
import collections, csv, subprocess, sys

def launches(path, title):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0].replace("void ", "")
        if "cub::" in name: name = "cub::" + name.split("cub::")[1].split("<")[0]
        try: val = float(r[mv].replace(",", ""))
        except ValueError: continue
        agg[name][0] += 1; agg[name][1] += val
    tot = sum(v[1] for v in agg.values())
    out = [f"# {title}", "", "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare shares.", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]/1e6:.2f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]/1e3:.1f} |")
    return "\n".join(out) + "\n"

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_active_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]

def raw(path, title):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# {title}", "", "`ncu --set full --clock-control none --import-source on`", ""]
    for r in rows[2:]:
        out.append(f"## {r[idx['Kernel Name']].split('(')[0]}  (id {r[idx['ID']]})")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for w in WANT:
            if w in idx: out.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
        out.append("")
    return "\n".join(out) + "\n"

if __name__ == "__main__":
    mode, path, title = sys.argv[1], sys.argv[2], sys.argv[3]
    sys.stdout.write(launches(path, title) if mode == "launches" else raw(path, title))
