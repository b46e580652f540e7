"""Summarise an .ncu-rep (ncu --set full) into the JSON kept under profiles/: one record per profiled launch with the
metrics the roofline discussion uses.  Usage: python tools/ncu_summary.py in.ncu-rep out.json "<command that was profiled>" [pairs_per_launch]"""
import csv, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active"]


def main():
    rep, out, command = sys.argv[1], sys.argv[2], sys.argv[3]
    pairs = int(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for i, name in enumerate(head):
        col.setdefault(name.split(".", 2)[-1] if name.count(".") >= 2 and name.split(".")[1] in ("TriageCompute",) else name, i)
        col.setdefault(name, i)
    kernels = []
    for r in body:
        rec = {"Kernel Name": r[head.index("Kernel Name")]}
        for m in WANT:
            i = col.get(m)
            if i is None:
                cands = [j for j, n in enumerate(head) if n.endswith(m)]
                i = cands[0] if cands else None
            if i is not None and i < len(r) and r[i] != "":
                rec[m] = f"{r[i]} {units[i]}".strip()
        kernels.append(rec)
    doc = {"command": command, "kernels": kernels}
    if pairs:
        doc["pairs_per_launch"] = pairs
    json.dump(doc, open(out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"][:50], k.get("gpu__time_duration.sum"), k.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
