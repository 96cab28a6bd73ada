"""key metrics + top stall lines of an .ncu-rep: python scripts/ncu_summary.py rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, r = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
for w in want:
    if w in h:
        print(f"{w:75s} {rows[1][h.index(w)]:12s} {r[h.index(w)]}")
for name in h:
    if "tensor" in name and "pct" in name and (".avg." in name or "realtime" in name) and float(r[h.index(name)] or 0) > 0:
        print(f"{name:75s} {r[h.index(name)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, rr in enumerate(rows) if "# Samples" in rr or "Sampling Data (All)" in rr)
h = rows[hi]; body = rows[hi + 1:]
scol = h.index("# Samples") if "# Samples" in h else h.index("Sampling Data (All)")
ix = {n: i for i, n in enumerate(h)}
def iv(x):
    try: return int(float(x))
    except Exception: return 0
tot = sum(iv(b[scol]) for b in body) or 1
print("total samples", tot, "lines", len(body))
top = sorted(range(len(body)), key=lambda i: -iv(body[i][scol]))[:ntop]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for i in sorted(top):
    b = body[i]
    dom = sorted(((iv(b[ix[n]]), n) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * iv(b[scol]) / tot:5.1f}% {b[ix['Source']].strip()[:100]:100s} {dom}")
