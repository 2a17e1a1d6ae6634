"""Digest of an ncu report exported with --page raw --csv and --page source --csv.
usage: python profiles/ncu_digest.py raw.csv src.csv [n_residues]"""
import csv
import sys
from collections import Counter

raw, src = sys.argv[1], sys.argv[2]
nres = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.05:
        print(f"{h:90s} {u:12s} {v}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
ops, samp = Counter(), Counter()
for r in data:
    t = r[ix["Source"]].strip().split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += f(r, "Instructions Executed")
    samp[op] += f(r, "# Samples")
print(f"total warp instructions {tot_i:.4g}" + (f" = {tot_i / nres:.1f} per residue" if nres else ""))
for op, c in ops.most_common(16):
    print(f"  {op:10s} {100 * c / tot_i:5.1f}% of instructions, {100 * samp[op] / tot_s:5.1f}% of stall samples" + (f", {c / nres:6.1f}/residue" if nres else ""))
# stall columns
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = Counter()
for r in data:
    for h in stall_cols:
        tot[h] += f(r, h)
print("stall samples by reason:", ", ".join(f"{k[6:]}={int(v)}" for k, v in tot.most_common(8)))
print("top stalled instructions:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:14]:
    reasons = sorted(((f(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  {int(f(r, '# Samples')):7d}  {r[ix['Source']].strip()[:60]:60s} {reasons}")
