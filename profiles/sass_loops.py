"""Opcode stream of the hot loops of an ncu source page (csv): python profiles/sass_loops.py src.csv [lo hi]
lo/hi = execution-count window relative to the maximum (default 0.4 0.6 picks loops that run once per position)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.4, 0.6)
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
ex = [float(r[ix['Instructions Executed']]) for r in data]
mx = max(ex)
sel = [i for i, e in enumerate(ex) if lo * mx < e < hi * mx]
groups, cur = [], [sel[0]]
for i in sel[1:]:
    if i - cur[-1] > 60:
        groups.append(cur)
        cur = [i]
    else:
        cur.append(i)
groups.append(cur)
tot_all = sum(float(r[ix['# Samples']]) for r in data)
for g in groups:
    ops, tot = [], 0
    for i in range(g[0], g[-1] + 1):
        s = data[i][ix['Source']].strip().split()
        op = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
        ops.append(op)
        tot += float(data[i][ix['# Samples']])
    print(f'--- loop at {g[0]}..{g[-1]}: {len(ops)} instructions, {ops.count("DMMA")} DMMA, {100 * tot / tot_all:.1f}% of all stall samples')
    out, prev, n = [], None, 0
    for op in ops:
        if op == prev:
            n += 1
        else:
            if prev:
                out.append(f'{prev}x{n}' if n > 1 else prev)
            prev, n = op, 1
    out.append(f'{prev}x{n}' if n > 1 else prev)
    print(' '.join(out))
