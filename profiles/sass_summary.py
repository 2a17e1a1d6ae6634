"""Summarise the SASS of one kernel: opcode histogram per basic block (loops show up as backward branches).
usage: python profiles/sass_summary.py <lib.so> <substring of mangled name>"""
import re
import subprocess
import sys
from collections import Counter

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if key in f.split("\n", 1)[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        addr, text = int(m.group(1), 16), m.group(2).strip()
        toks = text.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ins.append((addr, op, text))
print(body.split("\n", 1)[0], "instructions:", len(ins))
# basic blocks delimited by branch targets / branches
targets = set()
for a, op, t in ins:
    if op.startswith("BRA") or op.startswith("BSSY") or op.startswith("CALL"):
        m = re.search(r"0x([0-9a-f]+)", t)
        if m:
            targets.add(int(m.group(1), 16))
blocks, cur = [], []
for a, op, t in ins:
    if a in targets and cur:
        blocks.append(cur)
        cur = []
    cur.append((a, op, t))
    if op.startswith(("BRA", "EXIT", "RET", "BRX")):
        blocks.append(cur)
        cur = []
if cur:
    blocks.append(cur)
for b in blocks:
    c = Counter(op.split(".")[0] for _, op, _ in b)
    last = b[-1]
    back = ""
    if last[1].startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", last[2])
        if m and int(m.group(1), 16) <= last[0]:
            back = f"  <-- LOOP back to {m.group(1)}"
    if len(b) >= 12 or back:
        top = ", ".join(f"{k}:{v}" for k, v in c.most_common(14))
        print(f"[{b[0][0]:05x}-{last[0]:05x}] n={len(b):4d} {top}{back}")
