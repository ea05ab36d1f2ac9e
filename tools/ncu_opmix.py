"""Opcode mix of a kernel in an .ncu-rep, weighted by executed warp-instructions (read here, no GPU)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0
for r in data:
    try: n = int(r[idx["Instructions Executed"]])
    except Exception: continue
    s = r[idx["Source"]].strip()
    toks = s.split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    op = toks[0].split(".")[0] if toks else "?"
    if op in ("LDS", "STS", "LDL", "STL", "LDG", "STG", "RED", "SHFL", "MUFU", "F2FP", "LDTM", "BAR", "SYNCS"):
        op = ".".join(toks[0].split(".")[:2])
    ops[op] += n; tot += n
print("total warp-instructions:", tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(f"  {op:18s}{n:12d} {100*n/tot:5.1f}%")
