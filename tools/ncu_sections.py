"""Split a warp-specialised kernel's SASS at its USETMAXREG instructions and report, per role section,
static size, executed warp-instructions, stall samples and opcode mix (read here, no GPU)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
secs = []; cur = None
for r in data:
    s = r[idx["Source"]].strip()
    if cur is None or "USETMAXREG" in s:
        cur = dict(name=s if "USETMAXREG" in s else "prologue", n=0, ex=0, samples=0, ops=collections.Counter(), st=collections.Counter())
        secs.append(cur)
    try: ex = int(r[idx["Instructions Executed"]])
    except Exception: ex = 0
    cur["n"] += 1; cur["ex"] += ex; cur["samples"] += int(r[idx["# Samples"]] or 0)
    toks = s.split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    op = toks[0] if toks else "?"
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDL", "STL", "MUFU", "SHFL", "LDTM", "REDG", "SYNCS") else op.split(".")[0]
    cur["ops"][op] += ex
    for k in stalls:
        try: cur["st"][k[6:]] += int(r[idx[k]])
        except Exception: pass
for s in secs:
    print(f"== {s['name'][:60]}: static {s['n']} instr, executed {s['ex']}, samples {s['samples']}")
    print("   ops:", ", ".join(f"{o}={100*n/max(s['ex'],1):.1f}%" for o, n in s["ops"].most_common(16)))
    tot = sum(s["st"].values()) or 1
    print("   stalls:", ", ".join(f"{o}={100*n/tot:.0f}%" for o, n in s["st"].most_common(7)))
