"""Per-CUDA-source-line executed instructions and stall samples of one kernel (read here, no GPU).

usage: ncu_lines.py <rep> <nvdisasm -g -c output of the same build> <kernel symbol substring> [top N]
The .ncu-rep's SASS page and nvdisasm list the kernel's instructions in the same order; they are zipped by index."""
import csv, subprocess, sys, io, re, collections
rep, sass, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- line of every instruction, from nvdisasm
lines_of = []; cur = None; inside = False
for ln in open(sass):
    if ln.startswith(".text."):
        inside = sym in ln; continue
    if ln.startswith("//-----") or ln.lstrip().startswith(".section"):
        if inside and lines_of and ".text." in ln and sym not in ln: inside = False
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
if len(data) != len(lines_of):
    print(f"warning: {len(data)} instructions in the report vs {len(lines_of)} in the disassembly", file=sys.stderr)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: dict(n=0, ex=0, smp=0, st=collections.Counter()))
for r, key in zip(data, lines_of):
    a = agg[key]
    a["n"] += 1
    try: a["ex"] += int(r[idx["Instructions Executed"]])
    except Exception: pass
    a["smp"] += int(r[idx["# Samples"]] or 0)
    for k in stalls:
        try: a["st"][k[6:]] += int(r[idx[k]])
        except Exception: pass
tot_ex = sum(a["ex"] for a in agg.values()) or 1; tot_s = sum(a["smp"] for a in agg.values()) or 1
srcs = {}
print(f"{'file:line':26s}{'static':>7s}{'exec%':>7s}{'smp%':>7s}  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    f, l = key if key else ("?", 0)
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/demovlp_b200/csrc/{f}").read().split("\n")
        except Exception: srcs[f] = []
    text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
    st = ",".join(f"{k}={v}" for k, v in a["st"].most_common(3))
    print(f"{f+':'+str(l):26s}{a['n']:7d}{100*a['ex']/tot_ex:7.1f}{100*a['smp']/tot_s:7.1f}  {st} | {text}")
