"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall-reason totals + hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, val = rows[0], rows[-1]
print("kernel:", dict(zip(hdr, val)).get("Kernel Name", "?")[:100], "(last launch of the report)")
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg,", "sm__inst_executed.sum.per_cycle_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum,", "dram__bytes_write.sum,", "launch__registers_per_thread,",
        "lts__t_bytes.sum,", "smsp__inst_executed.sum,", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_op_red.sum,", "sm__warps_active.avg.pct_of_peak_sustained_active",
        # the shared-memory data pipe: LSU side (LDS / STS / SHFL) and tensor-core operand fetch
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, v in zip(hdr, val):
    if any((h + ",").startswith(w) or h == w.rstrip(",") for w in want):
        print(f"{h} = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# a report may hold several kernels: keep the last one's instruction table
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[starts[-1]:] if starts else rows
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        try: tot[s] += int(r[idx[s]])
        except Exception: pass
T = sum(tot.values()) or 1
print("SASS instructions:", len(data), "(~%d KB)" % (len(data) * 16 // 1024))
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:26s}{v:8d} {100 * v / T:5.1f}%")
si = idx["# Samples"]
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    print(r[si].rjust(6), r[idx["Source"]][:80], {s[6:]: r[idx[s]] for s in stalls if r[idx[s]] not in ("0", "")})
