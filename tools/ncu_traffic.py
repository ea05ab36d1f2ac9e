"""DRAM bytes per launch of the pair / gradient kernels from `ncu --set full` captures (read here, no GPU):
writes profiles/ncu_traffic.json, which bench.py copies into roofline.traffic for the dominant kernel.

usage: ncu_traffic.py fwd=<rep> bwd_pair=<rep> grad_gemm=<rep> > profiles/ncu_traffic.json"""
import csv, io, json, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME_US = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
out = {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    rd = wr = dur = 0.0
    names = []
    for r in launches:            # a report may hold several launches (e.g. both gradient GEMM modes): add them up
        rd += float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr += float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
        dur += float(r[col["gpu__time_duration.sum"]]) * TIME_US.get(units[col["gpu__time_duration.sum"]], 1.0)
        names.append(r[col["Kernel Name"]].split("(")[0])
    out[key] = {"kernels": names, "dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                "ncu_duration_us": round(dur, 1),
                "shape": "config 2: Bv=Bt=128, R=80, W=32, d=256 (16384 pairs per launch)",
                "source": "ncu --set full --clock-control none, tools/run_fwd.py 128 bwd (profiles/r1_*_ncu.txt)"}
print(json.dumps(out, indent=1))
