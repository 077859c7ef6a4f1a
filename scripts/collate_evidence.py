#!/usr/bin/env python3
"""Copy the outputs of scripts/gpu_evidence.sh from gpurun_out/ (scratch) into profiles/ (tracked) under a prefix and
refresh profiles/traffic.json from the ncu capture:  python scripts/collate_evidence.py r1s3"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pre = sys.argv[1]
for src, dst in [("bench.log", "bench_c2.json"), ("bench_ref.log", "bench_reference_arm.json"), ("launches.csv", "launches.csv"),
                 ("time_c5_full.log", "c5_full_one_gpu.json"), ("pytest_gpu.log", "pytest_gpu.log")]:
    shutil.copy(os.path.join(G, src), os.path.join(P, f"{pre}_{dst}"))
out = {}
for line in open(os.path.join(G, "time_configs.log")):
    if " {" in line[:12]:
        n, js = line.split(" ", 1)
        out[n] = json.loads(js)
json.dump(out, open(os.path.join(P, f"{pre}_configs_all.json"), "w"), indent=1)
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")], input=raw, capture_output=True, text=True).stdout
open(os.path.join(P, f"{pre}_ncu_c2_kernels.txt"), "w").write(summ)
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
iK, iR, iW = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
uR, uW = rows[1][iR], rows[1][iW]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"fc_fast_r2c": "fast_r2c_N512", "fc_fused_axis": "fused_axis_N512", "fc_fast_c2r": "fast_c2r_N512"}
traffic = {}
for r in rows[2:]:
    for k, v in names.items():
        if k in r[iK]:
            traffic[v] = int(float(r[iR].replace(",", "")) * mult[uR] + float(r[iW].replace(",", "")) * mult[uW])
traffic["_source"] = f"ncu --set full --clock-control none, one launch each (profiles/{pre}_ncu_c2_kernels.txt); dram__bytes_read.sum + dram__bytes_write.sum"
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(traffic)
for n, v in out.items():
    print(n, round(v["ours_ms_best"], 4), [(k["kernel"], round(k["ms"] * 1e3, 1), round(k["gbs"]) if k["gbs"] else None) for k in v.get("kernels", [])])
