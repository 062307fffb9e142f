"""Record the DRAM traffic per launch of a kernel from an `ncu --set full` capture into profiles/ncu_traffic.json (read by bench.py
for `roofline.traffic`).     python tools/ncu_traffic.py gpurun_out/r02_attn.ncu-rep attn_fwd_kernel cfg3 profiles/r02_attn_ncu_full.txt"""
import csv
import io
import json
import os
import subprocess
import sys

rep, kernel, workload, source = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
rd, wr, n = 0.0, 0.0, 0
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    if kernel in d.get("Kernel Name", ""):
        rd += float(d["dram__bytes_read.sum"].replace(",", "")) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr += float(d["dram__bytes_write.sum"].replace(",", "")) * scale[units[hdr.index("dram__bytes_write.sum")]]
        n += 1
assert n, f"no launch of {kernel} in {rep}"
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
db.setdefault(kernel, {})[workload] = {"dram_bytes": (rd + wr) / n, "dram_bytes_read": rd / n, "dram_bytes_write": wr / n, "launches": n,
                                       "source": source}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(db[kernel][workload]))
