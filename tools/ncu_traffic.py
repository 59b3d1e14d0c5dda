#!/usr/bin/env python
"""Extract per-launch DRAM traffic of the msda kernels from an .ncu-rep into profiles/ncu_traffic.json (read by bench.py
for roofline.traffic).   python tools/ncu_traffic.py gpurun_out/x/prof.ncu-rep <workload tag>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def val(r, name):
    i = hdr.index(name)
    v = float(r[i])
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


out = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = "fwd" if "msda_fwd" in name else "bwd" if "msda_bwd" in name else None
    if key is None:
        continue
    out[key] = {"kernel": name.split("(")[0].replace("void ", ""),
                "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])}
    out[key]["dram_bytes"] = out[key]["dram_bytes_read"] + out[key]["dram_bytes_write"]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
blob = json.load(open(path)) if os.path.exists(path) else {}
blob[tag] = {"source": rep, **out}
json.dump(blob, open(path, "w"), indent=1)
print(json.dumps(blob[tag], indent=1))
