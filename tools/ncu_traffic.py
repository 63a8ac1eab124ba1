#!/usr/bin/env python
"""DRAM traffic per wideband sample of the receive kernels, from `ncu --set full` captures of ONE launch each:

    tools/ncu_traffic.py <samples per launch> analyzer_kernel=<rep> sync_kernel=<rep> packet_decode_kernel=<rep> > profiles/r02_traffic.json

(dram__bytes_read.sum + dram__bytes_write.sum of the captured launch) / (wideband samples that launch processed);
bench.py reads the JSON for roofline.traffic."""
import csv
import json
import subprocess
import sys

samples = float(sys.argv[1])
out = {"samples_per_launch": samples, "kernels": {}}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))

    def val(key):
        unit, v = d[key]
        v = float(v.replace(",", ""))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return v * scale
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out["kernels"][name] = {"kernel": d["Kernel Name"][1], "report": rep.split("/")[-1], "dram_bytes_read": rd, "dram_bytes_write": wr,
                            "duration_us": float(d["gpu__time_duration.sum"][1].replace(",", "")) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}[d["gpu__time_duration.sum"][0]],
                            "dram_bytes_per_sample": (rd + wr) / samples}
json.dump(out, sys.stdout, indent=1)
print()
