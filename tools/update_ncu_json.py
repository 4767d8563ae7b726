"""Refresh one workload's entry of profiles/ncu_dram_traffic_r1.json from an `ncu --page raw --csv` export
(tools/ncu_capture.sh writes gpurun_out/raw_<tag>.csv; captures are 2 views per launch).
usage: python tools/update_ncu_json.py <workload> <raw.csv> <summary file name under profiles/> [views_per_launch]"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
workload, raw, summary = sys.argv[1], sys.argv[2], sys.argv[3]
views = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def nbytes(key):
    u, v = d[key]
    return float(v.replace(",", "")) * scale[u]


dram = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
inst = float(d["smsp__inst_executed.sum"][1].replace(",", ""))
path = ROOT / "profiles" / "ncu_dram_traffic_r1.json"
j = json.load(open(path))
e = j.setdefault(workload, {})
e.update({"dram_bytes_per_view": int(dram / views), "warp_inst_per_view": int(inst / views), "capture": f"profiles/{summary}",
          "kernel": d["Kernel Name"][1]})
json.dump(j, open(path, "w"), indent=1)
print(workload, e)
