#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from an `ncu --set full` report of k_stream<14,ENC,aligned> over 2^30 B:
per-launch DRAM bytes and L1/shared data-pipe wavefronts per 32 blocks (what bench.py quotes as
roofline.traffic and bound_smem_lookup.lsu_wavefronts_per_32_blocks, with this file as their source).
usage: tools/ncu_json.py <report.ncu-rep> <n_bytes> <source note>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, n_bytes, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
hdr, units, data = rows[0], rows[1], rows[2:]
d = [r for r in data if "k_stream" in r[hdr.index("Kernel Name")]][-1]


def val(k):
    i = hdr.index(k)
    v = float(d[i].replace(",", ""))
    u = units[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(u, 1.0)


wf = val("l1tex__data_pipe_lsu_wavefronts.sum")
res = {"source": note, "kernel": d[hdr.index("Kernel Name")], "n_bytes": n_bytes,
       "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
       "lsu_wavefronts": int(wf), "lsu_wavefronts_per_32_blocks": round(wf / (n_bytes / 16 / 32), 2),
       "lsu_wavefronts_pct_of_peak": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
       "gpu_time_ms": val("gpu__time_duration.sum") / 1e6 if "gpu__time_duration.sum" in hdr else None}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
