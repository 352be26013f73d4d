#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from an `ncu --set full` report of k_stream<14,ENC,aligned> over 2^30 B:
per-launch DRAM bytes and L1/shared data-pipe wavefronts per 32 blocks (what bench.py quotes as
roofline.traffic and bound_smem_lookup.lsu_wavefronts_per_32_blocks, with this file as their source).
usage: tools/ncu_json.py <report.ncu-rep | raw page .csv> <n_bytes> <source note>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, n_bytes, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
if rep.endswith(".csv"):   # the raw page, already exported (`ncu -i x.ncu-rep --page raw --csv`)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
hdr, units, data = rows[0], rows[1], rows[2:]
d = [r for r in data if "k_stream" in r[hdr.index("Kernel Name")]][-1]


def val(k):
    i = hdr.index(k)
    v = float(d[i].replace(",", ""))
    u = units[i].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(u, 1.0)


# ncu's raw page has the shared-memory wavefronts as a sum and the whole pipe only as a percentage: the total
# is shared wavefronts + 4 per fully coalesced 128-bit global request (512 B = 4 wavefronts of 128 B), and
# the percentage x elapsed cycles x SMs is kept beside it as a cross-check
shared = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
glob = 4 * (val("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum") + val("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"))
pct = val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
sms = int(val("launch__grid_size"))   # one persistent CTA per SM
from_pct = pct / 100.0 * val("sm__cycles_elapsed.avg") * sms
rows = n_bytes / 16 / 32
res = {"source": note, "kernel": d[hdr.index("Kernel Name")], "n_bytes": n_bytes,
       "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
       "lsu_wavefronts_shared": int(shared), "lsu_wavefronts_global": int(glob),
       "lsu_wavefronts_per_32_blocks": round((shared + glob) / rows, 2),
       "lsu_wavefronts_per_32_blocks_from_pct_of_peak": round(from_pct / rows, 2),
       "lsu_wavefronts_pct_of_peak": round(pct, 2),
       "gpu_time_ms": round(val("gpu__time_duration.sum"), 4)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(root, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
