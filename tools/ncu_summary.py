#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (ncu -i ... --page raw --csv) as a markdown table.
usage: tools/ncu_summary.py <report.ncu-rep> <kernel-name substring> [title]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'smsp__cycles_active.avg']

rep, pat = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else pat
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index('Kernel Name')
sel = [r for r in data if pat in r[ki]]
if not sel:
    sys.exit("no kernel matching %r" % pat)
d = sel[-1]
print("# %s\n" % title)
print("Kernel: `%s`  (report %s, %d matching launches, last one shown)\n" % (d[ki], rep.split('/')[-1], len(sel)))
print("| metric | value | unit |\n|---|---|---|")
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print("| %s | %s | %s |" % (k, d[i], units[i]))
