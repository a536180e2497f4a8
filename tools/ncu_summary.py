#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
KEYS = [
 'gpu__time_duration.sum','sm__cycles_elapsed.max','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct','lts__t_sectors.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
 'l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
]
def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    stall = [k for k in hdr if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio')] or \
            [k for k in hdr if 'issue_stalled' in k and 'ratio' in k]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('==', name[:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); print(f'  {k:80s} {r[i]:>16s} {units[i]}')
        st = []
        for k in stall:
            i = hdr.index(k)
            try: st.append((float(r[i]), k))
            except ValueError: pass
        for v, k in sorted(st, reverse=True)[:8]:
            print(f'  stall {k.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""):40s} {v:8.2f} warps/issue')
main()
