"""Scratch: print key metrics for every kernel in an .ncu-rep."""
import csv, subprocess, sys
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg.per_second']
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units=rows[0],rows[1]
    for r in rows[2:]:
        d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
        print("=====",d['Kernel Name'][:100])
        for k in keys:
            if k in d: print(f"  {k:75s} {d[k]:>14s} {u[k]}")
        st={h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):float(d[h]) for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')}
        print("  stalls:", ", ".join(f"{k}={v:.2f}" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:7]))
