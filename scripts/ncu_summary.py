import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','dram__throughput.avg.pct','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit','sm__throughput.avg.pct','l1tex__throughput.avg.pct','lts__throughput.avg.pct','sm__inst_executed_pipe_fp64.avg.pct','smsp__issue_active.avg.pct','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_bytes.sum ','l1tex__t_bytes.sum ','launch__grid_size','smsp__inst_executed.sum ','sm__inst_executed_pipe_alu','sm__inst_executed_pipe_fma','sm__inst_executed_pipe_lsu','smsp__average_warp','smsp__warps_issue_stalled','l1tex__data_pipe_lsu_wavefronts.sum ','l1tex__t_requests','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ','smsp__pcsamp']
for r in rows[2:]:
    print('==',r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i,h in enumerate(hdr):
        if any(w.strip() in h for w in want) and r[i] not in ('','0'):
            print(f'{h:90s} {units[i]:12s} {r[i]}')
