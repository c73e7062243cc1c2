import csv, sys, subprocess
rep=sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; units = rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum']
want += [h for h in hdr if 'average_warps_issue_stalled' in h]
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')][:18])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            try:
                v=float(r[i])
                if 'stalled' in w and v < 0.5: continue
            except: pass
            print('  ', w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio',''), '=', r[i], units[i])
