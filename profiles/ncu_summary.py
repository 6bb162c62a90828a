import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
hdr=r[0]; units=r[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max','launch__grid_size','launch__block_size','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__maximum_warps_per_active_cycle_pct','launch__shared_mem_per_block_dynamic','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active']
idx=[i for i,h in enumerate(hdr) if h in want or (h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'))]
ki=hdr.index('Kernel Name')
for row in r[2:]:
    print('---', row[ki][:90])
    for i in idx: print('   ',hdr[i], '=', row[i], units[i])
