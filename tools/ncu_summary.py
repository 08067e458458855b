import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','launch__shared_mem_per_block_dynamic','l1tex__t_sector_hit_rate.pct','smsp__cycles_active.avg','sm__cycles_elapsed.max']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr: print(' ',w, '=', r[hdr.index(w)], units[hdr.index(w)])
