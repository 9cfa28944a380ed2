MET=smsp__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_mio_throttle.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_membar.sum
for skip in 0 2; do
DFB_DEBUG_SKIP=$skip timeout 300 ncu --metrics $MET --clock-control none -k regex:igemm -s 100 -c 4 --csv --log-file gpurun_out/ncu_epi_skip$skip.csv python tools/_epi_cost.py child > /dev/null 2>&1
done
