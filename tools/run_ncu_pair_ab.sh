set -x
MET=lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg,lts__t_sectors_srcunit_tex_op_read.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
for k in conv ffo lin; do
for pair in 0 1; do
  DFB_PAIR=$pair timeout 120 python tools/one_gemm_for_ncu.py $k 2>&1 | tail -1
  DFB_PAIR=$pair timeout 300 ncu --metrics $MET --clock-control none -k regex:igemm -c 2 --csv --log-file gpurun_out/ncu_pair_${k}_$pair.csv python tools/one_gemm_for_ncu.py $k 1 > /dev/null 2>&1
done; done
