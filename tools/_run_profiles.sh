set -x
mkdir -p gpurun_out
V=$1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_$V.json 2> gpurun_out/r2_bench_n1_$V.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
timeout 300 python tools/trace_step.py 2 > gpurun_out/r2_trace_unet_fwd_beff2_$V.log 2>&1
timeout 300 python tools/trace_step.py 16 > gpurun_out/r2_trace_unet_fwd_beff16_$V.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_ncu_launches_unet_fwd_beff2_$V.csv python tools/profile_step.py 2 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:igemm --csv --log-file gpurun_out/r2_ncu_igemm_dram_beff2_$V.csv python tools/profile_step.py 2 > /dev/null 2>&1
timeout 300 python bench.py --workload cavp --clips-per-gpu 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_cavp_$V.json 2> gpurun_out/r2_bench_cavp_$V.err
tail -2 gpurun_out/r2_bench_n1_$V.err gpurun_out/r2_bench_cavp_$V.err
ls -la gpurun_out | tail -12
