set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
ncu --query-metrics 2>/dev/null | grep -i -E 'tensor|utc|tmem|pipe_tc' > gpurun_out/r2_ncu_metrics_tensor.txt
wc -l gpurun_out/r2_ncu_metrics_tensor.txt
timeout 300 python tools/microbench_attention.py 2 > gpurun_out/r2_attn_base_b2.log 2>&1
timeout 300 python tools/microbench_attention.py 16 > gpurun_out/r2_attn_base_b16.log 2>&1
cat gpurun_out/r2_attn_base_b2.log gpurun_out/r2_attn_base_b16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 2 -c 2 -o gpurun_out/r2_prof_attention_base python tools/microbench_attention.py 16 2 > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
timeout 300 python tools/trace_step.py 2 > gpurun_out/r2_trace_base_beff2.log 2>&1
head -3 gpurun_out/r2_trace_base_beff2.log; tail -12 gpurun_out/r2_trace_base_beff2.log
timeout 300 python tools/profile_ops.py 16 > gpurun_out/r2_profile_base_beff16.log 2>&1
head -40 gpurun_out/r2_profile_base_beff16.log
