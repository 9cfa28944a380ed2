timeout 600 python bench.py --no-extras > gpurun_out/bench_v20.json 2> gpurun_out/bench_v20.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v20.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'])"
timeout 300 python tools/trace_step.py 2 > gpurun_out/trace_v20_b2.log 2>&1; head -1 gpurun_out/trace_v20_b2.log
timeout 300 python tools/trace_step.py 16 > gpurun_out/trace_v20_b16.log 2>&1; head -1 gpurun_out/trace_v20_b16.log
