timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-extras > gpurun_out/bench_v22.json 2> gpurun_out/bench_v22.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v22.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'], d['roofline']['by_kind_ms_event_profile'])"
timeout 300 python tools/trace_step.py 2 > gpurun_out/trace_v22_b2.log 2>&1; head -1 gpurun_out/trace_v22_b2.log
