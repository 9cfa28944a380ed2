timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --no-extras > gpurun_out/bench_v24.json 2> gpurun_out/bench_v24.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v24.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'])"; done
DFB_DEBUG_SKIP=0 timeout 100 python tools/_epi_cost.py child
