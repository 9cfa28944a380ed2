for i in 1 2; do timeout 600 python bench.py --no-extras > gpurun_out/bench_v23.json 2> gpurun_out/bench_v23.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v23.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'])"; done
