timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== warm/cold a32"; timeout 100 python tools/_warm_cold.py
echo "== warm/cold no a32"; DFB_A32=0 timeout 100 python tools/_warm_cold.py
timeout 600 python bench.py --no-extras > gpurun_out/bench_v19.json 2> gpurun_out/bench_v19.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v19.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'])"
