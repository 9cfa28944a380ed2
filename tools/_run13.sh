timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_cavp_gpu.py tests/test_vae_gpu.py -x -q -m gpu 2>&1 | tail -3
for w in 1 0; do
echo "== DFB_WIDE=$w"
DFB_WIDE=$w timeout 300 python bench.py --workload cavp --clips-per-gpu 4 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cavp',d['value'],d['e2e']['value'])"
DFB_WIDE=$w timeout 300 python tools/bench_vae.py 2>&1 | tail -3
done
timeout 600 python bench.py > gpurun_out/bench_v25.json 2> gpurun_out/bench_v25.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_v25.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'], d['config3'])"
timeout 300 python tools/trace_step.py 16 > gpurun_out/trace_v25_b16.log 2>&1; head -1 gpurun_out/trace_v25_b16.log
