timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== slope"; DFB_DEBUG_SKIP=0 timeout 100 python tools/_slope2.py child
echo "== warm/cold"; timeout 100 python tools/_warm_cold.py
echo "== convs B16"; timeout 200 python tools/microbench_convs.py 5 2>&1 | tail -12
timeout 600 python bench.py --no-extras > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err; tail -c 1500 gpurun_out/bench_v18.json
