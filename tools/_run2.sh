timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== rank-minimal maps"; DFB_DEBUG_SKIP=0 timeout 100 python tools/_slope2.py child
echo "== 5-D maps"; DFB_TMA_RANK5=1 DFB_DEBUG_SKIP=0 timeout 100 python tools/_slope2.py child
echo "== rank-minimal, no MMA no W"; DFB_DEBUG_SKIP=160 timeout 100 python tools/_slope2.py child
echo "== warm/cold rank-minimal"; timeout 100 python tools/_warm_cold.py
echo "== warm/cold 5-D"; DFB_TMA_RANK5=1 timeout 100 python tools/_warm_cold.py
