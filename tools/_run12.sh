timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "wide" 2>&1 | tail -5
echo "== B16 convs: table plans"; timeout 200 python tools/microbench_convs.py 5 2>&1 | head -7
echo "== B16 convs: wide tiles"; timeout 200 python - <<'PY'
import sys; sys.path.insert(0, ".")
import runpy
from diff_foley_b200 import _lib as L
L.lib().dfb_debug_igemm_force(256, 1)
sys.argv = ["x", "5"]
runpy.run_path("tools/microbench_convs.py")
PY
