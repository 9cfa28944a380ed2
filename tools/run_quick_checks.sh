# ops + UNet parity tests and two --no-extras bench lines (the quick A/B used while tuning kernels)
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['whole_step']['unet_step_ms'])"; done
