for t in 32 64 128 256; do
echo "== DFB_GN_CTAS=$t"
DFB_GN_CTAS=$t timeout 600 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'],d['roofline']['whole_step']['unet_step_ms'], d['roofline']['by_kind_ms_event_profile']['groupnorm'])"
done
