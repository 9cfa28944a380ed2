echo "== table"; timeout 200 python tools/microbench_convs.py 5 2>&1 | head -7
echo "== pair BN=128 (DFB_NO_TUNED so that splits==1 model plans are paired)"; DFB_PAIR=1 timeout 200 python tools/microbench_convs.py 5 2>&1 | head -7
echo "== slope"; timeout 100 python tools/_slope.py
