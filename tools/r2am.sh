timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fit.py tests/test_gpu_dist.py -x -q 2>&1 | tail -15 | cut -c1-200
python tools/gram_probe.py 12500,10000,1 100000,10000,1 2>&1 | tail -2
EF_GRAM_TRANSPOSE=1 python tools/gram_probe.py 12500,10000,1 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras gram 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): print(json.dumps(json.loads(l)['gram'])[:700])
"
