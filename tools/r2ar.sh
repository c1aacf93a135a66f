timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fit.py -x -q 2>&1 | tail -4
python tools/jacobi_probe.py 321 512 590 640 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras fit 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'): print(json.dumps(json.loads(l)['fit'])[:900])
"
