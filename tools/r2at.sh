timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fit.py tests/test_gpu_dist.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras gram 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        g=json.loads(l)['gram']; print({k:g[k] for k in ('ms','ms_accumulating_call','ms_upper_triangle_only','store_equals_accumulate')})
"
python tools/c4_probe.py 100000 chol::1 2>&1 | grep "rep 1" | cut -c1-330
