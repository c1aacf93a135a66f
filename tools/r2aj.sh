set -x
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_pipeline.py tests/test_gpu_recognize.py tests/test_gpu_template.py -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras latency_b1,c5,c1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d.get('latency_b1'))); print(json.dumps(d.get('c5_video'))[:600])
"
