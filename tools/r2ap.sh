timeout 600 python -m pytest tests/test_gpu_recognize.py -x -q 2>&1 | tail -3
timeout 200 python tools/stream_probe.py stages 2>&1 | grep "us per batch\|stages" | grep -v "CTA 0" | cut -c1-260 | head -16
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('K=20', d['ms_per_step'], d['roofline']['frac'], d['serving_results_bit_identical_to_unpipelined'], d['pipeline_timeouts'], d['clocks']['sm_mhz'])
"
python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('K=200', d['ms_per_step'], d['roofline']['frac'], d['serving_results_bit_identical_to_unpipelined'])
"
