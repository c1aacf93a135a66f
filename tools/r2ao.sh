timeout 900 python -m pytest tests/test_gpu_recognize.py -x -q -k "band or small_gallery or few" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras shipped_shapes 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v['us'],1), v['launches_per_call']) for k,v in d['shipped_shapes'].items() if isinstance(v,dict)})
"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2ao_launches.csv python tools/models_probe.py > gpurun_out/r2ao_models.log 2>&1; tail -5 gpurun_out/r2ao_models.log | cut -c1-200
