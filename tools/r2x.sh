set -x
timeout 600 python -m pytest tests/test_gpu_recognize.py -x -q > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
timeout 300 python tools/stream_probe.py > gpurun_out/r2x_probe.log 2>&1
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench20.json 2> gpurun_out/r2x_bench20.err
grep -v "CTA 0" gpurun_out/r2x_probe.log | cut -c1-700
