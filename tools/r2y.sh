set -x
timeout 600 python -m pytest tests/test_gpu_recognize.py -x -q > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2y_bench20.json 2> gpurun_out/r2y_bench20.err
python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2y_bench200.json 2> gpurun_out/r2y_bench200.err
