set -x
timeout 900 python tools/c4_probe.py 100000 > gpurun_out/r2z_c4.log 2>&1
cut -c1-1200 gpurun_out/r2z_c4.log
python tools/preprocess_probe.py > gpurun_out/r2z_pre.log 2>&1; cat gpurun_out/r2z_pre.log
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_fit.py -x -q 2>&1 | tail -3
