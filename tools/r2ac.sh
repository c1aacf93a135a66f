set -x
python tools/dgemm_probe.py 2>&1 | tee gpurun_out/r2af_dgemm.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fit.py -x -q > gpurun_out/r2af_pytest.log 2>&1; tail -5 gpurun_out/r2af_pytest.log | cut -c1-250
timeout 600 python tools/c4_probe.py 100000 chol::1 > gpurun_out/r2af_c4.log 2>&1; cut -c1-400 gpurun_out/r2af_c4.log
