set -x
timeout 600 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_fit.py -x -q > gpurun_out/r2ac_pytest.log 2>&1; tail -15 gpurun_out/r2ac_pytest.log | cut -c1-250
timeout 600 python tools/c4_probe.py 100000 chol::1 > gpurun_out/r2ac_c4.log 2>&1; cut -c1-900 gpurun_out/r2ac_c4.log
EF_DGEMM_TC=0 timeout 600 python tools/c4_probe.py 100000 chol::1 2>&1 | cut -c1-300 | tail -4
