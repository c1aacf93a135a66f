set -x
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q > gpurun_out/r2aa_pytest.log 2>&1; tail -30 gpurun_out/r2aa_pytest.log | cut -c1-300
timeout 300 python tools/c3_probe.py 1000000 2>&1 | tail -3
