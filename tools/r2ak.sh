set -x
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_pipeline.py tests/test_gpu_recognize.py tests/test_gpu_blocks.py tests/test_gpu_dist.py -x -q 2>&1 | tail -12 | cut -c1-200
python tools/latency_probe.py 2>&1 | tail -1 | cut -c1-400
