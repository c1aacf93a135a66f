EF_STREAM_APPEND_DEBUG=1 timeout 120 python tools/stream_probe.py append 2>&1 | grep -v "ef_stream_probe" | tail -12
timeout 600 python -m pytest tests/test_gpu_recognize.py -x -q 2>&1 | tail -3
