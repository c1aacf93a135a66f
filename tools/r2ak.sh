timeout 900 python -m pytest tests/test_gpu_recognize.py tests/test_gpu_blocks.py tests/test_gpu_dist.py -x -q 2>&1 | tail -3
python tools/latency_probe.py > gpurun_out/r2ak_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2ak_lat_launches.csv python tools/latency_probe.py > gpurun_out/r2ak.log 2>&1
tail -1 gpurun_out/r2ak_plain.log | cut -c1-300
