set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ad_c4_launches.csv python tools/c4_probe.py 100000 chol::1 > gpurun_out/r2ad_c4.log 2>&1
tail -3 gpurun_out/r2ad_c4.log | cut -c1-300
