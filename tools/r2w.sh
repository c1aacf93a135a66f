set -x
timeout 300 python tools/stream_probe.py debug > gpurun_out/r2w_probe.log 2>&1
python tools/preprocess_probe.py > gpurun_out/r2w_pre_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 1 -c 1 -o gpurun_out/r2w_pre_gray python tools/preprocess_probe.py > gpurun_out/r2w_ncu_pre.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 45 -c 1 -o gpurun_out/r2w_pre_bgr python tools/preprocess_probe.py >> gpurun_out/r2w_ncu_pre.log 2>&1
cat gpurun_out/r2w_pre_plain.log
