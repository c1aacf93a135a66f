# ncu evidence of the round-2 tree (each profiled command exits 0 without ncu first; numbers printed under ncu are not bench values)
set -x
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --profile --steps 20 --warmup 3 > gpurun_out/r2_ev_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_ev_launches.csv python bench.py --profile --steps 20 --warmup 3 > gpurun_out/r2_ev_launch.log 2>&1
python tools/stream_probe.py > gpurun_out/r2_ev_stream_plain.log 2>&1 && \
timeout 600 $NCU -k regex:recognize_stream -s 3 -c 1 -o gpurun_out/r2_ev_stream python tools/stream_probe.py > gpurun_out/r2_ev_stream_ncu.log 2>&1
python tools/preprocess_probe.py > gpurun_out/r2_ev_pre_plain.log 2>&1 && \
timeout 600 $NCU -k regex:preprocess_kernel -s 1 -c 1 -o gpurun_out/r2_ev_pre_gray python tools/preprocess_probe.py > gpurun_out/r2_ev_pre_ncu.log 2>&1
timeout 600 $NCU -k regex:preprocess_kernel -s 45 -c 1 -o gpurun_out/r2_ev_pre_bgr python tools/preprocess_probe.py >> gpurun_out/r2_ev_pre_ncu.log 2>&1
python tools/c3_probe.py 1000000 > gpurun_out/r2_ev_c3_plain.log 2>&1 && \
timeout 600 $NCU -k regex:match_tc_kernel -s 1 -c 1 -o gpurun_out/r2_ev_match_tc python tools/c3_probe.py 1000000 > gpurun_out/r2_ev_c3_ncu.log 2>&1
python tools/gram_probe.py 12500,10000,1 > gpurun_out/r2_ev_gram_plain.log 2>&1 && \
timeout 600 $NCU -k regex:gram_tc_kernel -s 1 -c 1 -o gpurun_out/r2_ev_gram python tools/gram_probe.py 12500,10000,1 > gpurun_out/r2_ev_gram_ncu.log 2>&1
python tools/dgemm_probe.py > gpurun_out/r2_ev_dgemm_plain.log 2>&1 && \
timeout 600 $NCU -k regex:dgemm_tc_kernel -s 5 -c 1 -o gpurun_out/r2_ev_dgemm python tools/dgemm_probe.py > gpurun_out/r2_ev_dgemm_ncu.log 2>&1
ls -la gpurun_out/r2_ev_*
