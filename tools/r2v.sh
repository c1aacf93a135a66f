set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
python bench.py --impl reference > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2v_bench20.json 2> gpurun_out/r2v_bench20.err
python bench.py --profile --steps 20 --warmup 3 > gpurun_out/r2v_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --profile --steps 20 --warmup 3 > gpurun_out/r2v_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recognize_stream -c 2 -o gpurun_out/r2v_stream python bench.py --profile --steps 20 --warmup 3 --no-extras > gpurun_out/r2v_ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:preprocess_kernel -s 2 -c 2 -o gpurun_out/r2v_pre python bench.py --profile --steps 5 --warmup 3 --extras preprocess > gpurun_out/r2v_ncu_pre.log 2>&1
tail -3 gpurun_out/r2v_pytest.log
