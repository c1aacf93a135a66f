set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ag_pytest.log; tail -3 gpurun_out/r2ag_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"
