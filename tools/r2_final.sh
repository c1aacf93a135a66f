set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_k20.json 2> gpurun_out/r2f_bench_k20.err; echo "bench rc=$?"
python bench.py --no-extras > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"
