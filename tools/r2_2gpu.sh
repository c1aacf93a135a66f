set -x
nvidia-smi -L
python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2ai_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2ai_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ai_bench_2gpu.json 2> gpurun_out/r2ai_bench_2gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2ai_bench_2gpu.err
