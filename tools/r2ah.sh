set -x
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
for kb in 48 64 72 96; do
echo "stage KB $kb"
EF_PRE_STAGE_KB=$kb python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras preprocess 2>/dev/null | python tools/print_pre.py
done
echo "byte taps (debug 4)"; EF_PRE_DEBUG=4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --extras preprocess 2>/dev/null | python tools/print_pre.py
