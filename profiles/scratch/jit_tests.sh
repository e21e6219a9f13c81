set -x
HB_JIT_MIN_WORK=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_em.py tests/test_gpu_change_factor.py -m gpu -x -q 2>&1 | tail -n 8
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench/configs.py --only C3,C5 2>/dev/null
