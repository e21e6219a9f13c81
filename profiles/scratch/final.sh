set -x
python -m pytest tests/test_gpu_jit.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err; cat gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -n 1 gpurun_out/ncu_bench.log | cut -c1-200
python profiles/prof_pass.py 1048576 3
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --kernel-name-base demangled -k 'regex:prodsum_kernel|fused_kernel|hb_step_' -s 166 -c 83 --csv --log-file gpurun_out/pass_dram.csv python profiles/prof_pass.py 1048576 3 > gpurun_out/ncu_pass.log 2>&1; tail -n 1 gpurun_out/ncu_pass.log | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json | cut -c1-400
