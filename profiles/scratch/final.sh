set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err; cat gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -n 2 gpurun_out/ncu_bench.log | cut -c1-300
python profiles/prof_pass.py 1048576 3
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --kernel-name-base demangled -k 'regex:prodsum_kernel|fused_kernel|hb_step_' -s 166 -c 83 --csv --log-file gpurun_out/pass_dram.csv python profiles/prof_pass.py 1048576 3 > gpurun_out/ncu_pass.log 2>&1; tail -n 2 gpurun_out/ncu_pass.log | cut -c1-300
ncu --set full --import-source on --clock-control none -k 'regex:hb_step_(30|31|32)$' -s 6 -c 3 -o gpurun_out/jit_top python profiles/prof_pass.py 262144 3 > gpurun_out/ncu_full.log 2>&1; tail -n 2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
