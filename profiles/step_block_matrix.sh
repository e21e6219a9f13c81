#!/bin/bash
# Knob matrix behind the shapes of the re-reading generated step kernels (jit_step_shape / Gen in csrc/jit.cpp) on C3: register
# blocks over several output variables (HB_JIT_BLOCK_U entries in lock step, HB_JIT_BLOCK_V rows per thread, HB_JIT_BLOCK_MINB
# CTAs per SM, HB_JIT_BLOCK_BODY unrolled chain evaluations per group), digit order (HB_JIT_DIGIT_ORDER), split strip loops
# (HB_JIT_SPLIT_LOOP) and leaving out the `0 +` of every sum (HB_JIT_STEP_ELIDE for hb_step_<i>, HB_JIT_ZERO_ADD=1 keeps it
# everywhere).  One JSON line per run (profiles/step_probe.py) -> profiles/r2_step_block.jsonl (three sessions: the first
# lines were taken before digit order / split loops / elision existed, the second set with the elision ON in hb_step_<i>,
# the third with today's code); then the launch list of the default shape and the C2 on-chip kernel with / without the elision.
out=gpurun_out/step_block.jsonl
mkdir -p gpurun_out
: > $out
run() { # config, then VAR=value ...
  local cfg=$1; shift
  env HB_JIT_CACHE=0 "$@" timeout 200 python profiles/step_probe.py $cfg >> $out 2>> gpurun_out/step_block.err || echo "{\"config\": \"$cfg\", \"failed\": \"$*\"}" >> $out
}
run C3 HB_JIT_BLOCK_U=0 HB_JIT_DIGIT_ORDER=0
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320 HB_JIT_SPLIT_LOOP=0
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320 HB_JIT_DIGIT_ORDER=0
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320 HB_JIT_STEP_ELIDE=1
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320 HB_JIT_BLOCK_V=2
run C3 HB_JIT_BLOCK_U=16 HB_JIT_BLOCK_MINB=1 HB_JIT_BLOCK_BODY=320
run C3 HB_JIT_BLOCK_U=16 HB_JIT_BLOCK_MINB=1 HB_JIT_BLOCK_BODY=600
run C3 HB_JIT_BLOCK_U=8 HB_JIT_BLOCK_MINB=1 HB_JIT_BLOCK_BODY=600
run C3 HB_JIT_BLOCK_U=16 HB_JIT_BLOCK_MINB=2 HB_JIT_BLOCK_BODY=320
cat $out
for z in 0 1; do
  HB_JIT_CACHE=0 HB_JIT_ZERO_ADD=$z timeout 200 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/c2_zero_add$z.json 2>> gpurun_out/step_block.err
  cut -c1-200 gpurun_out/c2_zero_add$z.json
done
HB_JIT_CACHE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv \
    --log-file gpurun_out/c3_launches_block.csv python profiles/step_probe.py C3 > gpurun_out/c3_ncu_block.log 2>&1
