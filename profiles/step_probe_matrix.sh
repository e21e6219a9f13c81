#!/bin/bash
# the knob matrix behind the launch-shape choices of the generated step kernels (profiles/step_probe.py); one JSON line per run
out=gpurun_out/step_probe.jsonl
mkdir -p gpurun_out
: > $out
run() { # config, then VAR=value ...
  local cfg=$1; shift
  env "$@" timeout 300 python profiles/step_probe.py $cfg >> $out 2>> gpurun_out/step_probe.err || echo "{\"config\": \"$cfg\", \"failed\": \"$*\"}" >> $out
}
run C5 HB_JIT_INTERLEAVE=0
run C5 HB_JIT_INTERLEAVE=1
run C5 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=2
run C5 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=4
run C5 HB_JIT_INTERLEAVE=1 HB_JIT_WAVES=8
run C5 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=4 HB_JIT_WAVES=8
run C5 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=2 HB_JIT_WAVES=4 HB_JIT_MAX_KERNELS=400 HB_JIT_BUDGET=100000
run C4 HB_JIT=0
run C4 HB_JIT_INTERLEAVE=1
run C4 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=4
run C4 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=4 HB_JIT_WAVES=8
run C3 HB_JIT_INTERLEAVE=0
run C3 HB_JIT_INTERLEAVE=1
run C3 HB_JIT_INTERLEAVE=1 HB_JIT_UNROLL=4 HB_JIT_WAVES=4
run C3 HB_JIT_INTERLEAVE=2 HB_JIT_UNROLL=2
cat $out
