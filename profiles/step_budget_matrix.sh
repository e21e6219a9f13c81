#!/bin/bash
# C3: how many steps get a generated kernel (HB_JIT_BUDGET = unrolled chain evaluations in all) and how many graph lanes
out=gpurun_out/step_budget.jsonl
mkdir -p gpurun_out
: > $out
run() { # config, then VAR=value ...
  local cfg=$1; shift
  env HB_JIT_CACHE=0 "$@" timeout 240 python profiles/step_probe.py $cfg >> $out 2>> gpurun_out/step_budget.err || echo "{\"config\": \"$cfg\", \"failed\": \"$*\"}" >> $out
}
run C3 HB_LANES=32
run C3 HB_LANES=8
run C3 HB_JIT_BUDGET=20000
run C3 HB_JIT_BUDGET=100000
cat $out
