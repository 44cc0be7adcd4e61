#!/bin/bash
# GPU box: solver-kernel time of bench.py's workload under the dataflow / barrier knobs
run() {
  echo "== $*"
  env "$@" python bench.py --no-cpu-baseline --steps 20 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'median', round(d['step_ms_stats']['median'],4), 'e2e', round(d['e2e']['ms_per_step'],3))"
}
run S2B_DATAFLOW=0
run S2B_DATAFLOW=0 S2B_SOLVE_BLOCKS_PER_SM=1
run S2B_DATAFLOW=1
run S2B_DATAFLOW=1 S2B_FLOW_SLEEP_NS=100
run S2B_DATAFLOW=1 S2B_FLOW_SLEEP_NS=400
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=1
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=1 S2B_FLOW_SLEEP_NS=100
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=1 S2B_SOLVE_THREADS=128
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=1 S2B_SOLVE_THREADS=128 S2B_FLOW_SLEEP_NS=100
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=1 S2B_SOLVE_THREADS=64
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=2 S2B_SOLVE_THREADS=128
run S2B_DATAFLOW=1 S2B_SOLVE_BLOCKS_PER_SM=2 S2B_SOLVE_THREADS=64
