#!/usr/bin/env bash
mkdir -p gpurun_out
for m in 0 1; do for c in 2 0; do
  STAPLE_STREAMED_TRACE=1 timeout 300 python bench.py --no-solver --no-cpu-baseline --steps 12 --warmup 3 --stream-chunk $c --stream-mode $m 2> gpurun_out/trace_m${m}_c${c}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('TRACE mode $m chunk $c e2e ms', round(e['ms_per_step'],4), 'plain ms', round(e['unpipelined_ms_per_step'],4))"
  grep -n "streamed trace" gpurun_out/trace_m${m}_c${c}.err | tail -1
  tail -40 gpurun_out/trace_m${m}_c${c}.err | grep -A 40 "streamed trace" | head -40
done; done
