#!/bin/bash
# resident blocks per SM of the persistent closest-hit kernel: is it latency-bound (scales with warps) or throughput-bound?
for B in 1 2 3 4 5; do
  IZPI_TRACE_BLOCKS_PER_SM=$B python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/SM=$B', round(d['value'],1))"
done
