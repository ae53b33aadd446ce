#!/bin/bash
# rebuilds the library with different register caps for the 4-lanes-per-ray kernel and benches each
for MB in 5 6 7 8; do
  sed -i "s/^#define IZPI_G4_MIN_BLOCKS .*/#define IZPI_G4_MIN_BLOCKS $MB/" izpi_b200/csrc/device/trace.cu
  python izpi_b200/build.py --force -v 2>&1 | grep -A2 "trace_g4_kernelILb0" | grep -E "registers|spill" | tr '\n' ' '
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MIN_BLOCKS=$MB', round(d['value'],1), round(d['roofline']['frac'],3))"
done
