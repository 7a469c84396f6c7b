#!/bin/bash
# A/B: rebuild the engine with different -D flags on the GPU box and run the device-resident bench.
# usage: scripts/ab_build_bench.sh "<flags A>" "<flags B>" ...
set -e
cd "$(dirname "$0")/.."
for flags in "$@"; do
  nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
    -o limap_b200/lib/liblimap_b200.so limap_b200/csrc/*.cu 2>/dev/null
  echo "== flags: $flags"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']/1e6,1),'M rows/s', d['config'].get('pairs'))"
done
