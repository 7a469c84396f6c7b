#!/bin/bash
# A/B of prebuilt engine variants (limap_b200/lib/variants/*.so, built locally with different -D flags):
# each is copied over liblimap_b200.so, the parity suite of the triangulation path and the device-resident bench run.
# usage: scripts/ab_variants.sh name1 name2 ...
cd "$(dirname "$0")/.."
cp limap_b200/lib/liblimap_b200.so /tmp/orig.so
for v in "$@"; do
  cp limap_b200/lib/variants/$v.so limap_b200/lib/liblimap_b200.so
  echo "== variant $v"
  python -m pytest tests/test_tri_parity_gpu.py -x -q --timeout 600 2>&1 | tail -2
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-lm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']/1e6,1),'M rows/s', d['config'].get('pairs'), d['config']['candidates_per_step_rank0'], d['config']['valid_connections_rank0'])"
done
cp /tmp/orig.so limap_b200/lib/liblimap_b200.so
