#!/bin/bash
# Round-end measurement on ONE B200 (run under gpurun; everything lands in gpurun_out/ and is copied into profiles/
# by hand afterwards): GPU tests, smoke, the bench line, one `ncu --set full` capture per shipped hot kernel and the
# launch list of a short bench run. Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
tag=${1:-r02f}
python -m pytest tests -m gpu -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
tail -c 400 gpurun_out/${tag}_bench_1gpu.err
for k in tri:tri_node_kernel lm:lm_refine_kernel remerge:remerge_pairs_kernel vp:jlinkage_kernel; do
  w=${k%%:*}; kn=${k#*:}
  ncu --set full --clock-control none --import-source on -k regex:$kn -c 1 -f -o gpurun_out/${tag}_$w python scripts/ncu_target.py $w 1 > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
