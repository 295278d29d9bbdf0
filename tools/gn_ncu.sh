#!/bin/bash
# ncu --set full of the GroupNorm slab kernel on one shape; the report comes back in gpurun_out/.
mkdir -p gpurun_out
idx=${1:-0}
ncu --set full --clock-control none --import-source on -k regex:gn_slab -s 4 -c 1 -f -o gpurun_out/gn_slab_$idx \
  python tools/bench_norm.py --only gn --gn-index $idx --reps 2 > gpurun_out/gn_ncu_$idx.log 2>&1
tail -n 3 gpurun_out/gn_ncu_$idx.log
python tools/ncu_summary.py gpurun_out/gn_slab_$idx.ncu-rep
