#!/bin/bash
# whole-forward A/B of the GroupNorm kernels on one box, interleaved (the box warms up / power-caps over a run)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k "groupnorm or fp16_activations" -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/gn_test.log 2>&1
echo "tests exit $?"; tail -n 4 gpurun_out/gn_test.log | cut -c1-160
for i in 1 2; do
  for s in 0 1; do
    CPD_GN_SLAB=$s timeout 600 python tools/profile_layers.py --reps 5 > gpurun_out/layers_slab${s}_$i.txt 2>&1
    echo "CPD_GN_SLAB=$s run $i: $(head -n 1 gpurun_out/layers_slab${s}_$i.txt) $(tail -n 1 gpurun_out/layers_slab${s}_$i.txt | cut -c1-90)"
  done
done
