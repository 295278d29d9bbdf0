#!/bin/bash
# A/B of the GroupNorm kernels on one box: tests, micro-benchmark under the slab-kernel switches, whole-forward profile.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -k "groupnorm or fp16_activations" -m gpu -q -x -s --no-header -p no:cacheprovider > gpurun_out/gn_test.log 2>&1
echo "tests exit $?"; tail -n 22 gpurun_out/gn_test.log | cut -c1-120
: > gpurun_out/gn_bench.log
for cfg in "${@:-CPD_GN_SLAB=1}"; do
  echo "== $cfg" | tee -a gpurun_out/gn_bench.log
  env $cfg timeout 300 python tools/bench_norm.py --only gn 2>&1 | tee -a gpurun_out/gn_bench.log
done
