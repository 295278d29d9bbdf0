#!/bin/bash
# Runs the GPU test groups in separate processes (a device trap in one group must not hide the others).
# Logs go to gpurun_out/.  Usage: bash tools/gpu_check.sh [group ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then groups=(step gemm conv attn misc unet e2e); fi
rc=0
for g in "${groups[@]}"; do
  case $g in
    step) sel="tests/test_gpu_sampling.py -k 'fused_loop or denoiser_forward or unsupported or more_samplers or threshold or clip or churn'";;
    gemm) sel="tests/test_gpu_ops.py -k 'gemm or fp16_activations'";;
    conv) sel="tests/test_gpu_ops.py -k 'conv3x3 or conv1x1'";;
    attn) sel="tests/test_gpu_ops.py -k 'attention'";;
    misc) sel="tests/test_gpu_ops.py -k 'groupnorm or layernorm or timestep or conv_in_out'";;
    unet) sel="tests/test_gpu_sampling.py -k 'unet_forward'";;
    e2e)  sel="tests/test_gpu_sampling.py -k 'end_to_end or config or vae or softmax_rows'";;
    *) sel="$g";;
  esac
  echo "=== group $g"
  eval timeout 600 python -m pytest $sel -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/test_$g.log 2>&1
  code=$?
  echo "group $g exit $code"; tail -n 6 gpurun_out/test_$g.log
  if [ $code -ne 0 ]; then rc=1; fi
done
exit $rc
