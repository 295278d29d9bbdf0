#!/bin/bash
# Interleaved whole-forward A/B of one environment setting on one box.  Usage: bash tools/env_ab.sh "VAR=value [VAR2=value]" [rounds]
mkdir -p gpurun_out
cfg=$1; rounds=${2:-2}
for i in $(seq 1 $rounds); do
  timeout 600 python tools/profile_layers.py --reps 5 > gpurun_out/layers_base_$i.txt 2>&1
  echo "base run $i: $(head -n 1 gpurun_out/layers_base_$i.txt) $(tail -n 1 gpurun_out/layers_base_$i.txt | cut -c1-100)"
  env $cfg timeout 600 python tools/profile_layers.py --reps 5 > gpurun_out/layers_alt_$i.txt 2>&1
  echo "alt  run $i: $(head -n 1 gpurun_out/layers_alt_$i.txt) $(tail -n 1 gpurun_out/layers_alt_$i.txt | cut -c1-100)"
done
