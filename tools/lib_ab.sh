#!/bin/bash
# Interleaved whole-forward A/B of two builds of the library on one box: CPD_B200_LIB selects the .so.
# Usage: bash tools/lib_ab.sh <old.so> [rounds]   (the product build is the other arm)
mkdir -p gpurun_out
old=$1; rounds=${2:-2}
for i in $(seq 1 $rounds); do
  CPD_B200_LIB=$PWD/$old timeout 600 python tools/profile_layers.py --reps 5 > gpurun_out/layers_old_$i.txt 2>&1
  echo "old run $i: $(head -n 1 gpurun_out/layers_old_$i.txt) $(tail -n 1 gpurun_out/layers_old_$i.txt | cut -c1-100)"
  timeout 600 python tools/profile_layers.py --reps 5 > gpurun_out/layers_new_$i.txt 2>&1
  echo "new run $i: $(head -n 1 gpurun_out/layers_new_$i.txt) $(tail -n 1 gpurun_out/layers_new_$i.txt | cut -c1-100)"
done
