#!/bin/bash
# Round profiles: (1) ncu launch list of one bench step (gpu__time_duration only), (2) DRAM bytes of the dominant kernel's launches
# INSIDE that step (-> profiles/<tag>_traffic.json, what bench.py prints as roofline.traffic), (3) ncu --set full of the hot kernels
# on their hot shapes through the micro-benchmark tools, summarised to text on the box, (4) the SASS mnemonic counts that prove
# tcgen05 / TMA.  Usage (under gpurun): bash tools/capture_profiles.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r02}
mkdir -p gpurun_out /tmp/prof
BENCH="python bench.py --sampler-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-decode --no-bf16-leg"
CPD_BENCH_NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${tag}_launches.csv $BENCH > /tmp/prof/launch.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
CPD_BENCH_NCU=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:gemm2_kernel -c 160 --csv --log-file gpurun_out/${tag}_gemm2_dram.csv $BENCH > /tmp/prof/dram.log 2>&1
python tools/traffic_from_ncu.py gpurun_out/${tag}_gemm2_dram.csv gpurun_out/${tag}_traffic.json >> gpurun_out/${tag}_launches_summary.txt 2>&1
out=gpurun_out/${tag}_ncu_full.txt
: > $out
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/prof/$name "$@" > /tmp/prof/$name.log 2>&1
  echo "## $name: $*" >> $out
  python tools/ncu_summary.py /tmp/prof/$name.ncu-rep >> $out 2>&1
}
cap gemm_conv3_64_320 gemm2_kernel 3 1 python tools/bench_gemm.py --only 0 --reps 2
cap gemm_lin_65536_320_320 gemm2_kernel 3 1 python tools/bench_gemm.py --only 4 --reps 2
cap gemm_lin_16384_640_640 gemm2_kernel 3 1 python tools/bench_gemm.py --only 5 --reps 2
cap gemm_geglu_65536_320_2560 gemm2_kernel 3 1 python tools/bench_gemm.py --only 9 --reps 2
cap gemm_qkv_folded_ln_transposed_tail_65536_320_1152 gemm2_kernel 4 1 python tools/bench_fold.py --profile qkv
cap gemm_residual_ln_partial_sums_65536_384_320 gemm2_kernel 4 1 python tools/bench_fold.py --profile producer
cap attention_self_4096_d40 attention4_kernel 3 1 python tools/bench_attn.py --only 0 --reps 2
cap attention_cross_4096x77_d40 attention5_kernel 3 1 python tools/bench_attn.py --only 3 --reps 2
cap groupnorm "gn_" 6 4 python tools/bench_norm.py --only gn --reps 2
cap layernorm layernorm_sub 3 1 python tools/bench_norm.py --only ln --reps 2
cap sampler_step sampler_step_kernel 4 3 python tools/bench_step.py
cap conv_out conv_out_tiled 3 1 python tools/bench_small.py
cp /tmp/prof/gemm_lin_65536_320_320.ncu-rep gpurun_out/${tag}_gemm_lin320.ncu-rep 2>/dev/null
python tools/sass_summary.py > gpurun_out/${tag}_sass_summary.txt 2>&1
ls -la gpurun_out/${tag}_* | cat
