#!/bin/bash
# A/B helper: GEMM micro-benchmark of the epilogue-bound layer shapes (CUDA events, cold operands), 3 repetitions
python tools/bench_gemm.py 2>&1 | tail -30
