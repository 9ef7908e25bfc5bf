#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for h in 0 1 3 4 5 7; do
  for c in "fwd_ffn1" "dgrad_ffn2*gelu"; do
    POLUS_GEMM_L2HINT=$h GEMM_ONLY="$c" timeout 120 python tools/gemm_shapes.py ${1:-128} 2>&1 | grep case | sed "s/^/hint=$h /"
  done
done | tee $OUT/r02_gemm_l2hint_b${1:-128}.log
