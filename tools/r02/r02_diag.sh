#!/bin/bash
# 1-GPU diagnostics behind DESIGN.md's reading of the two epilogue-heavy GEMMs: pipe throughputs and what each part costs
OUT=gpurun_out; mkdir -p $OUT
tools/ubench/alu 2>&1 | tee $OUT/r02_ubench_alu.log
for b in 128 32; do
  echo "== batch $b"
  GEMM_EXTRA=1 GEMM_ONLY="ffn1" timeout 200 python tools/gemm_shapes.py $b 2>&1 | grep case
  GEMM_ONLY="ffn2" timeout 200 python tools/gemm_shapes.py $b 2>&1 | grep case
done | tee $OUT/r02_gemm_epilogue_parts.log
