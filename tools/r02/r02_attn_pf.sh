#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for pf in 0 1 2 0 1; do
  POLUS_ATTN_PREFETCH=$pf timeout 200 python tools/kernel_times.py ${1:-128} 2>&1 | grep "attention" | sed "s/^/prefetch=$pf /"
done | tee $OUT/r02_attn_bwd_prefetch_b${1:-128}.log
(timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -3)
