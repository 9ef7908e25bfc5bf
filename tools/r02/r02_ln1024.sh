#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for pf in 0 1 0 1; do
  POLUS_LN_PREFETCH=$pf KT_H=1024 KT_S=512 KT_ONLY=ln timeout 100 python tools/kernel_times.py 32 2>&1 | grep "ln_res_bwd" | sed "s/^/prefetch=$pf /"
done | tee $OUT/r02_ln_bwd_h1024.log
(timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "ln_residual or large" 2>&1 | tail -3)
