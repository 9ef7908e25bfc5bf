#!/bin/bash
# ncu --set full captures of the kernels VERDICT r01 names, each only after the same command exited 0 without ncu.
# The reports embed the whole cubin (25-55 MB each): they are reduced on the box to the raw-metric and per-instruction
# CSV pages and deleted, so that gpurun_out/ stays below its 64 MiB limit.
OUT=gpurun_out; mkdir -p $OUT
prof() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  "$@" > $OUT/plain_$name.log 2>&1 && \
  timeout 280 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r02_ncu_$name "$@" > $OUT/ncu_$name.log 2>&1
  local rc=$?
  if [ -f /tmp/r02_ncu_$name.ncu-rep ]; then
    ncu -i /tmp/r02_ncu_$name.ncu-rep --page raw --csv > $OUT/r02_ncu_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/r02_ncu_$name.ncu-rep --page source --csv > $OUT/r02_ncu_${name}_source.csv 2>/dev/null
    ncu -i /tmp/r02_ncu_$name.ncu-rep --page details > $OUT/r02_ncu_${name}_details.txt 2>/dev/null
    gzip -f $OUT/r02_ncu_${name}_source.csv
  fi
  echo "ncu $name rc=$rc $(ls -la $OUT/r02_ncu_${name}_* 2>/dev/null | awk '{print $5}' | tr '\n' ' ')"
}
if [ "$1" != "noattn" ]; then
prof attn_fwd_p01 attn_fwd_kernel 2 python tools/kernel_times.py 128
prof attn_fwd_p0 attn_fwd_kernel 38 python tools/kernel_times.py 128
prof attn_bwd attn_bwd_kernel 2 python tools/kernel_times.py 128
fi
prof ln_bwd ln_res_bwd 2 python tools/kernel_times.py 128
prof gemm_ffn1 gemm_tc 3 env GEMM_ONLY=fwd_ffn1 python tools/gemm_shapes.py 128
prof gemm_dgrad_ffn2 gemm_tc 3 env "GEMM_ONLY=dgrad_ffn2*gelu" python tools/gemm_shapes.py 128
prof gemm_plain_ffn2 gemm_tc 3 env "GEMM_ONLY=fwd_ffn2" python tools/gemm_shapes.py 128
(timeout 300 python -m pytest tests/test_fullsize_gpu.py tests/test_trainer_contract_gpu.py -q -m gpu 2>&1 | tail -30) > $OUT/r02_tests_d.log 2>&1; tail -3 $OUT/r02_tests_d.log
du -sh $OUT
