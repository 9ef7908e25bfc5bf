#!/bin/bash
# one GPU call: GEMM epilogue A/B (default build vs -DPOLUS_EPI_LATE_WAIT=0 -DPOLUS_EPI_BIAS_EARLY=0), the GPU test suite,
# the ncu capture behind profiles/r02_gemm_roofline.json and the launch list of one op-by-op step at batch 256
OUT=gpurun_out; mkdir -p $OUT
for lib in "" _epi_old; do
  for c in "fwd_ffn1" "dgrad_ffn2*gelu" "fwd_ffn2"; do
    POLUS_LIB=polus_b200/libpolus_b200$lib.so GEMM_ONLY="$c" timeout 120 python tools/gemm_shapes.py 128 2>&1 | grep case | sed "s/^/lib$lib /"
  done
done | tee $OUT/r02_gemm_epi_ab.log
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > $OUT/r02_tests_gpu.log 2>&1; tail -4 $OUT/r02_tests_gpu.log
GEMM_NCU=1 timeout 120 python tools/gemm_shapes.py 256 > $OUT/plain_gemm_ncu.log 2>&1 && \
  GEMM_NCU=1 timeout 600 ncu --set full --clock-control none -k regex:gemm_tc -c 12 -f -o /tmp/r02_gemm12 python tools/gemm_shapes.py 256 > $OUT/ncu_gemm12.log 2>&1
[ -f /tmp/r02_gemm12.ncu-rep ] && ncu -i /tmp/r02_gemm12.ncu-rep --page raw --csv > $OUT/r02_ncu_gemm12_b256_raw.csv 2>/dev/null
ls -la $OUT/r02_ncu_gemm12_b256_raw.csv
timeout 200 python tools/profile_step.py 256 > $OUT/plain_profile_step.log 2>&1 && \
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02_launches_b256.csv python tools/profile_step.py 256 > $OUT/ncu_profile_step.log 2>&1
tail -2 $OUT/plain_profile_step.log; wc -l $OUT/r02_launches_b256.csv
