#!/bin/bash
# ncu --set full with source-level stall sampling of ONE GEMM case:  tools/r02/r02_ncu_src.sh NAME "GEMM_ONLY pattern" [GEMM_EXTRA]
OUT=gpurun_out; mkdir -p $OUT
name=$1; pat=$2; extra=${3:-0}
GEMM_EXTRA=$extra GEMM_ONLY="$pat" python tools/gemm_shapes.py 128 > $OUT/plain_$name.log 2>&1 && \
GEMM_EXTRA=$extra GEMM_ONLY="$pat" timeout 280 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o /tmp/ncu_$name python tools/gemm_shapes.py 128 > $OUT/ncu_$name.log 2>&1
ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $OUT/r02_ncu_${name}_raw.csv 2>/dev/null
ncu -i /tmp/ncu_$name.ncu-rep --page source --csv > $OUT/r02_ncu_${name}_source.csv 2>/dev/null
gzip -f $OUT/r02_ncu_${name}_source.csv
ls -la $OUT/r02_ncu_${name}_*; cat $OUT/plain_$name.log | grep case
