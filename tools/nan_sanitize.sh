#!/bin/bash
# compute-sanitizer (initcheck, then memcheck) over a 2-rank run of two consecutive models (tools/nan_hunt.py, tiny batch)
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
for tool in initcheck memcheck; do
  POLUS_LOGGER_LEVEL=ERROR NANHUNT_VARIANTS=base,base timeout 420 compute-sanitizer --tool $tool --target-processes all --print-limit 30 \
      --log-file $OUT/sanitize_${tool}_%p.log \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      tools/nan_hunt.py 4 4 > $OUT/sanitize_$tool.out 2> $OUT/sanitize_$tool.err
  echo "== $tool rc=$?"; grep -c NANHUNT $OUT/sanitize_$tool.out
  for f in $OUT/sanitize_${tool}_*.log; do echo "-- $f: $(grep -c '=========' $f) lines"; grep -E "ERROR SUMMARY|Uninitialized|Invalid|at .*\+0x|by .*kernel" $f | sort | uniq -c | sort -rn | head -12; done
done
