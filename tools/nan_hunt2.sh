#!/bin/bash
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() { local name=$1; shift
  env "$@" POLUS_LOGGER_LEVEL=ERROR timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port $((29600 + RANDOM % 200)) tools/nan_hunt.py 64 20 > $OUT/nanhunt_$name.log 2> $OUT/nanhunt_$name.err
  echo "== $name rc=$?"; grep NANHUNT $OUT/nanhunt_$name.log | tr '}' '\n' | grep '"rank": 0' | cut -c1-330; tail -2 $OUT/nanhunt_$name.err; }
run cfg4first NANHUNT_VARIANTS=base:cfg4@64,base:cfg5@32,base:ner_base@128,base:cfg4@64,base:cfg5@32 NANHUNT_CHECK_EVERY=1
