#!/bin/bash
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
V=base,base,base,base,base,base,base
run() { local name=$1; shift
  env "$@" NANHUNT_VARIANTS=$V POLUS_LOGGER_LEVEL=ERROR timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port $((29600 + RANDOM % 200)) tools/nan_hunt.py 128 20 > $OUT/nanhunt_$name.log 2> $OUT/nanhunt_$name.err
  echo "== $name rc=$? nonfinite trials (rank 0): $(grep -o 'NANHUNT {[^}]*}' $OUT/nanhunt_$name.log | grep '"rank": 0' | grep -c 'first_nonfinite_step": [0-9]') of $(grep -o 'NANHUNT {[^}]*}' $OUT/nanhunt_$name.log | grep -c '"rank": 0')  steps: $(grep -o 'NANHUNT {[^}]*}' $OUT/nanhunt_$name.log | grep '"rank": 0' | grep -o 'first_nonfinite_step": [0-9a-z]*' | cut -d' ' -f2 | tr '\n' ' ')"; tail -2 $OUT/nanhunt_$name.err | cut -c1-200; }
V=base:cfg5@32,base:cfg4@64,base:cfg5@32,base:ner_base@128,base:cfg5@32,base:cfg4@64,base:cfg5@32
run fixed_cfgs2 X=1
run fixed_cfgs3 POLUS_NCCL_MAX_CTAS=8
