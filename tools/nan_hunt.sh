#!/bin/bash
# tools/nan_hunt.sh N : the NaN hunt under the NCCL settings of the round-2 scale probe (run on the GPU box)
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() { local name=$1; shift
  env "$@" POLUS_LOGGER_LEVEL=ERROR timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port $((29600 + RANDOM % 200)) tools/nan_hunt.py 128 60 > $OUT/nanhunt_$name.log 2> $OUT/nanhunt_$name.err
  echo "== $name rc=$?"; grep NANHUNT $OUT/nanhunt_$name.log | tr '}' '\n' | grep '"rank": 0' | cut -c1-260; tail -2 $OUT/nanhunt_$name.err; }
run cta8_order POLUS_NCCL_MAX_CTAS=8 NANHUNT_VARIANTS=noearly,noearly,base,base,noside,base,noearly NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$OUT/nccl_cta8_%h_%p.log
run cta8_noreg POLUS_NCCL_MAX_CTAS=8 NCCL_GRAPH_REGISTER=0 NANHUNT_VARIANTS=base,base,base,base
run cta8_nopdl POLUS_NCCL_MAX_CTAS=8 POLUS_PDL=0 NANHUNT_VARIANTS=base,base,base,base
grep -il "regist" $OUT/nccl_cta8_* | head; grep -i "regist\|NVLS\|Algo\|proto" $OUT/nccl_cta8_* | cut -c1-200 | sort | uniq -c | sort -rn | head -30
