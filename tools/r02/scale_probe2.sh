#!/bin/bash
# bucket-size probe of the in-step exposed exchange time:  tools/r02/scale_probe2.sh N
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
for mb in 64 32 128 16; do
  env POLUS_BUCKET_MB=$mb timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra-configs --no-strong --no-dp-parity --preheat-steps 60 \
      > $OUT/r02_bucket_probe_n${N}_mb$mb.log 2> $OUT/r02_bucket_probe_n${N}_mb$mb.err
  python - <<PY
import json
try:
    d=json.loads([x for x in open("$OUT/r02_bucket_probe_n${N}_mb$mb.log") if x.startswith("{")][-1]); a=d["allreduce"]
    print("bucket_mb=$mb", {k: round(v,3) if isinstance(v,float) else v for k,v in dict(value=d["value"], ms=d["ms_per_step"], buckets=a["buckets"], ar_ms=a["ms"], busbw=a["busbw_GBps"], exposed=a["in_step_exposed_ms"], loss=d["config"]["loss_last"]).items()})
except Exception as e:
    print("bucket_mb=$mb no line:", e)
PY
done | tee $OUT/r02_bucket_probe_n$N.log
