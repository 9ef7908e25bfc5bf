#!/bin/bash
# N-GPU probe of the gradient-exchange settings (run on the GPU box):  tools/r02/scale_probe.sh N [extra bench args]
# One bench line per variant into gpurun_out/r02_scale${N}_<variant>.log; the dp_parity harness runs first.
N=${1:-2}; shift
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  env "$@" BENCH_DEBUG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 --no-extra-configs $EXTRA \
      > $OUT/r02_scale${N}_${name}.log 2> $OUT/r02_scale${N}_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    l=[x for x in open("$OUT/r02_scale${N}_${name}.log") if x.startswith("{")][-1]; d=json.loads(l)
    a=d.get("allreduce",{}); s=d.get("strong_gb256",{})
    print({k:round(v,3) if isinstance(v,float) else v for k,v in dict(value=d["value"], ms=d["ms_per_step"], e2e=d["e2e"]["value"], ar_ms=a.get("ms"), busbw=a.get("busbw_GBps"), exposed=a.get("in_step_exposed_ms"), wire=a.get("wire_dtype"), buckets=a.get("buckets"), strong=s.get("seq_s"), strong_ms=s.get("ms_per_step"), dp=d.get("dp_parity",{}).get("vs_single_rank_rel"), ident=d.get("dp_parity",{}).get("replicas_identical")).items()})
except Exception as e:
    print("no line:", e)
PY
}
EXTRA="$@"
DP_PARITY_WATCHDOG=120 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    tests/dp_parity.py > $OUT/r02_dp_parity_n${N}.log 2> $OUT/r02_dp_parity_n${N}.err
echo "== dp_parity rc=$?"; grep DPPARITY $OUT/r02_dp_parity_n${N}.log; tail -5 $OUT/r02_dp_parity_n${N}.err
run f32 POLUS_GRAD_WIRE=f32
run bf16 POLUS_GRAD_WIRE=bf16
run f32_cta8 POLUS_GRAD_WIRE=f32 POLUS_NCCL_MAX_CTAS=8
run bf16_cta8 POLUS_GRAD_WIRE=bf16 POLUS_NCCL_MAX_CTAS=8
run bf16_cta4_b32 POLUS_GRAD_WIRE=bf16 POLUS_NCCL_MAX_CTAS=4 POLUS_BUCKET_MB=32
