#!/usr/bin/env python
"""Headline benchmark: training sequences/s of BERT-base NER (+CRF) at seq 256 on N B200s (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU port of the reference step (oracle/torch_ref.py)

One "step" = one full optimisation step (forward, CRF loss, backward, [allreduce], Adam) of
polus_b200.training.ClassifierTrainer.train_step on one synthetic batch per GPU.
`value`  : inputs already resident in HBM, K captured-graph replays, CUDA events, max over ranks.
`e2e`    : the same K steps fed from HOST numpy batches through the public trainer API: pinned-memory
           H2D copy of every input and the D2H read of the loss are inside the timed region.
`roofline`: tcgen05 GEMM launches (the dominant kernel family, ~all FLOPs) bracketed one by one with CUDA
           events in an extra op-by-op step; achieved = sum(2mnk) / sum(durations) vs the measured bf16 peak.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ = 256
TRAIN_GFLOP_PER_SEQ = 137.71  # BASELINE.md §3 (GEMM FLOPs, fwd+bwd = 3x fwd), BERT-base S=256


def workload_name(batch):
    return (f"BERT-base (L12 H768 nh12 I3072) polus.ner token classification + CRF, seq {SEQ}, "
            f"batch {batch}/GPU, dropout 0.1, Adam+warmup")


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batches(n, batch, seq, vocab, K, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ids = rng.integers(1000, vocab, size=(batch, seq)).astype(np.int32)
        ids[:, 0], ids[:, -1] = 101, 102
        mask = np.ones((batch, seq), np.int32)  # throughput config: full-length sequences
        tt = np.zeros((batch, seq), np.int32)
        tags = rng.integers(1, K, size=(batch, seq))
        y = np.eye(K, dtype=np.float32)[tags]
        out.append(({"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}, y))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_ref
    batch = args.ref_batch
    sps, dt, threads = torch_ref.time_train_steps(batch=batch, seq=SEQ, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "train_sequences_per_second", "value": sps, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "global_batch": args.batch * args.gpus, "seq_len": SEQ,
                       "parallelism": f"dp{args.gpus}", "per_step_sample": f"{batch} sequences of the same workload per timed step",
                       "note": "CPU port of the reference step (TensorFlow is not installable in this image; oracle/torch_ref.py)"},
            "cpu_baseline": {"value": sps, "unit": "sequences/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps of {batch} sequences x {SEQ} tokens"},
            "e2e": {"value": sps, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


_T0 = time.time()


def dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - _T0:6.1f}s] {msg}\n")
        sys.stderr.flush()


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    import ctypes as C
    import polus_b200
    from polus_b200 import _lib, comm, device, ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.optimizers import Adam
    from polus_b200.schedulers import warmup_scheduler
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed

    device.init(local_rank)
    dbg("device ready")
    polus_b200.PolusContext()  # brings NCCL up when WORLD_SIZE > 1
    dbg("context ready")
    set_random_seed(42)
    K = 4
    cfg = BertConfig()  # BERT-base: L12 H768 nh12 I3072 vocab 30522, dropout 0.1
    model = BertNERModel(cfg, output_classes=K, hidden_space=128, droupout_p=0.1)
    opt = Adam(warmup_scheduler(10000, 5e-5))
    trainer = ClassifierTrainer(model, opt, model.loss)
    batches = synthetic_batches(4, args.batch, SEQ, cfg.vocab_size, K, seed=1 + rank)

    def barrier():
        device.device_sync()
        if world > 1:
            comm.barrier()

    dbg("model built")
    for i in range(max(args.warmup, 3)):
        loss = trainer.train_step(*batches[i % len(batches)])
        if i < 3:
            device.device_sync()
            dbg(f"warmup step {i} done")
    float(loss)
    dbg("warmup done")
    if trainer.use_horovod:
        trainer.broadcast_init_vars()
    dev_batches = []
    for x, y in batches:
        dev_batches.append(({k: tensor.Tensor.from_numpy(v, tensor.I32) for k, v in x.items()}, tensor.Tensor.from_numpy(y, tensor.F32)))

    def ev():
        e = C.c_void_p()
        _lib.call("polus_event_create", C.byref(e))
        return e

    def timed(feed):
        e0, e1 = ev(), ev()
        barrier()
        l0 = _lib.call("polus_launch_count")
        _lib.call("polus_event_record", e0, device.stream())
        last = None
        for i in range(args.steps):
            last = trainer.train_step(*feed[i % len(feed)])
        _lib.call("polus_event_record", e1, device.stream())
        float(last)
        barrier()
        ms = C.c_float()
        _lib.call("polus_event_elapsed_ms", e0, e1, C.byref(ms))
        return ms.value, _lib.call("polus_launch_count") - l0, float(last)

    # bring the GPU to its steady operating point (power-capped clocks) before either timed region: the first seconds
    # after start-up run at boost clocks, which made whichever region came first look 3-5 % faster than the other
    # (a FIXED number of steps: every rank must issue the same sequence of collectives)
    for i in range(args.preheat_steps):
        last = trainer.train_step(*dev_batches[i % len(dev_batches)])
        if i % 10 == 9:
            float(last)
    dbg("pre-heat done")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_dev, launches, loss_dev = timed(dev_batches)
    dbg("timed (resident) done")
    ms_e2e, _, loss_e2e = timed(batches)
    h2d = trainer.last_h2d_bytes  # bytes the public API copied host -> device for ONE step of the e2e region
    dbg("timed (e2e) done")
    clocks = sampler.stop() if sampler else None
    if world > 1:
        all_ms = [json.loads(b.decode()) for b in comm._host_allgather(json.dumps([ms_dev, ms_e2e]).encode())]
        ms_dev, ms_e2e = max(a[0] for a in all_ms), max(a[1] for a in all_ms)

    # ---- roofline of the dominant kernel family (tcgen05 GEMMs, ~all FLOPs of the step)
    # (1) one op-by-op step counts the launches and their algorithmic FLOPs (2mnk each);
    # (2) a second captured graph of the SAME step in which every kernel entry point except polus_gemm_tc is a
    #     no-op (polus_b200._lib.set_only) replays exactly the step's GEMM launches -- same order, shapes, buffers
    #     and streams -- and is timed with CUDA events like the step itself: duration of the GEMM launches of one
    #     step, free of host launch latency (bracketing single launches op-by-op charges ~5 us of host time each);
    # (3) the marginal figure (step minus the step with polus_gemm_tc ablated) is reported next to it.
    roof = None
    prof = []
    ops.GEMM_PROFILE = prof
    trainer.use_graph = False
    float(trainer.train_step(*batches[0]))  # every rank runs it (it contains the gradient allreduce)
    ops.GEMM_PROFILE = None
    trainer.use_graph = True
    tot_flop = sum(flop for kind, flop, _, _ in prof if kind == "tc")
    n_tc = sum(1 for kind, _, _, _ in prof if kind == "tc")
    barrier()

    def replay_ms(only=None, ablate=None):
        saved = (trainer._compiled, trainer._warm, _lib._ONLY, _lib._ABLATE)
        trainer._compiled, trainer._warm = {}, set()
        if only is not None:
            _lib.set_only(only)
        if ablate is not None:
            _lib._ABLATE = frozenset(ablate)
        try:
            for i in range(3):
                float(trainer.train_step(*dev_batches[i % len(dev_batches)]))
            ms, _, _ = timed(dev_batches)
        finally:
            trainer._compiled, trainer._warm = saved[0], saved[1]
            _lib._ONLY, _lib._ABLATE = saved[2], saved[3]
        return ms / args.steps

    gemm_only_ms = replay_ms(only={"polus_gemm_tc"})
    no_gemm_ms = replay_ms(ablate={"polus_gemm_tc"})
    dbg("gemm-only replays done")
    if world > 1:
        all_g = [json.loads(b.decode()) for b in comm._host_allgather(json.dumps([gemm_only_ms, no_gemm_ms]).encode())]
        gemm_only_ms, no_gemm_ms = max(a[0] for a in all_g), max(a[1] for a in all_g)
    if rank == 0:
        peaks, how = read_peaks()
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        achieved = tot_flop / (gemm_only_ms * 1e-3) / 1e12 if gemm_only_ms > 0 else 0.0
        step_ms = ms_dev / args.steps
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": 222.7e6 * args.batch / 128.0,  # dram__bytes_read+write per launch, mean of 12 launches, batch 128
                "traffic_source": "profiles/r01_ncu_summary_v16_gemm_{fwd,bwd}.txt (ncu --set full, batch 128); scaled linearly with batch",
                "tensor_pipe_active_pct": 65.5,  # sm__pipe_tensor_cycles_active, time-weighted over the same 12 launches
                "algorithmic_bytes_per_launch": 244.2e6 * args.batch / 128.0,
                "kernel": "gemm_tc_kernel (all tcgen05 GEMM launches of one step)", "launches": n_tc,
                "gemm_ms_per_step": gemm_only_ms, "gemm_gflop_per_step": tot_flop / 1e9,
                "method": "CUDA-event time of a captured replay of the step's GEMM launches alone (same order/buffers/streams)",
                "gemm_marginal_ms_per_step": step_ms - no_gemm_ms,
                "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step)",
                "step_mfu": (args.batch * TRAIN_GFLOP_PER_SEQ * 1e9 / (step_ms * 1e-3)) / 1e12 / peak}
    # ---- the exchange step alone: bucketed NCCL allreduce of the fp32 gradient arena, back to back (in the step it is
    # overlapped with backward); bus bandwidth = bytes * 2(N-1)/N / time  (NCCL's definition)
    allreduce = None
    if world > 1:
        buckets = comm.plan_buckets(trainer.trainable_weights)
        nbytes = sum(n for _, _, n, _ in buckets) * 4
        st = device.stream()

        def ar_all():
            for ch, off, n, _ in buckets:
                _lib.call("polus_comm_allreduce_f32", ch.g.ptr + off * 4, n, st)
        for _ in range(3):
            ar_all()
        e0, e1 = ev(), ev()
        barrier()
        _lib.call("polus_event_record", e0, st)
        for _ in range(10):
            ar_all()
        _lib.call("polus_event_record", e1, st)
        barrier()
        ms = C.c_float()
        _lib.call("polus_event_elapsed_ms", e0, e1, C.byref(ms))
        ar_ms = max(json.loads(b.decode()) for b in comm._host_allgather(json.dumps(ms.value / 10).encode()))
        # leave the gradient arena zeroed, as the optimizer kernel does
        for ch, off, n, _ in buckets:
            _lib.call("polus_memset", ch.g.ptr + off * 4, 0, n * 4, st)
        device.device_sync()
        allreduce = {"bytes": int(nbytes), "buckets": len(buckets), "ms": ar_ms,
                     "busbw_GBps": nbytes * 2 * (world - 1) / world / (ar_ms * 1e-3) / 1e9,
                     "nvlink_peak_GBps": 900.0, "nvlink_measured_allreduce_GBps": 725.0,
                     "note": "standalone, back to back; inside the step it runs on a side stream under backward"}
    if rank != 0:
        return
    seqs = args.batch * world * args.steps
    value = seqs / (ms_dev * 1e-3)
    e2e_value = seqs / (ms_e2e * 1e-3)
    line = {"metric": "train_sequences_per_second", "value": value, "unit": "sequences/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.batch),
                       "global_batch": args.batch * world, "seq_len": SEQ, "parallelism": f"dp{world}",
                       "preheat_steps": args.preheat_steps,
                       "l2": "per-step working set (weights 0.2 GB bf16 + activations > 3 GB) exceeds the 126 MB L2; no flush needed",
                       "loss_last": loss_dev},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sequences/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roof}
    if allreduce is not None:
        line["allreduce"] = allreduce
    if world == 1 and not args.no_cpu_baseline:
        from oracle import torch_ref
        sps, dt, threads = torch_ref.time_train_steps(batch=args.ref_batch, seq=SEQ, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": sps, "unit": "sequences/s", "cores": threads, "kind": "port",
                                "sample": f"2 timed steps of {args.ref_batch} sequences x {SEQ} tokens (torch-CPU port, fp32)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="sequences per GPU per step (weak scaling: global batch = batch x N)")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--preheat-steps", type=int, default=100, help="untimed steps before the timed regions (~2.4 s at batch 128)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
