"""N-rank data-parallel parity check (test infrastructure; also called by bench.py at N > 1 so that the driver's
scaling run records it).  Horovod semantics (polus/training.py:88-96,182-185): N ranks, each on its 1/N slice of a
global batch with lr x N and averaged gradients, must follow the SAME trajectory as one process on the whole batch with
lr x N (the loss is a batch mean), and the replicas must stay bit-identical.

Every rank runs both arms: the data-parallel one through the real trainer (NCCL allreduce inside the captured step) and
the single-rank one with the collective backend swapped for the one-rank mock, in the same process."""
import json
import os

import numpy as np


def _build(seed):
    from polus_b200 import ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.utils import set_random_seed
    tensor.reset_arena()
    set_random_seed(seed)
    ops.set_step(0)
    cfg = BertConfig(vocab_size=300, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                     max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    return BertNERModel(cfg, output_classes=4, droupout_p=0.0)


def run_dp_parity(steps=3, per_rank=4, seq=32, base_lr=1e-3, seed=11, log=None):
    """Returns {"world", "replicas_identical", "vs_single_rank_rel", "lr", "losses_dp_mean", "losses_single"} on every
    rank (identical content).  vs_single_rank_rel = max over steps of |mean_r loss_r - loss_single| / |loss_single|,
    plus the final weights' max / mean abs difference.  (Adam turns the sign of a near-zero gradient into a +-lr step, so
    single weights may differ by up to 2 x lr x steps between two correct runs that only differ in the order of fp32
    additions; the MEAN difference is the sharp figure.)"""
    from polus_b200 import comm, hvd
    from polus_b200.mock import horovod as mock_hvd
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from tests.parity import make_batch
    log = log or (lambda m: None)
    h = hvd()
    world, rank = h.size(), (h.rank() if hasattr(h, "rank") else 0)
    GB = per_rank * world
    rng = np.random.default_rng(5)
    ids, mask, tt, tags = make_batch(rng, GB, seq, 300, 4)
    onehot = np.eye(4, dtype=np.float32)[tags]

    def feed(lo, hi):
        return {"input_ids": ids[lo:hi], "attention_mask": mask[lo:hi], "token_type_ids": tt[lo:hi]}, onehot[lo:hi]

    # ---- arm 1: N ranks, each on its slice (the trainer multiplies lr by N and averages gradients itself)
    model = _build(seed)
    x, y = feed(rank * per_rank, (rank + 1) * per_rank)
    model(**x, training=False)
    opt = Adam(base_lr)
    trainer = ClassifierTrainer(model, opt, model.loss)
    if trainer.use_horovod:
        trainer.broadcast_init_vars()
    log("dp arm built")
    losses = []
    for i in range(steps):
        losses.append(float(trainer.train_step(x, y)))
        log(f"dp step {i} loss {losses[-1]:.5f}")
    w_dp = [w.numpy().copy() for w in model.weights]
    digest = float(sum(np.abs(w.astype(np.float64)).sum() for w in w_dp))
    probe = w_dp[-3].reshape(-1)[:16].astype(np.float64).tolist()
    lr_used = float(opt.learning_rate.read_value())
    mine = json.dumps({"rank": rank, "losses": losses, "digest": digest, "probe": probe}).encode()
    everyone = [json.loads(b.decode()) for b in (comm._host_allgather(mine) if world > 1 else [mine])]
    everyone.sort(key=lambda d: d["rank"])
    identical = all(d["digest"] == everyone[0]["digest"] and d["probe"] == everyone[0]["probe"] for d in everyone)
    losses_dp_mean = np.mean([d["losses"] for d in everyone], axis=0)
    trainer.release_graphs()
    del trainer, model

    # ---- arm 2: ONE rank on the whole global batch with lr x N, no collective
    model = _build(seed)
    xg, yg = feed(0, GB)
    model(**xg, training=False)
    opt1 = Adam(base_lr * world)
    trainer1 = ClassifierTrainer(model, opt1, model.loss)
    if trainer1.use_horovod:   # undo what the constructor did for the N-rank world
        opt1.learning_rate.assign(base_lr * world)
        opt1.grad_scale = 1.0
        trainer1.use_horovod, trainer1.hvd = False, mock_hvd
    losses1 = [float(trainer1.train_step(xg, yg)) for _ in range(steps)]
    w_1 = [w.numpy().copy() for w in model.weights]
    log("single arm done")
    rel = float(np.max(np.abs(losses_dp_mean - np.asarray(losses1)) / np.maximum(np.abs(losses1), 1e-12)))
    wdiff = float(max(np.abs(a - b).max() for a, b in zip(w_dp, w_1)))
    wmean = float(sum(np.abs(a - b).sum() for a, b in zip(w_dp, w_1)) / sum(a.size for a in w_dp))
    trainer1.release_graphs()
    del trainer1, model
    from polus_b200 import tensor
    tensor.reset_arena()
    # hvd.allgather_object with device tensors inside (what ValidationDataCallback gathers, polus/callbacks.py:249): the
    # tensors travel through ncclAllGather, the host labels through the socket store
    pred = tensor.Tensor.from_numpy(np.arange(12, dtype=np.int32).reshape(3, 4) + 100 * rank, tensor.I32)
    got = h.allgather_object((pred, {"labels": np.full(3, rank), "rank": rank}))
    gather_ok = len(got) == world and all(
        np.array_equal(got[r][0].numpy() if hasattr(got[r][0], "numpy") else np.asarray(got[r][0]),
                       np.arange(12, dtype=np.int32).reshape(3, 4) + 100 * r)
        and got[r][1]["rank"] == r and np.array_equal(got[r][1]["labels"], np.full(3, r)) for r in range(world))
    return {"allgather_ok": bool(gather_ok), "world": world, "replicas_identical": bool(identical), "vs_single_rank_rel": rel, "weights_max_abs_diff": wdiff,
            "weights_mean_abs_diff": wmean, "lr": lr_used, "steps": steps, "losses_dp_mean": [float(v) for v in losses_dp_mean],
            "losses_single": [float(v) for v in losses1]}


if __name__ == "__main__":   # python -m torch.distributed.run ... tests/dp_parity.py
    import faulthandler
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    faulthandler.dump_traceback_later(int(os.environ.get("DP_PARITY_WATCHDOG", "150")), exit=True)
    t0 = time.time()
    r = int(os.environ.get("RANK", "0"))

    def _log(m):
        sys.stderr.write(f"[dp_parity r{r} +{time.time() - t0:5.1f}s] {m}\n")
        sys.stderr.flush()
    import polus_b200
    from polus_b200 import comm, device
    device.init(int(os.environ.get("LOCAL_RANK", r)))
    _log("device ready")
    polus_b200.PolusContext()
    _log("context ready")
    out = run_dp_parity(steps=4, log=_log)
    if r == 0:
        print("DPPARITY " + json.dumps(out), flush=True)
    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    os._exit(0)   # like bench.py: leave without tearing the communicator down (graphs holding NCCL kernels are still alive)
