"""N>1 host path on CPU: two processes (world_size 2) over the package's own socket store (no torch).  Covers rendezvous, hvd-shaped
allgather_object, dataset sharding across ranks, learning-rate scaling and rank-0-only callbacks.
Device collectives (NCCL) are exercised on the GPU box by bench.py --gpus N."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    import polus_b200
    from polus_b200 import PolusContext, comm, hvd
    from polus_b200.data import DataLoader
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.callbacks import runs_if_root
    assert PolusContext().is_horovod_enabled()
    h = hvd()
    assert h is comm and h.size() == 2 and h.rank() == int(os.environ["RANK"])
    gathered = h.allgather_object({"rank": h.rank(), "payload": list(range(h.rank() + 1))})
    def gen():
        for i in range(11):
            yield {"i": i}
    mine = [int(s["i"]) for s in DataLoader(gen).to_tfDataset()]
    class M:
        name = "m"; trainable_weights = []
    opt = Adam(1e-3)
    tr = ClassifierTrainer(M(), opt, None)
    class C:
        @runs_if_root
        def f(self): return "ran"
    comm.barrier()
    assert "torch" not in sys.modules, "the product path must not import torch (north_star)"
    comm.shutdown()
    print(json.dumps({"rank": h.rank(), "gathered": gathered, "mine": mine, "lr": opt.learning_rate.read_value(),
                      "grad_scale": opt.grad_scale, "root_only": C().f()}))
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("store", ["unix", "tcp"])
def test_two_rank_host_path_over_socket_store(store):
    import json
    port, store_port = _free_port(), _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), POLUS_COMM_HOST_ONLY="1", POLUS_LOGGER_LEVEL="ERROR", POLUS_STORE=store,
                   POLUS_STORE_PORT=str(store_port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    for d in outs:
        assert d["gathered"] == [{"rank": 0, "payload": [0]}, {"rank": 1, "payload": [0, 1]}]
        assert abs(d["lr"] - 2e-3) < 1e-12 and d["grad_scale"] == 0.5     # LR x size (training.py:90-94), op=Average
    assert outs[0]["mine"] == [0, 2, 4, 6, 8, 10] and outs[1]["mine"] == [1, 3, 5, 7, 9]  # i mod N (data.py:94-96)
    assert outs[0]["root_only"] == "ran" and outs[1]["root_only"] is None


def test_file_rendezvous_fallback(tmp_path):
    """Same gather over a shared directory (POLUS_RENDEZVOUS_DIR), for launchers that export no MASTER_ADDR."""
    code = textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        from polus_b200 import comm
        assert comm.init(use_device=False) is None
        print(json.dumps(comm.allgather_object(("r", comm.rank()))))
    """) % ROOT
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", POLUS_RENDEZVOUS_DIR=str(tmp_path))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        o, e = p.communicate(timeout=60)
        assert p.returncode == 0, e[-2000:]
        assert o.strip().splitlines()[-1] == '[["r", 0], ["r", 1]]'


def test_broadcast_and_bucket_order_do_not_depend_on_device_addresses():
    """Collectives must be issued in the same order, with the same sizes, on every rank -- but device addresses (and
    host object ids) differ from rank to rank.  Two simulated ranks whose master / shadow / optimizer-slot buffers sit in
    opposite address order must plan identical ncclBroadcast sequences and identical gradient buckets (round 2: spans
    were sorted by local address and buckets by id(chunk); rebuilt models on >= 2 GPUs went non-finite)."""
    from polus_b200 import comm
    from polus_b200.tensor import Param, Tensor, F32, BF16

    class Chunk:
        def __init__(self, index, base):
            self.index, self.base = index, base

    def fake_rank(p_bases, pb_bases, g_base, slot_bases, reverse_ids):
        chunks = [Chunk(1, p_bases[0]), Chunk(2, p_bases[1])]
        if reverse_ids:
            chunks = chunks[::-1]          # (only so that id() order differs; .index stays the creation order)
            chunks.sort(key=lambda c: c.index)
        params = []
        sizes = [(1000, 0), (64, 0), (5000, 0), (300, 1), (70000, 1)]     # (elements, chunk)
        used = [0, 0]
        for n, c in sizes:
            w = Param.__new__(Param)
            Tensor.__init__(w, (n,), F32, ptr=p_bases[c] + used[c] * 4, block=chunks[c])
            w.chunk, w.offset = chunks[c], used[c]
            w.shadow = Tensor((n,), BF16, ptr=pb_bases[c] + used[c] * 2, block=chunks[c])
            w.grad = Tensor((n,), F32, ptr=g_base[c] + used[c] * 4, block=chunks[c])
            used[c] += (n + 63) & ~63
            params.append(w)
        slots = [Tensor((used[0],), F32, ptr=slot_bases[0], block=object()), Tensor((used[0],), F32, ptr=slot_bases[1], block=object())]
        return params, slots

    GB = 1 << 30
    a_params, a_slots = fake_rank([10 * GB, 20 * GB], [30 * GB, 40 * GB], [50 * GB, 60 * GB], [70 * GB, 80 * GB], False)
    b_params, b_slots = fake_rank([20 * GB, 10 * GB], [5 * GB, 2 * GB], [60 * GB, 50 * GB], [80 * GB, 70 * GB], True)
    sizes = lambda spans: [n for _, n in spans]
    assert sizes(comm.plan_broadcast_spans(a_params)) == sizes(comm.plan_broadcast_spans(b_params))
    assert sizes(comm.plan_broadcast_spans(a_slots)) == sizes(comm.plan_broadcast_spans(b_slots))
    assert len(comm.plan_broadcast_spans(a_params)) == 4            # master + shadow of two chunks, each merged into one span
    bk = lambda ps: [(ch.index, off, n) for ch, off, n, _ in comm.plan_buckets(ps, bucket_bytes=64 * 1024)]
    assert bk(a_params) == bk(b_params)
    assert bk(a_params)[0][0] == 2                                  # the last-created chunk's variables are exchanged first
