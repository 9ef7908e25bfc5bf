"""Collective backend with the surface polus uses from horovod.tensorflow -- the six functions of the
reference's polus/mock/horovod.py:5-24 (+ rank()):

    init, size, local_rank, DistributedGradientTape, broadcast_variables, allgather_object

One process per GPU (polus/__init__.py:122).  Device collectives are NCCL over NVLink, issued from
libpolus_b200.so on a dedicated high-priority stream; the gradient allreduce runs bucket by bucket
*while backward is still producing earlier layers' gradients* (reference: Horovod's background
fusion thread, polus/training.py:182-185).  The host-side rendezvous (sharing the NCCL unique id,
pickled-object gathers for validation) rides on torch.distributed's gloo store when the process was
started by torchrun, or on a shared directory (POLUS_RENDEZVOUS_DIR) otherwise -- plumbing only.
"""
import ctypes as C
import os
import pickle
import time

import numpy as np

from . import _lib, device
from .tensor import Param

_state = {"initialised": False, "rank": 0, "local_rank": 0, "size": 1, "nccl": False, "gloo": False,
          "comm_stream": None, "seq": 0}

BUCKET_BYTES = int(os.environ.get("POLUS_BUCKET_MB", "64")) * 1024 * 1024


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def init(use_device=True):
    """hvd.init(): returns "mock" when the process is alone (same contract as polus/mock/horovod.py:5-6)."""
    if _state["initialised"]:
        return None if _state["size"] > 1 else "mock"
    size = _env_int("WORLD_SIZE", 1)
    rank = _env_int("RANK", 0)
    local_rank = _env_int("LOCAL_RANK", rank)
    _state.update(rank=rank, local_rank=local_rank, size=size, initialised=True)
    if size <= 1:
        return "mock"
    _host_rendezvous_init()
    if use_device:
        device.init(local_rank)
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            _lib.call("polus_comm_unique_id", uid.ctypes.data)
        uid = np.frombuffer(_host_broadcast(uid.tobytes(), 0), np.uint8).copy()
        _lib.call("polus_comm_init", rank, size, uid.ctypes.data)
        s = C.c_void_p()
        _lib.call("polus_stream_create", C.byref(s), 1)
        _state["comm_stream"] = s.value
        _state["nccl"] = True
    return None


def size():
    return _state["size"]


def rank():
    return _state["rank"]


def local_rank():
    return _state["local_rank"]


def shutdown():
    if _state["nccl"]:
        _lib.call("polus_comm_destroy")
        _state["nccl"] = False


# ------------------------------------------------------------------------------------------------
# host-side rendezvous (pickled objects, a few hundred bytes)
# ------------------------------------------------------------------------------------------------
def _host_rendezvous_init():
    if os.environ.get("POLUS_RENDEZVOUS_DIR"):
        os.makedirs(os.environ["POLUS_RENDEZVOUS_DIR"], exist_ok=True)
        return
    import torch.distributed as dist  # plumbing: store + gloo object collectives
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=_state["rank"], world_size=_state["size"])
    _state["gloo"] = True


def _file_allgather(payload):
    d = os.environ["POLUS_RENDEZVOUS_DIR"]
    seq = _state["seq"]
    _state["seq"] += 1
    mine = os.path.join(d, f"ag_{seq}_{_state['rank']}")
    with open(mine + ".tmp", "wb") as f:
        f.write(payload)
    os.replace(mine + ".tmp", mine)
    out = []
    for r in range(_state["size"]):
        path = os.path.join(d, f"ag_{seq}_{r}")
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > 300:
                raise TimeoutError(f"rendezvous: rank {r} never wrote {path}")
            time.sleep(0.005)
        with open(path, "rb") as f:
            out.append(f.read())
    return out


def _host_allgather(payload: bytes):
    if _state["size"] == 1:
        return [payload]
    if _state["gloo"]:
        import torch.distributed as dist
        out = [None] * _state["size"]
        dist.all_gather_object(out, payload)
        return out
    return _file_allgather(payload)


def _host_broadcast(payload: bytes, root=0):
    return _host_allgather(payload)[root]


def barrier():
    _host_allgather(b"")


def allgather_object(obj):
    """hvd.allgather_object (polus/callbacks.py:249): list with every rank's object, rank order."""
    from .tensor import Tensor

    def to_host(o):
        if isinstance(o, Tensor):
            return o.numpy()
        if isinstance(o, (list, tuple)):
            return type(o)(to_host(x) for x in o)
        return o
    return [pickle.loads(b) for b in _host_allgather(pickle.dumps(to_host(obj)))]


# ------------------------------------------------------------------------------------------------
# device collectives
# ------------------------------------------------------------------------------------------------
def broadcast_variables(variables, root_rank=0):
    """hvd.broadcast_variables (polus/training.py:208-211): rank-0 values to every rank.  Adjacent
    arena variables are merged so BERT-base ships in a handful of ncclBroadcast calls."""
    if _state["size"] == 1 or not _state["nccl"]:
        return
    spans = []
    for v in variables:
        spans.append((v.ptr, v.nbytes))
        if isinstance(v, Param):
            spans.append((v.shadow.ptr, v.shadow.nbytes))
    spans.sort()
    merged = []
    for ptr, n in spans:
        if merged and 0 <= ptr - (merged[-1][0] + merged[-1][1]) <= 256 and (ptr - merged[-1][0]) % 4 == 0:
            merged[-1][1] = ptr + n - merged[-1][0]
        else:
            merged.append([ptr, n])
    st = device.stream()
    for ptr, n in merged:
        _lib.call("polus_comm_broadcast", ptr, n, root_rank, st)


def plan_buckets(weights, bucket_bytes=None):
    """Group the gradient arena spans of `weights` into contiguous buckets of ~bucket_bytes, ordered
    from the END of the arena (the last-created variables get their gradients first in backward).
    Returns [(chunk, offset_elems, n_elems, [params])...]."""
    bucket_bytes = bucket_bytes or BUCKET_BYTES
    ps = sorted((w for w in {id(w): w for w in weights if isinstance(w, Param)}.values()),
                key=lambda w: (id(w.chunk), w.offset), reverse=True)
    buckets = []
    cur = None
    for w in ps:
        n = (w.size + 63) & ~63
        if cur is not None and cur["chunk"] is w.chunk and w.offset + n == cur["off"] and cur["n"] * 4 < bucket_bytes:
            cur["off"] = w.offset
            cur["n"] += n
            cur["params"].append(w)
        else:
            cur = {"chunk": w.chunk, "off": w.offset, "n": n, "params": [w]}
            buckets.append(cur)
    return [(b["chunk"], b["off"], b["n"], b["params"]) for b in buckets]


class _DistributedTape:
    """tape.gradient() + bucketed NCCL allreduce overlapped with the rest of backward."""

    def __init__(self, tape):
        self.tape = tape

    def gradient(self, loss, weights, on_bucket_ready=None):
        from . import ops
        tape = self.tape
        comm_stream, main = _state["comm_stream"], device.stream()
        events = []

        def launch(bucket):
            ch, off, n, _ = bucket
            ev = C.c_void_p()
            _lib.call("polus_event_create", C.byref(ev))
            _lib.call("polus_event_record", ev, main)
            _lib.call("polus_stream_wait_event", comm_stream, ev)
            ops.side_join(comm_stream)  # weight gradients of this bucket issued on the background stream
            _lib.call("polus_comm_allreduce_f32", ch.g.ptr + off * 4, n, comm_stream)
            events.append(ev)
            if on_bucket_ready is not None:
                on_bucket_ready(ch, off, n, after=comm_stream)  # the update waits for the reduced bucket only

        before_node, flush = ops.bucket_schedule(tape.nodes, weights, launch)
        grads = ops.run_backward(tape.nodes, loss, before_node)
        flush()
        tape.nodes = []
        # join: the optimizer (main stream) must see every reduced bucket
        done = C.c_void_p()
        _lib.call("polus_event_create", C.byref(done))
        _lib.call("polus_event_record", done, comm_stream)
        _lib.call("polus_stream_wait_event", main, done)
        self.events = events + [done]
        return [w.grad if isinstance(w, Param) else ops._materialise(grads.get(id(w))) for w in weights]


def DistributedGradientTape(tape, **kwargs):
    """hvd.DistributedGradientTape(tape) (polus/training.py:182).  op=Average: the sum is taken here,
    the 1/size lands in the optimizer's grad_scale."""
    if _state["size"] == 1 or not _state["nccl"]:
        return tape
    return _DistributedTape(tape)
