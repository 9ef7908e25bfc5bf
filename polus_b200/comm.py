"""Collective backend with the surface polus uses from horovod.tensorflow -- the six functions of the
reference's polus/mock/horovod.py:5-24 (+ rank()):

    init, size, local_rank, DistributedGradientTape, broadcast_variables, allgather_object

One process per GPU (polus/__init__.py:122).  Device collectives are NCCL over NVLink, issued from
libpolus_b200.so on a dedicated high-priority stream; the gradient allreduce runs bucket by bucket
*while backward is still producing earlier layers' gradients* (reference: Horovod's background
fusion thread, polus/training.py:182-185).  The host-side rendezvous (sharing the NCCL unique id,
pickled-object gathers) is a small socket store from the standard library keyed on MASTER_ADDR / MASTER_PORT
(whatever launcher exported RANK / WORLD_SIZE / MASTER_*; torchrun does), or a shared directory
(POLUS_RENDEZVOUS_DIR).  No torch anywhere on this path.
"""
import ctypes as C
import os
import pickle
import socket
import struct
import time

import numpy as np

from . import _lib, device
from .tensor import Param

_state = {"initialised": False, "rank": 0, "local_rank": 0, "size": 1, "nccl": False, "store": None,
          "comm_stream": None, "seq": 0}

BUCKET_BYTES = int(os.environ.get("POLUS_BUCKET_MB", "64")) * 1024 * 1024
# wire format of the gradient exchange: "f32" = Horovod's default (no compression, polus/training.py:182 passes none);
# "bf16" = half the NVLink bytes, one bf16 rounding per partial sum (like hvd.Compression.fp16); POLUS_GRAD_WIRE selects
WIRE_DTYPE = os.environ.get("POLUS_GRAD_WIRE", "f32")
assert WIRE_DTYPE in ("f32", "bf16"), "POLUS_GRAD_WIRE must be f32 or bf16"


def wire_bytes_per_element():
    return 2 if WIRE_DTYPE == "bf16" else 4


def allreduce_bucket(ch, off, n, stream):
    """Sum-allreduce n gradient elements of arena chunk `ch` starting at element `off`, in place."""
    if WIRE_DTYPE == "bf16":
        buf = getattr(ch, "wire_scratch", None)   # bf16 staging of the chunk's gradients; lives and dies with the chunk
        if buf is None:
            buf = ch.wire_scratch = device.Buffer(ch.capacity * 2)
        _lib.call("polus_comm_allreduce_bf16", ch.g.ptr + off * 4, buf.ptr + off * 2, n, stream)
    else:
        _lib.call("polus_comm_allreduce_f32", ch.g.ptr + off * 4, n, stream)


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
        _lib.call("polus_comm_init_cfg", rank, size, uid.ctypes.data, _env_int("POLUS_NCCL_MAX_CTAS", 0))
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
    st = _state.get("store")
    if st is not None:
        for c in (st if isinstance(st, list) else [st]):
            try:
                c.close()
            except OSError:
                pass
        _state["store"] = None


# ------------------------------------------------------------------------------------------------
# host-side rendezvous (pickled objects, a few hundred bytes) -- python standard library only
# ------------------------------------------------------------------------------------------------
# The launcher (torchrun, or anything that exports RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT) tells every rank
# where rank 0 is (see _store_endpoint).  Rank 0 binds and accepts one connection per peer during init(); afterwards every
# host collective is one synchronous round on those sockets -- each rank sends its frame to rank 0, rank 0 answers
# with all frames in rank order -- served inside rank 0's own call (every rank makes the same sequence of calls, as with
# Horovod), so there is no background thread.
_STORE_TIMEOUT = float(os.environ.get("POLUS_STORE_TIMEOUT", "600"))


def _send_frame(sock, payload):
    sock.sendall(struct.pack("<Q", len(payload)) + payload)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(n - len(buf), 1 << 20))
        if not chunk:
            raise ConnectionError("polus store: peer closed the connection (another rank died?)")
        buf += chunk
    return bytes(buf)


def _recv_frame(sock):
    (n,) = struct.unpack("<Q", _recv_exact(sock, 8))
    return _recv_exact(sock, n)


def _store_endpoint():
    """(family, address) of rank 0's store.  One node (MASTER_ADDR is loopback -- the 8 GPUs of one box): an abstract
    Unix socket named after MASTER_PORT, which cannot collide with the launcher's own TCP store on that port or with
    any other listener.  Several nodes: TCP on POLUS_STORE_PORT, default MASTER_PORT + 1."""
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = _env_int("MASTER_PORT", 29500)
    kind = os.environ.get("POLUS_STORE", "unix" if addr in ("127.0.0.1", "localhost", "::1") else "tcp")
    if kind == "unix":
        return socket.AF_UNIX, "\0polus-store-%d-%s" % (port, os.environ.get("TORCHELASTIC_RUN_ID", ""))
    return socket.AF_INET, (addr, _env_int("POLUS_STORE_PORT", 0) or port + 1)


def _host_rendezvous_init():
    if os.environ.get("POLUS_RENDEZVOUS_DIR"):
        os.makedirs(os.environ["POLUS_RENDEZVOUS_DIR"], exist_ok=True)
        return
    family, where = _store_endpoint()
    rank, size = _state["rank"], _state["size"]
    tcp = family == socket.AF_INET
    if rank == 0:
        srv = socket.socket(family, socket.SOCK_STREAM)
        if tcp:
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        try:
            srv.bind(("", where[1]) if tcp else where)
        except OSError as e:
            raise RuntimeError(f"polus store: rank 0 cannot listen on {where!r} ({e}); set POLUS_STORE_PORT / POLUS_STORE") from e
        srv.listen(size)
        srv.settimeout(_STORE_TIMEOUT)
        peers = {}
        while len(peers) < size - 1:
            try:
                c, _ = srv.accept()
            except socket.timeout:
                raise TimeoutError(f"polus store: only {len(peers) + 1} of {size} ranks arrived within {_STORE_TIMEOUT:.0f} s")
            c.settimeout(_STORE_TIMEOUT)
            (r,) = struct.unpack("<I", _recv_exact(c, 4))
            peers[r] = c
        srv.close()
        _state["store"] = [peers[r] for r in range(1, size)]
    else:
        t0 = time.time()
        while True:
            c = socket.socket(family, socket.SOCK_STREAM)
            try:
                c.connect(where)
                break
            except OSError:
                c.close()
                if time.time() - t0 > _STORE_TIMEOUT:
                    raise TimeoutError(f"polus store: rank {rank} could not reach rank 0 at {where!r}")
                time.sleep(0.05)
        c.settimeout(_STORE_TIMEOUT)
        c.sendall(struct.pack("<I", rank))
        _state["store"] = c
    if tcp:
        for c in (_state["store"] if rank == 0 else [_state["store"]]):
            c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)


def _store_allgather(payload):
    if _state["rank"] == 0:
        frames = [payload] + [_recv_frame(c) for c in _state["store"]]
        blob = struct.pack("<I", len(frames)) + b"".join(struct.pack("<Q", len(f)) + f for f in frames)
        for c in _state["store"]:
            _send_frame(c, blob)
        return frames
    c = _state["store"]
    _send_frame(c, payload)
    blob = _recv_frame(c)
    (n,) = struct.unpack_from("<I", blob, 0)
    out, pos = [], 4
    for _ in range(n):
        (m,) = struct.unpack_from("<Q", blob, pos)
        out.append(blob[pos + 8:pos + 8 + m])
        pos += 8 + m
    return out


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
    if _state.get("store") is not None:
        return _store_allgather(payload)
    return _file_allgather(payload)


def _host_broadcast(payload: bytes, root=0):
    return _host_allgather(payload)[root]


def barrier():
    _host_allgather(b"")


def allgather_object(obj):
    """hvd.allgather_object (polus/callbacks.py:249): list with every rank's object, rank order.

    Device tensors inside `obj` (the int32 predictions ValidationDataCallback gathers every validation batch) travel
    over NVLink: one ncclAllGather per tensor into a [size, ...] buffer and ONE device-to-host read, instead of a
    device read + pickle + socket round trip per rank.  The picklable rest (host labels, shapes) rides on the host
    store; ranks whose tensors differ in shape (a ragged last batch) fall back to the host path for that call."""
    from .tensor import Tensor

    leaves = []

    def strip(o):
        if isinstance(o, Tensor):
            leaves.append(o)
            return ("__polus_tensor__", len(leaves) - 1, tuple(o.shape), o.dtype)
        if isinstance(o, (list, tuple)):
            return type(o)(strip(x) for x in o)
        if isinstance(o, dict):
            return {k: strip(v) for k, v in o.items()}
        return o

    skeleton = strip(obj)
    size = _state["size"]
    metas = [pickle.loads(b) for b in _host_allgather(pickle.dumps(skeleton))]
    sig = lambda m: pickle.dumps(_tensor_sigs(m))
    same = all(sig(m) == sig(metas[0]) for m in metas)
    if leaves and _state["nccl"] and same:
        st = device.stream()
        gathered = []
        for t in leaves:
            recv = Tensor((size,) + tuple(t.shape), t.dtype)
            _lib.call("polus_comm_allgather", t.ptr, recv.ptr, t.nbytes, st)
            gathered.append(recv.numpy())                      # [size, ...] on the host, one copy
        per_rank_leaves = [[g[r] for g in gathered] for r in range(size)]
    else:
        mine = pickle.dumps([t.numpy() for t in leaves])
        per_rank_leaves = [pickle.loads(b) for b in _host_allgather(mine)] if leaves else [[] for _ in range(size)]

    def fill(o, vals):
        if isinstance(o, tuple) and len(o) == 4 and o[0] == "__polus_tensor__":
            return vals[o[1]]
        if isinstance(o, (list, tuple)):
            return type(o)(fill(x, vals) for x in o)
        if isinstance(o, dict):
            return {k: fill(v, vals) for k, v in o.items()}
        return o
    return [fill(m, per_rank_leaves[r]) for r, m in enumerate(metas)]


def _tensor_sigs(o, out=None):
    out = [] if out is None else out
    if isinstance(o, tuple) and len(o) == 4 and o[0] == "__polus_tensor__":
        out.append((o[2], o[3]))
    elif isinstance(o, (list, tuple)):
        for x in o:
            _tensor_sigs(x, out)
    elif isinstance(o, dict):
        for v in o.values():
            _tensor_sigs(v, out)
    return out


# ------------------------------------------------------------------------------------------------
# device collectives
# ------------------------------------------------------------------------------------------------
def plan_broadcast_spans(variables):
    """[[ptr, nbytes], ...] in the order the ncclBroadcast calls are issued.

    The sequence of calls (and their sizes) must be the same on every rank.  Device addresses are NOT the same on every
    rank, so the spans keep the caller's variable order (arena creation order: fp32 masters first, then the bf16
    shadows) and only neighbours IN THAT ORDER that are also adjacent in memory are merged.  (Round 2: the spans used to
    be sorted by local address.  Models built after the first one of a process get their buffers from a fragmented free
    list, the master / shadow / optimizer-slot buffers then come out in a different address order on different ranks,
    and the ranks paired a 438 MB fp32 span with a 219 MB bf16 one, or Adam's m with v -- garbage weights, or a negative
    second moment, on the receiving ranks: every weight non-finite a few steps later; profiles/r02_nan_hunt.txt.)"""
    variables = list(variables)
    # (ptr, nbytes, owner): spans are merged only inside one allocation (an arena chunk's master or shadow buffer), so
    # that two allocations which happen to be neighbours on ONE rank cannot change that rank's call sequence
    spans = [(v.ptr, v.nbytes, ("p", v.chunk.index) if isinstance(v, Param) else ("t", id(getattr(v, "block", None)) or i))
             for i, v in enumerate(variables)]
    spans += [(v.shadow.ptr, v.shadow.nbytes, ("pb", v.chunk.index)) for v in variables if isinstance(v, Param)]
    merged, owner = [], None
    for ptr, n, own in spans:
        if merged and own == owner and 0 <= ptr - (merged[-1][0] + merged[-1][1]) <= 256 and (ptr - merged[-1][0]) % 4 == 0:
            merged[-1][1] = ptr + n - merged[-1][0]
        else:
            merged.append([ptr, n])
            owner = own
    return merged


def broadcast_variables(variables, root_rank=0):
    """hvd.broadcast_variables (polus/training.py:208-211): rank-0 values to every rank.  Adjacent
    arena variables are merged so BERT-base ships in a handful of ncclBroadcast calls."""
    if _state["size"] == 1 or not _state["nccl"]:
        return
    merged = plan_broadcast_spans(variables)
    st = device.stream()
    for ptr, n in merged:
        _lib.call("polus_comm_broadcast", ptr, n, root_rank, st)


def plan_buckets(weights, bucket_bytes=None):
    """Group the gradient arena spans of `weights` into contiguous buckets of ~bucket_bytes, ordered
    from the END of the arena (the last-created variables get their gradients first in backward).
    A variable of at least bucket_bytes / 2 (the word-embedding table: 94 MB fp32 in BERT-base, whose gradient backward
    produces LAST) is never merged with its neighbours: it closes the running bucket and is cut into its own
    ~bucket_bytes / 2 pieces, so that the variables before it are exchanged as soon as THEY are ready and the tail of
    the step pipelines allreduce(piece i+1) with the optimizer update of piece i instead of one long exposed exchange.
    Returns [(chunk, offset_elems, n_elems, [params])...]."""
    bucket_bytes = bucket_bytes or BUCKET_BYTES
    # (chunk creation index, offset): the same order on every rank -- id(chunk) is a host address and differs between
    # processes, which for a model of several chunks (BERT-large dims: three) made the ranks exchange different buckets
    ps = sorted((w for w in {id(w): w for w in weights if isinstance(w, Param)}.values()),
                key=lambda w: (w.chunk.index, w.offset), reverse=True)
    buckets = []
    cur = None
    for w in ps:
        n = (w.size + 63) & ~63
        if n * 4 >= bucket_bytes // 2:
            cur = None
            piece = max(64, ((bucket_bytes // 8) + 63) & ~63)          # elements per piece (bucket_bytes / 2 bytes)
            pieces = []
            pos = 0
            while pos < n:
                m = min(piece, n - pos)
                if n - pos - m < piece // 4:                           # no tiny last piece
                    m = n - pos
                pieces.append({"chunk": w.chunk, "off": w.offset + pos, "n": m, "params": [w]})
                pos += m
            buckets.extend(reversed(pieces))
            continue
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
            allreduce_bucket(ch, off, n, comm_stream)
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
