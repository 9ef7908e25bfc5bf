"""Device memory, streams and numpy<->device transfer for polus_b200 (thin layer over the C ABI)."""
import ctypes as C
import numpy as np

from . import _lib

_initialised = False
_default_stream = None  # None == legacy default stream until init() creates one


def init(device=0):
    """Pin this process to one GPU (reference: polus/__init__.py:122 set_visible_devices)."""
    global _initialised, _default_stream
    if _initialised:
        return
    _lib.call("polus_init", int(device))
    s = C.c_void_p()
    _lib.call("polus_stream_create", C.byref(s), 0)
    _default_stream = s.value
    _initialised = True


def stream():
    if not _initialised:
        init(0)
    return _default_stream


def synchronize():
    _lib.call("polus_stream_sync", stream())


def device_sync():
    _lib.call("polus_device_sync")


# ------------------------------------------------------------------ bf16 <-> fp32 on the host
def f32_to_bf16_bits(x):
    """Round-to-nearest-even fp32 -> bf16 storage (uint16), NaN preserved."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (0x7FFF + ((u >> 16) & 1))
    out = (rounded >> 16).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def bf16_bits_to_f32(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(x):
    return bf16_bits_to_f32(f32_to_bf16_bits(x))


_NP_OF = {_lib.F32: np.float32, _lib.BF16: np.uint16, _lib.I32: np.int32, _lib.U8: np.uint8}
_SIZE_OF = {_lib.F32: 4, _lib.BF16: 2, _lib.I32: 4, _lib.U8: 1}


def dtype_size(dt):
    return _SIZE_OF[dt]


class Buffer:
    """Owning handle of a device allocation; freed when collected."""

    __slots__ = ("ptr", "nbytes", "_owned")

    def __init__(self, nbytes, zero=False):
        if not _initialised:
            init(0)
        p = C.c_void_p()
        _lib.call("polus_malloc", C.byref(p), int(nbytes))
        self.ptr = p.value
        self.nbytes = int(nbytes)
        self._owned = True
        if zero and nbytes:
            _lib.call("polus_memset", self.ptr, 0, self.nbytes, stream())

    def __del__(self):
        try:
            if self._owned and self.ptr:
                _lib.load().polus_free(self.ptr)
        except Exception:
            pass


def upload(buf_ptr, arr):
    arr = np.ascontiguousarray(arr)
    _lib.call("polus_memcpy_h2d", buf_ptr, arr.ctypes.data, arr.nbytes, stream())
    synchronize()  # pageable source: make the call safe to return from


def download(buf_ptr, shape, np_dtype):
    out = np.empty(shape, dtype=np_dtype)
    _lib.call("polus_memcpy_d2h", out.ctypes.data, buf_ptr, out.nbytes, stream())
    synchronize()
    return out


class PinnedArray:
    """numpy view over pinned host memory (async H2D staging for input batches)."""

    def __init__(self, shape, np_dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(np_dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.call("polus_host_alloc", C.byref(p), max(n, 16))
        self.ptr = p.value
        self.nbytes = n
        raw = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().polus_host_free(self.ptr)
        except Exception:
            pass
