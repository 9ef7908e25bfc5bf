"""numpy restatement of the device dropout RNG (Philox4x32-10, Salmon et al. SC'11) -- TEST INFRASTRUCTURE ONLY.

The reference draws dropout masks from TensorFlow's stateful RNG (tf.nn.dropout inside HF TFBertLayer and
tf.keras.layers.Dropout, polus/ner/models.py:58); that stream cannot be reproduced without TF,
so parity runs either use p=0 or regenerate the *device's* masks here bit-for-bit and feed them to the oracle.
Counter = (idx_lo, idx_hi, site, step), key = (seed_lo, seed_hi); one call covers 8 consecutive elements,
16 random bits each: keep <=> bits >= round(p*65536)."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, np.uint32).copy() for c in (c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    for _ in range(10):
        p0 = _M0 * c0.astype(np.uint64)
        p1 = _M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        with np.errstate(over="ignore"):
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_keep_mask(n_elems, p, seed, site, step):
    """Boolean keep mask for a flat tensor of n_elems (multiple of 8), identical to the device's."""
    assert n_elems % 8 == 0
    n8 = n_elems // 8
    idx = np.arange(n8, dtype=np.uint64)
    r = philox4x32_10((idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                      np.full(n8, site, np.uint32), np.full(n8, step, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    thresh = np.uint32(int(np.rint(np.float32(p) * np.float32(65536.0))))
    keep = np.empty((n8, 8), bool)
    for i, w in enumerate(r):
        keep[:, 2 * i] = (w & np.uint32(0xFFFF)) >= thresh
        keep[:, 2 * i + 1] = (w >> np.uint32(16)) >= thresh
    return keep.reshape(-1)


def dropout_scale_mask(shape, p, seed, site, step):
    """Pre-scaled mask (0 or 1/(1-p)) as float32, in the layout the oracle's `masks` arguments expect."""
    keep = dropout_keep_mask(int(np.prod(shape)), p, seed, site, step).reshape(shape)
    return keep.astype(np.float32) * np.float32(1.0 / (1.0 - p))
