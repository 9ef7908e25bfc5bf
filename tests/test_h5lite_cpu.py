"""polus_b200.h5lite -- the `.h5` weight files of polus/models.py:42-48,127-133 without h5py.

Pinning: tests/golden/hdf5_library_written.mat is a file written by the real HDF5 library (MATLAB v7.3 = HDF5 behind a
512-byte user block; a copy of scipy's test datum scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD licence): the
reader must find its one dataset `testdouble` = 0, pi/4, ... 2 pi as float64 [9,1].  The writer is checked through the
reader, structurally against the specification's invariants (alignment, sorted B-tree keys, end-of-file address), and
against the exact structure the library-written file uses (same superblock version, group style and message classes)."""
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reader_on_a_file_written_by_the_hdf5_library():
    from polus_b200 import h5lite
    d = h5lite.read_h5(os.path.join(HERE, "golden", "hdf5_library_written.mat"))
    assert list(d) == ["testdouble"]
    a = d["testdouble"]
    assert a.dtype == np.float64 and a.shape == (9, 1)
    np.testing.assert_allclose(a[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_weight_file_round_trip_and_layout(tmp_path):
    from polus_b200 import h5lite
    rng = np.random.default_rng(0)
    ws = [rng.standard_normal((768, 64)).astype(np.float32), rng.standard_normal(64).astype(np.float32),
          np.zeros((0,), np.float32), np.arange(6, dtype=np.int32).reshape(2, 3), rng.standard_normal((2, 2, 2)),
          np.float32(3.5)] + [rng.standard_normal(5).astype(np.float32) for _ in range(400)]   # > one symbol node
    path = str(tmp_path / "m.h5")
    h5lite.write_weights(path, ws)
    back = h5lite.read_weights(path)
    assert len(back) == len(ws)
    for a, b in zip(back, ws):
        b = np.asarray(b)
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8        # superblock v0, 8-byte offsets
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                                          # end-of-file address
    names = list(h5lite.read_h5(path))
    assert names == sorted(names)                                                                    # group B-tree order = strcmp order
    # same structural choices as the library-written file
    ref = open(os.path.join(HERE, "golden", "hdf5_library_written.mat"), "rb").read()[512:]
    assert ref[8:16:1][:1] == raw[8:9] and ref[13:15] == raw[13:15]                                  # version, offset / length sizes
    assert b"TREE" in raw and b"HEAP" in raw and b"SNOD" in raw
    for sig in (b"TREE", b"HEAP", b"SNOD"):
        assert raw.index(sig) % 8 == 0                                                               # every structure 8-byte aligned


def test_reader_refuses_what_it_does_not_implement(tmp_path):
    from polus_b200 import h5lite
    bad = tmp_path / "x.h5"
    bad.write_bytes(b"not hdf5 at all" * 100)
    with pytest.raises(h5lite.H5Error):
        h5lite.read_h5(str(bad))
    with pytest.raises(h5lite.H5Error):
        h5lite.write_h5(str(tmp_path / "y.h5"), [("a", np.array(["text"]))])
    with pytest.raises(h5lite.H5Error):
        h5lite.write_h5(str(tmp_path / "z.h5"), [("a", np.zeros(2)), ("a", np.zeros(2))])
