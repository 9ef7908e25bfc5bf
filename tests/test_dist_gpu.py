"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): 2-rank NCCL data parallel == 1 rank on the global batch."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _results(out):
    return [json.loads(l.split("DPRESULT ", 1)[1]) for l in out.splitlines() if l.startswith("DPRESULT ")]


def test_two_rank_data_parallel_matches_single_rank_global_batch():
    from polus_b200 import _lib
    import ctypes as C
    n = C.c_int(0)
    _lib.call("polus_device_count", C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    worker = os.path.join(ROOT, "tests", "dist_gpu_worker.py")
    env = dict(os.environ, POLUS_LOGGER_LEVEL="ERROR")
    single = subprocess.run([sys.executable, worker, "2"], capture_output=True, text=True, timeout=300, env=env)
    assert single.returncode == 0, single.stderr[-2000:]
    ref = _results(single.stdout)[0]
    multi = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", worker],
                           capture_output=True, text=True, timeout=900, env=env)  # first `import torch` on a fresh box is slow
    assert multi.returncode == 0, multi.stderr[-3000:]
    res = sorted(_results(multi.stdout), key=lambda r: r["rank"])
    assert len(res) == 2 and abs(res[0]["lr"] - 2e-3) < 1e-9
    # replicas stay bit-identical (same averaged gradients, same update)
    assert res[0]["checksum"] == res[1]["checksum"] and res[0]["w0"] == res[1]["w0"]
    # mean of the two half-batch losses == global-batch loss, step by step; weights follow the same trajectory
    mean_losses = np.mean([res[0]["losses"], res[1]["losses"]], axis=0)
    np.testing.assert_allclose(mean_losses, ref["losses"], rtol=2e-3)
    np.testing.assert_allclose(res[0]["w0"], ref["w0"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(res[0]["checksum"], ref["checksum"], rtol=1e-4)
