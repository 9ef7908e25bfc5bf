"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): 2-rank NCCL data parallel == 1 rank on the global batch,
replicas bit-identical (polus/training.py:88-96,182-185,208-211).  The same check runs inside bench.py at N > 1 and is
printed in its JSON line as "dp_parity", so the driver's scaling run records it even when this test is skipped."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# bars: the two arms differ only in the order of fp32 additions (per-rank sums then NCCL sum vs one batch sum)
LOSS_REL = 2e-3
WEIGHT_MEAN_ABS = 2e-5


def check(report):
    assert report["replicas_identical"], report
    assert report["allgather_ok"], report                               # device tensors through ncclAllGather (callbacks.py:249)
    assert report["vs_single_rank_rel"] < LOSS_REL, report
    assert report["weights_mean_abs_diff"] < WEIGHT_MEAN_ABS, report
    assert report["weights_max_abs_diff"] <= 2.02 * report["lr"] * report["steps"], report
    assert abs(report["lr"] - 1e-3 * report["world"]) < 1e-9, report   # LR x size (training.py:90-94)


def test_two_rank_data_parallel_matches_single_rank_global_batch():
    from polus_b200 import _lib
    import ctypes as C
    n = C.c_int(0)
    _lib.call("polus_device_count", C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, POLUS_LOGGER_LEVEL="ERROR")
    multi = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dp_parity.py")],
                           capture_output=True, text=True, timeout=600, env=env)  # first `import torch` on a fresh box is slow
    assert multi.returncode == 0, multi.stderr[-3000:]
    lines = [l for l in multi.stdout.splitlines() if l.startswith("DPPARITY ")]
    assert len(lines) == 1, multi.stdout[-2000:]
    report = json.loads(lines[0].split(" ", 1)[1])
    assert report["world"] == 2
    check(report)


@pytest.mark.parametrize("arena_params", ["0", "40000000"])
def test_two_rank_rebuilt_models_stay_finite(arena_params):
    """A process that trains one model through a captured step, drops it and builds the next one (HPO loops; bench.py's
    parity harness followed by the timed model and two more configurations) must keep finite weights.  Round 2 found
    (profiles/r02_nan_hunt.txt) that models built after the first one of a process went non-finite on >= 2 GPUs: the
    broadcast spans were sorted by local device address and the buckets / optimizer slots by id(chunk), so ranks whose
    allocators had handed out the buffers in a different order issued different collective sequences.  The second case
    splits the arena into two chunks (POLUS_ARENA_PARAMS), the situation of a BERT-large-sized model."""
    from polus_b200 import _lib
    import ctypes as C
    import re
    n = C.c_int(0)
    _lib.call("polus_device_count", C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, POLUS_LOGGER_LEVEL="ERROR", NANHUNT_VARIANTS="base,base,base,base,base")
    if arena_params != "0":
        env["POLUS_ARENA_PARAMS"] = arena_params
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tools", "nan_hunt.py"), "32", "20"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, run.stderr[-3000:]
    reports = [json.loads(m) for m in re.findall(r"NANHUNT (\{.*?\})(?=NANHUNT|\n|$)", run.stdout)]
    assert len(reports) == 10, run.stdout[-2000:]          # 5 models x 2 ranks
    assert all(r["first_nonfinite_step"] is None for r in reports), reports


def test_dp_parity_harness_single_rank():
    """World of one: both arms are the same computation -- the harness itself must report exact agreement."""
    from polus_b200 import device
    from tests.dp_parity import run_dp_parity
    device.init(0)
    r = run_dp_parity(steps=3)
    assert r["world"] == 1 and r["replicas_identical"]
    assert r["vs_single_rank_rel"] < 1e-4 and r["weights_mean_abs_diff"] < 1e-6, r
