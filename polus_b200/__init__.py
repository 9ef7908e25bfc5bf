"""polus_b200 -- B200-native drop-in for the data-parallel training step of bioinformatics-ua/polus.

Host code is Python (as the reference is); all device work is hand-written CUDA for sm_100a in
libpolus_b200.so, bound through the C ABI of include/polus_b200.h.  There is no CPU fallback.

Module map (same names as the reference package so `import polus_b200 as polus` reads the same):
training, models, layers, losses, callbacks, data, metrics, core, utils, schedulers, ner, ir, mock.horovod.
"""
import logging
import os
import sys

__version__ = "0.2.1+b200.r1"

logger = logging.getLogger("polus")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s: %(message)s"))
    logger.addHandler(_h)
# reference: DEBUG by default, POLUS_LOGGER_LEVEL overrides (polus/__init__.py:66-73); unlike the
# reference no logs/ directory is created at import time.
logger.setLevel(os.environ.get("POLUS_LOGGER_LEVEL", "WARNING"))

from .utils import Singleton  # noqa: E402


class PolusContext(metaclass=Singleton):
    """Process/rank/device bootstrap (reference polus/__init__.py:102-125): multi-process data parallel is
    enabled when the launcher (torchrun / polus_b200.launch) set WORLD_SIZE > 1; the process is pinned
    to GPU `local_rank`."""

    def __init__(self):
        self.use_horovod = False
        world = int(os.environ.get("WORLD_SIZE", "1") or 1)
        if world > 1 and os.environ.get("POLUS_DISABLE_COMM", "0") != "1":
            from . import comm
            if comm.init(use_device=os.environ.get("POLUS_COMM_HOST_ONLY", "0") != "1") != "mock":
                if comm.local_rank() == 0:
                    logger.info(f"MultiGPU training enabled, using {comm.size()} processes ")
                self.use_horovod = True

    def is_horovod_enabled(self):
        return self.use_horovod


def hvd():
    """The collective module in force: polus_b200.comm under a multi-process launch, else the mock."""
    if PolusContext().is_horovod_enabled():
        from . import comm
        return comm
    from .mock import horovod
    return horovod
