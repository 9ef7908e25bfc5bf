"""One op-by-op BERT-base NER training step inside a cudaProfilerStart/Stop window (for ncu launch lists):
    python tools/profile_step.py && ncu --profile-from-start off --metrics gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
Prints the step's own CUDA-event time so the kernel SHARES of the ncu list can be cross-checked."""
import ctypes as C
import os
import sys

os.environ["POLUS_EAGER"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from polus_b200 import _lib, device  # noqa: E402
from polus_b200.models import BertConfig  # noqa: E402
from polus_b200.ner.models import BertNERModel  # noqa: E402
from polus_b200.optimizers import Adam  # noqa: E402
from polus_b200.schedulers import warmup_scheduler  # noqa: E402
from polus_b200.training import ClassifierTrainer  # noqa: E402
from polus_b200.utils import set_random_seed  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
device.init(0)
workload = os.environ.get("PS_WORKLOAD", "ner_base")   # ner_base | cfg4 | cfg5 (bench.build_workload)
trainer, batches, _, label, _ = bench.build_workload(workload, batch)
for i in range(3):
    float(trainer.train_step(*batches[i % len(batches)]))
e0, e1 = C.c_void_p(), C.c_void_p()
_lib.call("polus_event_create", C.byref(e0)); _lib.call("polus_event_create", C.byref(e1))
device.device_sync()
_lib.call("polus_profiler_start")
_lib.call("polus_event_record", e0, device.stream())
loss = trainer.train_step(*batches[0])
_lib.call("polus_event_record", e1, device.stream())
float(loss)
device.device_sync()
_lib.call("polus_profiler_stop")
ms = C.c_float()
_lib.call("polus_event_elapsed_ms", e0, e1, C.byref(ms))
print(f"profiled step (op-by-op, {label}): {ms.value:.3f} ms, loss {float(loss):.4f}")
