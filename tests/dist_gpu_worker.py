"""Worker for tests/test_dist_gpu.py (run under torchrun, one process per GPU): data-parallel training of a tiny
BERT-NER model; every rank prints its losses and a weight checksum.  With Horovod semantics (polus/training.py:
88-96,182-185) N ranks on N half-batches with lr*N must follow the same trajectory as ONE process on the
concatenated batch with lr*N (the loss is a batch mean, gradients are averaged)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import polus_b200  # noqa: E402
from polus_b200 import device, ops, tensor  # noqa: E402
from polus_b200.models import BertConfig  # noqa: E402
from polus_b200.ner.models import BertNERModel  # noqa: E402
from polus_b200.optimizers import Adam  # noqa: E402
from polus_b200.training import ClassifierTrainer  # noqa: E402
from polus_b200.utils import set_random_seed  # noqa: E402
from tests.parity import make_batch  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
device.init(int(os.environ.get("LOCAL_RANK", rank)))
ctx = polus_b200.PolusContext()
set_random_seed(11)
cfg = BertConfig(vocab_size=300, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                 max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
model = BertNERModel(cfg, output_classes=4, droupout_p=0.0)
rng = np.random.default_rng(5)
GB, S = 8, 32
ids, mask, tt, tags = make_batch(rng, GB, S, 300, 4)
total = int(sys.argv[1]) if len(sys.argv) > 1 else world  # emulate `total` ranks worth of LR scaling
lo, hi = (rank * GB // world, (rank + 1) * GB // world)
x = {"input_ids": ids[lo:hi], "attention_mask": mask[lo:hi], "token_type_ids": tt[lo:hi]}
y = np.eye(4, dtype=np.float32)[tags[lo:hi]]
opt = Adam(1e-3 * (total if world == 1 else 1))  # the trainer multiplies by hvd.size() itself when world > 1
trainer = ClassifierTrainer(model, opt, model.loss)
if trainer.use_horovod:
    model(**x, training=False)
    trainer.trainable_weights = model.trainable_weights
    trainer.broadcast_init_vars()
losses = [float(trainer.train_step(x, y)) for _ in range(4)]
chk = float(sum(np.abs(w.numpy()).sum() for w in model.weights))
w0 = model.hidden.kernel.numpy().reshape(-1)[:8].tolist()
print("DPRESULT " + json.dumps({"rank": rank, "world": world, "losses": losses, "checksum": chk, "w0": w0,
                                "lr": float(opt.learning_rate.read_value())}), flush=True)
