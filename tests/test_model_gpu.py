"""GPU parity of the assembled path through the public polus-shaped API (ClassifierTrainer.train_step):
emissions, loss and gradients of a tiny BERT-NER+CRF vs the CPU oracle, eager step vs captured-graph
replay, and the tutorial classifier (config #1 of BASELINE.json) reaching macro-F1 > 0.9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tiny_bert_ner_crf_matches_oracle():
    from tests.parity import run_tiny_ner_parity
    r = run_tiny_ner_parity(steps=4)
    assert r["emis_ok"], r
    assert abs(r["loss_dev"] - r["loss_ref"]) / abs(r["loss_ref"]) < 1e-2, r
    assert r["min_grad_cos"] >= 0.999, r          # gradient cosine >= 0.999
    assert r["loss_traj_rel"] < 1e-2, r           # loss after N Adam steps within 1e-2 rel


def test_bert_shapes_seq256_heads12():
    """The BASELINE shapes (H=768, 12 heads, S=256) on one layer, forward+backward vs oracle."""
    from tests.parity import run_tiny_ner_parity
    r = run_tiny_ner_parity(steps=2, B=2, S=256, H=768, nh=12, I=3072, L=1, vocab=2000)
    assert r["ok"], r


def test_graph_replay_equals_eager(monkeypatch):
    """The captured CUDA graph must reproduce the op-by-op step bit for bit (same kernels, same order)."""
    import polus_b200
    from polus_b200 import device, ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed
    from tests.parity import make_batch

    def run(eager):
        monkeypatch.setenv("POLUS_EAGER", "1" if eager else "0")
        tensor.reset_arena()
        set_random_seed(3)
        ops.set_step(0)
        cfg = BertConfig(vocab_size=500, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                         max_position_embeddings=64)  # dropout 0.1 ON: replay must regenerate the same masks
        model = BertNERModel(cfg, output_classes=4)
        rng = np.random.default_rng(0)
        ids, mask, tt, tags = make_batch(rng, 4, 32, 500, 4)
        x = {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}
        y = np.eye(4, dtype=np.float32)[tags]
        tr = ClassifierTrainer(model, Adam(1e-3), model.loss)
        return [float(tr.train_step(x, y)) for _ in range(5)], [w.numpy() for w in model.weights]
    le, we = run(True)
    lg, wg = run(False)
    # Same kernels, same order, same Philox counters: the trajectories agree up to the order of fp32 atomic
    # accumulations (split-K TMA reduce-add, embedding scatter-add).  Adam turns a sign flip of a near-zero
    # gradient into a +-lr step, so individual weights may differ by a few lr; the bulk must agree tightly.
    np.testing.assert_allclose(lg, le, rtol=1e-3)
    for a, b in zip(wg, we):
        d = np.abs(a - b)
        assert d.max() <= 5e-3 and d.mean() <= 2e-4, (d.max(), d.mean())


@pytest.mark.gpu
def test_graph_replay_gradients_equal_eager(monkeypatch):
    """Sharp check of the captured step's scheduling (programmatic dependent launches, weight gradients on the
    background stream, csrc/common.cuh + ops.side_fork): with lr = 0 the weights never move, so the gradients of
    every step are a fixed function of (weights, batch, Philox counters) and Adam's first moment m after 4 steps
    must agree between op-by-op execution and graph replay up to the order of fp32 atomic accumulation.  A kernel
    that ran ahead of its producer (or a wgrad that read a recycled buffer) shows up as an O(1) relative error."""
    from polus_b200 import ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed
    from tests.parity import make_batch

    def run(eager):
        monkeypatch.setenv("POLUS_EAGER", "1" if eager else "0")
        tensor.reset_arena()
        set_random_seed(5)
        ops.set_step(0)
        cfg = BertConfig(vocab_size=1000, hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=1024,
                         max_position_embeddings=128)
        model = BertNERModel(cfg, output_classes=4)
        rng = np.random.default_rng(0)
        ids, mask, tt, tags = make_batch(rng, 8, 128, 1000, 4)
        x = {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}
        y = np.eye(4, dtype=np.float32)[tags]
        opt = Adam(0.0)
        tr = ClassifierTrainer(model, opt, model.loss)
        losses = [float(tr.train_step(x, y)) for _ in range(4)]
        m = opt.variables()[0].numpy()
        return losses, m
    le, me = run(True)
    lg, mg = run(False)
    np.testing.assert_allclose(lg, le, rtol=1e-5)
    scale = np.abs(me).max()
    assert scale > 0
    assert np.abs(mg - me).max() <= 2e-4 * scale, (np.abs(mg - me).max(), scale)


def test_tutorial_classifier_macro_f1():
    """Config #1: tutorials/classifier_example.py (784->128 relu->10, Adam 1e-3, batch 128 drop_remainder) on
    synthetic class-dependent blobs instead of MNIST (no network); reference assertion: macro-F1 > 0.9
    (tests/test_integration.py:17-20)."""
    from polus_b200 import nn, tensor
    from polus_b200.callbacks import EarlyStop, LossSmoothCallback, TimerCallback, ValidationDataCallback, ConsoleLogCallback
    from polus_b200.data import DataLoader
    from polus_b200.losses import SparseCategoricalCrossentropy
    from polus_b200.metrics import MacroF1Score
    from polus_b200.models import SequentialPolusClassifier
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed

    tensor.reset_arena()
    set_random_seed(0)
    rng = np.random.default_rng(0)
    centers = rng.uniform(0, 255, (10, 28, 28))

    def make(n):
        y = rng.integers(0, 10, n)
        x = np.clip(centers[y] + rng.normal(0, 60, (n, 28, 28)), 0, 255).astype(np.uint8)
        return x, y

    x_train, y_train = make(4096)
    x_test, y_test = make(1024)

    def train_gen(dx, dy):
        def generator():
            for i in range(len(dx)):
                yield {"x": dx[i], "y": dy[i]}
        return generator

    norm = lambda d: (d["x"].astype(np.float32) / 255.0, np.int32(d["y"]))
    ds_train = DataLoader(train_gen(x_train, y_train)).to_tfDataset().map(norm).cache().shuffle(len(x_train), seed=0).batch(128, drop_remainder=True).prefetch(-1)
    ds_test = DataLoader(train_gen(x_test, y_test)).to_tfDataset().map(norm).batch(128).cache().prefetch(-1)
    model = SequentialPolusClassifier([nn.Flatten(input_shape=(28, 28)), nn.Dense(128, activation='relu'), nn.Dense(10)])
    trainer = ClassifierTrainer(model, Adam(0.001), SparseCategoricalCrossentropy(from_logits=True),
                                metrics=[MacroF1Score(num_classes=10)])
    val = ValidationDataCallback(ds_test, name="Synthetic_Test")
    callbacks = [LossSmoothCallback(output=True), TimerCallback(), val, ConsoleLogCallback(log_interval=1000), EarlyStop()]
    trainer.changing_train_config(tf_dataset=ds_train, epochs=5, callbacks=callbacks)
    trainer.train()
    f1 = val.get_metrics()["MacroF1Score"][-1]
    assert f1 > 0.9, f1


def test_host_feed_ring_keeps_batches_apart_when_host_runs_ahead(monkeypatch):
    """The input feed (training._CompiledStep.feed): pinned staging ring + copy stream.  Feeding 10 DIFFERENT host batches
    without ever waiting for a loss must train on exactly the batches a fully synchronous loop trains on."""
    from polus_b200 import ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed
    from tests.parity import make_batch

    rng = np.random.default_rng(9)
    batches = []
    for _ in range(10):
        ids, mask, tt, tags = make_batch(rng, 4, 64, 800, 4)
        batches.append(({"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}, np.eye(4, dtype=np.float32)[tags]))

    def run(sync_every_step):
        monkeypatch.setenv("POLUS_EAGER", "0")
        tensor.reset_arena()
        set_random_seed(11)
        ops.set_step(0)
        cfg = BertConfig(vocab_size=800, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                         max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = BertNERModel(cfg, output_classes=4, droupout_p=0.0)
        tr = ClassifierTrainer(model, Adam(1e-3), model.loss)
        losses = []
        for x, y in batches:
            l = tr.train_step(x, y)
            losses.append(float(l) if sync_every_step else l)
        return [float(l) for l in losses]
    np.testing.assert_allclose(run(False), run(True), rtol=2e-3)


def test_full_size_bert_base_ner_properties():
    """BASELINE config 2 at its FULL size (BERT-base: 12 layers, H=768, 12 heads, I=3072, vocab 30522; S=256, K=4), checked
    through size-independent properties, because the numpy oracle cannot run 110 M parameters in test time:
      * Viterbi paths and the CRF negative log-likelihood are recomputed by the oracle FROM THE DEVICE'S OWN EMISSIONS:
        paths bit-exact (north_star), loss to fp32 tolerance;
      * the loss is a batch mean: loss(batch) == mean(loss(first half), loss(second half)) with dropout off;
      * inference is deterministic and independent of the other rows of the batch (row i alone == row i in the batch);
      * padded key positions do not influence real tokens: changing input_ids under attention_mask == 0 leaves the
        emissions of the real tokens unchanged;
      * five Adam steps through the captured-graph trainer on one fixed batch lower the loss."""
    from oracle import numpy_ref as R
    from polus_b200 import device, ops, tensor
    from polus_b200.models import BertConfig
    from polus_b200.ner.models import BertNERModel
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed
    from tests.parity import make_batch
    device.init(0)
    tensor.reset_arena()
    set_random_seed(123)
    ops.set_step(0)
    cfg = BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)  # BERT-base defaults otherwise
    K, B, S = 4, 8, 256
    model = BertNERModel(cfg, output_classes=K, hidden_space=128, droupout_p=0.0)
    rng = np.random.default_rng(123)
    ids, mask, tt, tags = make_batch(rng, B, S, cfg.vocab_size, K)
    x = {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}
    y = np.eye(K, dtype=np.float32)[tags]
    model(**x, training=False)
    # 109,581,204 = HF BertModel (109,482,240) + NER head + CRF; the pooler Dense (768*768 + 768) gets no gradient
    # on this path (polus/models.py:216 bypasses it) and is not instantiated here
    assert sum(int(np.prod(w.shape)) for w in model.trainable_weights) == 109_581_204 - (768 * 768 + 768)
    # spread the emissions (random-init heads give nearly flat ones: every path would tie)
    out_kernel0, trans0 = model.out.kernel.numpy().copy(), model.crf.transitions.numpy().copy()
    model.out.kernel.assign(out_kernel0 * 30.0)
    model.crf.transitions.assign(rng.standard_normal((K, K)).astype(np.float32))

    emis32 = model.emissions(**x, training=False).numpy().astype(np.float32)
    emis = emis32.astype(np.float64)
    assert np.isfinite(emis).all() and emis.std() > 0.05
    trans32 = model.crf.transitions.numpy().astype(np.float32)
    trans = trans32.astype(np.float64)
    lens = np.full(B, S)  # polus/layers.py:74-76: the layer decodes every row at full length
    # Viterbi: bit-exact given identical emissions (fp32 sums in the same order on both sides, ties -> lowest index)
    paths_dev = model.inference(x).numpy()
    paths_ref, _ = R.crf_decode(emis32, lens, trans32)
    assert paths_dev.shape == (B, S) and np.array_equal(paths_dev.astype(np.int64), np.asarray(paths_ref).astype(np.int64))
    assert len(np.unique(paths_dev)) > 1
    # CRF loss from the same emissions
    loss_dev = float(model.loss(tensor.Tensor.from_numpy(y, tensor.F32), model.crf(model.emissions(**x, training=True), training=True)))
    loss_ref = float(np.mean(-R.crf_log_likelihood(emis, tags, lens, trans)))
    assert abs(loss_dev - loss_ref) <= 2e-3 * abs(loss_ref), (loss_dev, loss_ref)

    def loss_of(sl):
        xs = {k: v[sl] for k, v in x.items()}
        e = model.emissions(**xs, training=True)
        return float(model.loss(tensor.Tensor.from_numpy(y[sl], tensor.F32), model.crf(e, training=True)))
    whole, h0, h1 = loss_of(slice(0, B)), loss_of(slice(0, B // 2)), loss_of(slice(B // 2, B))
    assert abs(whole - 0.5 * (h0 + h1)) <= 1e-4 * abs(whole), (whole, h0, h1)

    # rows are independent; inference is deterministic
    e_row = model.emissions(**{k: v[3:4] for k, v in x.items()}, training=False).numpy()
    e_all = model.emissions(**x, training=False).numpy()
    assert np.array_equal(e_all, model.emissions(**x, training=False).numpy())
    np.testing.assert_allclose(e_row[0], e_all[3], atol=2e-2, rtol=2e-2)  # (batch 1 and batch 8 take different GEMM tile paths)
    # tokens under the padding mask are invisible to the real ones
    ids2 = ids.copy()
    ids2[mask == 0] = rng.integers(0, cfg.vocab_size, size=int((mask == 0).sum()))
    e_pad = model.emissions(input_ids=ids2, attention_mask=mask, token_type_ids=tt, training=False).numpy()
    assert np.array_equal(e_pad[mask == 1], e_all[mask == 1])

    model.out.kernel.assign(out_kernel0)  # back to the initial head for the optimisation check
    model.crf.transitions.assign(trans0)
    trainer = ClassifierTrainer(model, Adam(2e-5), model.loss)
    losses = [float(trainer.train_step(x, y)) for _ in range(5)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
