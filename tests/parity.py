"""Shared parity harness: tiny BERT-NER+CRF on the device vs the CPU oracle (used by the GPU tests and by
__graft_entry__.smoke()).  Tolerances are the ones BASELINE.md states for the bf16 path: loss rel 1e-2,
emissions atol/rtol 2e-2, gradient cosine >= 0.999 per tensor."""
import numpy as np


def make_batch(rng, B, S, vocab, K, ragged=True):
    ids = rng.integers(0, vocab, size=(B, S)).astype(np.int32)
    lens = rng.integers(S // 4, S + 1, size=B) if ragged else np.full(B, S)
    mask = (np.arange(S)[None, :] < lens[:, None]).astype(np.int32)
    tt = np.zeros((B, S), np.int32)
    tags = rng.integers(1, K, size=(B, S)).astype(np.int32) * mask  # PAD = 0 outside the length
    return ids, mask, tt, tags


def device_params_to_oracle(model):
    from polus_b200 import device
    emb = model.bert.bert.embeddings
    r = lambda p: p.numpy().astype(np.float64)
    params = {"emb": {"word": r(emb.word), "pos": r(emb.position), "type": r(emb.token_type),
                      "emb_ln_g": r(emb.ln_gamma), "emb_ln_b": r(emb.ln_beta)},
              "layers": [], "head": {"Wa": r(model.hidden.kernel), "ba": r(model.hidden.bias),
                                     "Wb": r(model.out.kernel), "bb": r(model.out.bias)},
              "trans": r(model.crf.transitions)}
    for l in model.bert.bert.encoder.layer:
        params["layers"].append({k: r(getattr(l, k)) for k in
                                 ("Wqkv", "bqkv", "Wo", "bo", "ln1_g", "ln1_b", "W1", "b1", "W2", "b2", "ln2_g", "ln2_b")})
    return params


def device_grads(model):
    emb = model.bert.bert.embeddings
    r = lambda p: p.grad.numpy().astype(np.float64)
    g = {"emb": {"word": r(emb.word), "pos": r(emb.position), "type": r(emb.token_type),
                 "emb_ln_g": r(emb.ln_gamma), "emb_ln_b": r(emb.ln_beta)},
         "layers": [], "head": {"Wa": r(model.hidden.kernel), "ba": r(model.hidden.bias),
                                "Wb": r(model.out.kernel), "bb": r(model.out.bias)},
         "trans": r(model.crf.transitions)}
    for l in model.bert.bert.encoder.layer:
        g["layers"].append({k: r(getattr(l, k)) for k in
                            ("Wqkv", "bqkv", "Wo", "bo", "ln1_g", "ln1_b", "W1", "b1", "W2", "b2", "ln2_g", "ln2_b")})
    return g


def round_weights_to_bf16(model):
    """Make the fp32 masters exactly bf16-representable so device (bf16 GEMM operands) and oracle start
    from identical values."""
    from polus_b200 import device
    for w in model.weights:
        w.assign(device.bf16_round(w.numpy()))


def run_tiny_ner_parity(steps=3, verbose=False, B=4, S=64, H=128, nh=2, I=512, L=2, vocab=1000, K=4, lr=1e-3, seed=7):
    import polus_b200
    from oracle import ner_model as O
    from polus_b200 import _lib, device, nn, ops, tensor
    from polus_b200.ner.models import BertNERModel
    from polus_b200.models import BertConfig
    from polus_b200.optimizers import Adam
    from polus_b200.training import ClassifierTrainer
    from polus_b200.utils import set_random_seed

    device.init(0)
    tensor.reset_arena()
    set_random_seed(seed)
    ops.set_step(0)
    cfg = BertConfig(vocab_size=vocab, hidden_size=H, num_hidden_layers=L, num_attention_heads=nh, intermediate_size=I,
                     max_position_embeddings=max(S, 64), hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = BertNERModel(cfg, output_classes=K, hidden_space=128, droupout_p=0.0)
    rng = np.random.default_rng(seed)
    ids, mask, tt, tags = make_batch(rng, B, S, vocab, K)
    x = {"input_ids": ids, "attention_mask": mask, "token_type_ids": tt}
    y = np.eye(K, dtype=np.float32)[tags]
    # build lazily-created variables with one inference call, then align weights
    model(**x, training=False)
    # non-trivial LayerNorm / bias values so their gradients paths are exercised
    for w in model.weights:
        if w.name.endswith(("gamma",)):
            w.assign(1.0 + 0.1 * rng.standard_normal(w.shape))
        elif w.name.endswith(("beta", "bias")):
            w.assign(0.05 * rng.standard_normal(w.shape))
    round_weights_to_bf16(model)
    params = device_params_to_oracle(model)

    # ---- gradients of the first step (tape only, no optimizer)
    launches0 = _lib.call("polus_launch_count")
    with ops.GradientTape() as tape:
        e_dev = model.emissions(**x, training=True)
        loss_dev_t = model.loss(y, model.crf(e_dev, training=True))
    tape.gradient(loss_dev_t, model.trainable_weights)
    loss_dev = float(loss_dev_t)
    e_dev_h = e_dev.numpy()
    g_dev = O.flatten(device_grads(model))
    launches = _lib.call("polus_launch_count") - launches0
    for w in model.weights:  # clear the arena gradients before training
        _lib.call("polus_memset", w.grad.ptr, 0, w.grad.nbytes, device.stream())
    loss_ref, e_ref, g_ref = O.loss_and_grads(params, ids, mask, tt, tags, nh)
    g_ref = O.flatten(g_ref)
    cos_min, rel_max, worst = 1.0, 0.0, None
    for k, gr in g_ref.items():
        gd = g_dev[k]
        denom = np.linalg.norm(gr) * np.linalg.norm(gd)
        cos = float((gr * gd).sum() / denom) if denom > 0 else 1.0
        rel = float(np.abs(gr - gd).max() / (np.abs(gr).max() + 1e-12))
        if cos < cos_min:
            cos_min, worst = cos, k
        rel_max = max(rel_max, rel)
    emis_err = float(np.abs(e_dev_h - e_ref).max())
    emis_ok = bool(np.allclose(e_dev_h, e_ref, atol=2e-2, rtol=2e-2))

    # ---- N optimisation steps through the public trainer (step 1 eager, step 2 captured, then replays)
    trainer = ClassifierTrainer(model, Adam(lr), model.loss)
    dev_losses = [float(trainer.train_step(x, y)) for _ in range(steps)]
    ref_losses, _ = O.adam_train(params, [(ids, mask, tt, tags)], nh, lr=lr, steps=steps)
    loss_rel = max(abs(a - b) / max(abs(b), 1e-6) for a, b in zip(dev_losses, ref_losses))
    report = dict(loss_dev=loss_dev, loss_ref=float(loss_ref), emis_max_abs_err=emis_err, emis_ok=emis_ok,
                  min_grad_cos=cos_min, worst_grad=worst, max_rel_grad_err=rel_max, dev_losses=dev_losses,
                  ref_losses=ref_losses, loss_traj_rel=loss_rel, launches=int(launches))
    report["ok"] = bool(emis_ok and cos_min >= 0.999 and abs(loss_dev - loss_ref) / abs(loss_ref) < 1e-2 and loss_rel < 1e-2)
    if verbose:
        print(report)
    return report
