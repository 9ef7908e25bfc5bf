"""torch-CPU restatement of the reference training step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference's own CPU path is TensorFlow (HF TFBertModel + polus.ner head + tfa CRF + Keras Adam,
polus/training.py:150-193); TensorFlow is not installable in this image, so bench.py's
`cpu_baseline` and `--impl reference` legs time this port: the same HuggingFace BERT architecture from the
same `transformers` library (torch classes, eager attention, additive -10000 mask as polus/models.py:190-193),
the NER head of polus/ner/models.py:26-67, the CRF negative log-likelihood of tfa.text.crf_log_likelihood
(polus/layers.py:86-99, full-length sequences) and Adam with Keras's epsilon (1e-7).  fp32, all host threads.
"""
import os
import time

import numpy as np


def build(hidden=768, layers=12, heads=12, inter=3072, vocab=30522, max_pos=512, K=4, head_hidden=128, dropout=0.1):
    import torch
    from transformers import BertConfig
    from transformers.models.bert.modeling_bert import BertEmbeddings, BertLayer

    cfg = BertConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=inter, max_position_embeddings=max_pos, hidden_dropout_prob=dropout,
                     attention_probs_dropout_prob=dropout, layer_norm_eps=1e-12)
    cfg._attn_implementation = "eager"

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = BertEmbeddings(cfg)
            self.layers = torch.nn.ModuleList([BertLayer(cfg) for _ in range(layers)])
            self.drop = torch.nn.Dropout(dropout)
            self.fc1 = torch.nn.Linear(hidden, head_hidden)
            self.fc2 = torch.nn.Linear(head_hidden, K)
            self.trans = torch.nn.Parameter(torch.empty(K, K).uniform_(-0.5, 0.5))

        def forward(self, ids, mask, tt, tags):
            add = (1.0 - mask.float())[:, None, None, :] * -10000.0
            h = self.emb(input_ids=ids, token_type_ids=tt)
            for l in self.layers:
                h = l(h, attention_mask=add)
                h = h[0] if isinstance(h, tuple) else h
            e = self.fc2(torch.nn.functional.silu(self.fc1(self.drop(h))))
            B, T, _ = e.shape
            alpha = e[:, 0]
            for t in range(1, T):
                alpha = e[:, t] + torch.logsumexp(alpha[:, :, None] + self.trans[None], dim=1)
            logz = torch.logsumexp(alpha, dim=1)
            score = e.gather(2, tags[:, :, None]).squeeze(2).sum(1) + self.trans[tags[:, :-1], tags[:, 1:]].sum(1)
            return (logz - score).mean()

    return Net()


def time_train_steps(batch=8, seq=256, steps=2, warmup=1, threads=None, **model_kw):
    """Returns (sequences/s, seconds per step, threads used)."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = build(**model_kw)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=5e-5, eps=1e-7)
    vocab = model_kw.get("vocab", 30522)
    K = model_kw.get("K", 4)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, vocab, (batch, seq), generator=g)
    mask = torch.ones(batch, seq, dtype=torch.long)
    tt = torch.zeros(batch, seq, dtype=torch.long)
    tags = torch.randint(1, K, (batch, seq), generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = net(ids, mask, tt, tags)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, threads


# ------------------------------------------------------------------------------------------------------------------
# Parity helpers (tests/test_fullsize_gpu.py): the SAME weights on both sides, loss / emissions / gradients / N Keras-Adam
# steps of the full 12-layer model in fp32 on the CPU.
# ------------------------------------------------------------------------------------------------------------------
def load_weights(net, hf_state, head):
    """hf_state: HuggingFace-named numpy state dict of the encoder (polus_b200.pretrained.export_hf_bert_weights layout);
    head: {"Wa" [H,h], "ba", "Wb" [h,K], "bb", "trans" [K,K]} in the Keras [in, out] layout."""
    import torch
    with torch.no_grad():
        sd = {}
        for k, v in hf_state.items():
            v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
            if k.startswith("embeddings."):
                sd["emb." + k[len("embeddings."):]] = v
            elif k.startswith("encoder.layer."):
                sd["layers." + k[len("encoder.layer."):]] = v
        missing, unexpected = net.load_state_dict(sd, strict=False)
        bad = [m for m in missing if not (m.startswith(("fc1", "fc2", "trans")) or m.endswith("position_ids") or m.endswith("token_type_ids"))]
        assert not bad and not unexpected, (bad, unexpected)
        net.fc1.weight.copy_(torch.from_numpy(np.ascontiguousarray(head["Wa"].T, dtype=np.float32)))
        net.fc1.bias.copy_(torch.from_numpy(np.asarray(head["ba"], np.float32)))
        net.fc2.weight.copy_(torch.from_numpy(np.ascontiguousarray(head["Wb"].T, dtype=np.float32)))
        net.fc2.bias.copy_(torch.from_numpy(np.asarray(head["bb"], np.float32)))
        net.trans.copy_(torch.from_numpy(np.asarray(head["trans"], np.float32)))
    return net


def emissions(net, ids, mask, tt):
    import torch
    add = (1.0 - mask.float())[:, None, None, :] * -10000.0
    h = net.emb(input_ids=ids, token_type_ids=tt)
    for l in net.layers:
        h = l(h, attention_mask=add)
        h = h[0] if isinstance(h, tuple) else h
    return net.fc2(torch.nn.functional.silu(net.fc1(net.drop(h))))


def named_grads(net, H):
    """Gradients under the names tests/parity.py uses for the device model (Keras [in, out] layout, fused q|k|v)."""
    g = lambda p: p.grad.detach().numpy().astype(np.float64)
    out = {"emb/word": g(net.emb.word_embeddings.weight), "emb/pos": g(net.emb.position_embeddings.weight),
           "emb/type": g(net.emb.token_type_embeddings.weight), "emb/emb_ln_g": g(net.emb.LayerNorm.weight),
           "emb/emb_ln_b": g(net.emb.LayerNorm.bias), "head/Wa": g(net.fc1.weight).T, "head/ba": g(net.fc1.bias),
           "head/Wb": g(net.fc2.weight).T, "head/bb": g(net.fc2.bias), "trans": g(net.trans)}
    for i, l in enumerate(net.layers):
        a = l.attention
        out[f"layers/{i}/Wqkv"] = np.concatenate([g(a.self.query.weight).T, g(a.self.key.weight).T, g(a.self.value.weight).T], axis=1)
        out[f"layers/{i}/bqkv"] = np.concatenate([g(a.self.query.bias), g(a.self.key.bias), g(a.self.value.bias)])
        out[f"layers/{i}/Wo"], out[f"layers/{i}/bo"] = g(a.output.dense.weight).T, g(a.output.dense.bias)
        out[f"layers/{i}/ln1_g"], out[f"layers/{i}/ln1_b"] = g(a.output.LayerNorm.weight), g(a.output.LayerNorm.bias)
        out[f"layers/{i}/W1"], out[f"layers/{i}/b1"] = g(l.intermediate.dense.weight).T, g(l.intermediate.dense.bias)
        out[f"layers/{i}/W2"], out[f"layers/{i}/b2"] = g(l.output.dense.weight).T, g(l.output.dense.bias)
        out[f"layers/{i}/ln2_g"], out[f"layers/{i}/ln2_b"] = g(l.output.LayerNorm.weight), g(l.output.LayerNorm.bias)
    return out


def keras_adam_steps(net, batch, steps, lr, beta1=0.9, beta2=0.999, eps=1e-7, bf16_compute_weights=False):
    """`steps` optimisation steps with Keras Adam (epsilon outside the bias-corrected sqrt, polus/training.py:191 with
    tf.keras.optimizers.Adam); returns the per-step losses.  torch.optim.Adam places epsilon differently.

    bf16_compute_weights: the storage format of the device path -- fp32 MASTER weights that Adam updates, and a
    bf16-rounded copy of them in every forward / backward (activations and all arithmetic stay fp32 here).  With it
    the comparison is about the arithmetic; without it, an update far below bf16's resolution (2^-9 relative) moves
    the fp32 reference but not yet the bf16 operand, and the two trajectories differ for that reason alone."""
    import torch
    ids, mask, tt, tags = batch
    params = [p for p in net.parameters() if p.requires_grad]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    master = [p.detach().clone() for p in params] if bf16_compute_weights else None
    losses = []
    for t in range(1, steps + 1):
        for p in params:
            p.grad = None
        if master is not None:
            with torch.no_grad():
                for p, w in zip(params, master):
                    p.copy_(w.to(torch.bfloat16).to(torch.float32))
        loss = net(ids, mask, tt, tags)
        loss.backward()
        losses.append(float(loss.detach()))
        lr_t = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
        with torch.no_grad():
            for i, (p, mi, vi) in enumerate(zip(params, m, v)):
                if p.grad is None:
                    continue
                mi.mul_(beta1).add_(p.grad, alpha=1.0 - beta1)
                vi.mul_(beta2).addcmul_(p.grad, p.grad, value=1.0 - beta2)
                (master[i] if master is not None else p).sub_(lr_t * mi / (vi.sqrt() + eps))
    return losses
