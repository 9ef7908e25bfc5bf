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
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, threads
