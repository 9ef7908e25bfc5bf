"""Generates tests/golden/bert_hf_torch.npz: inputs, weights, outputs and gradients of HuggingFace
`transformers` (torch) BERT blocks, the independent implementation that pins oracle/numpy_ref.py.

Run in the build container (needs torch + transformers, CPU only):  python tests/golden/make_golden.py
The reference's own model code (polus/models.py:157-216) delegates to the TF twins of exactly these
classes (TFBertModel / TFBertLayer); TF is not installable here, the torch classes are the
same architecture from the same library.  Mask constant is polus's -10000 (polus/models.py:190-193).
"""
import os

import numpy as np
import torch
from transformers import BertConfig
from transformers.models.bert.modeling_bert import BertEmbeddings, BertLayer

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(True)
    cfg = BertConfig(vocab_size=100, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                     max_position_embeddings=32, type_vocab_size=2, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0, layer_norm_eps=1e-12)
    cfg._attn_implementation = "eager"
    emb = BertEmbeddings(cfg).double()
    layers = [BertLayer(cfg).double() for _ in range(2)]
    for m in [emb] + layers:
        for name, p in m.named_parameters():
            with torch.no_grad():
                if "LayerNorm.weight" in name:
                    p.copy_(1 + 0.1 * torch.randn_like(p))
                elif name.endswith("bias"):
                    p.copy_(0.05 * torch.randn_like(p))
                else:
                    p.copy_(0.1 * torch.randn_like(p))
    B, S = 2, 16
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 100, (B, S), generator=g)
    tt = torch.randint(0, 2, (B, S), generator=g)
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 11:] = 0
    add_mask = (1.0 - mask.double())[:, None, None, :] * -10000.0
    h = emb(input_ids=ids, token_type_ids=tt)
    hs = [h]
    for l in layers:
        h = l(h, attention_mask=add_mask)
        h = h[0] if isinstance(h, tuple) else h
        hs.append(h)
    R = torch.randn(h.shape, generator=g, dtype=torch.double)
    loss = (h * R).sum()
    loss.backward()
    out = {"ids": ids.numpy(), "tt": tt.numpy(), "mask": mask.numpy(), "R": R.numpy(), "loss": np.float64(loss.item())}
    for i, t in enumerate(hs):
        out[f"h{i}"] = t.detach().numpy()
    n = lambda p: p.detach().numpy()
    e = emb
    for tag, f in (("w", lambda p: n(p)), ("g", lambda p: n(p.grad))):
        out[f"{tag}/emb/word"] = f(e.word_embeddings.weight)
        out[f"{tag}/emb/pos"] = f(e.position_embeddings.weight)
        out[f"{tag}/emb/type"] = f(e.token_type_embeddings.weight)
        out[f"{tag}/emb/emb_ln_g"] = f(e.LayerNorm.weight)
        out[f"{tag}/emb/emb_ln_b"] = f(e.LayerNorm.bias)
        for i, l in enumerate(layers):
            a, o = l.attention.self, l.attention.output
            # torch Linear stores [out,in]; Keras kernels are [in,out]
            out[f"{tag}/layers/{i}/Wqkv"] = np.concatenate([f(a.query.weight).T, f(a.key.weight).T, f(a.value.weight).T], 1)
            out[f"{tag}/layers/{i}/bqkv"] = np.concatenate([f(a.query.bias), f(a.key.bias), f(a.value.bias)])
            out[f"{tag}/layers/{i}/Wo"] = f(o.dense.weight).T
            out[f"{tag}/layers/{i}/bo"] = f(o.dense.bias)
            out[f"{tag}/layers/{i}/ln1_g"] = f(o.LayerNorm.weight)
            out[f"{tag}/layers/{i}/ln1_b"] = f(o.LayerNorm.bias)
            out[f"{tag}/layers/{i}/W1"] = f(l.intermediate.dense.weight).T
            out[f"{tag}/layers/{i}/b1"] = f(l.intermediate.dense.bias)
            out[f"{tag}/layers/{i}/W2"] = f(l.output.dense.weight).T
            out[f"{tag}/layers/{i}/b2"] = f(l.output.dense.bias)
            out[f"{tag}/layers/{i}/ln2_g"] = f(l.output.LayerNorm.weight)
            out[f"{tag}/layers/{i}/ln2_b"] = f(l.output.LayerNorm.bias)
    np.savez_compressed(os.path.join(HERE, "bert_hf_torch.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("wrote bert_hf_torch.npz:", len(out), "arrays, loss", loss.item())


if __name__ == "__main__":
    main()
