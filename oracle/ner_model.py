"""CPU restatement of one full training step of the assembled BERT-NER+CRF model -- TEST INFRASTRUCTURE ONLY.

Follows the step order of the reference's BaseTrainer.train_step (polus/training.py:173-193): forward
(HF TFBertEmbeddings -> TFBertLayer x L, polus/models.py:205-213; head polus/ner/models.py:26-67; CRF loss
polus/layers.py:86-99 with sequence_lengths = full length, layers.py:74-76), backward, Keras Adam.
Built from oracle/numpy_ref.py; see that module's header for how parity is pinned.
"""
import numpy as np

from . import numpy_ref as R


def forward(params, ids, mask, tt, nh, masks=None, activation="swish", dtype=np.float64):
    masks = masks or {}
    cast = lambda d: {k: (np.asarray(v, dtype) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    emb = cast(params["emb"])
    h, ec = R.bert_embeddings_fwd(ids, tt, emb, masks.get("emb"))
    add_mask = R.attention_mask_additive(mask).astype(dtype) if mask is not None else None
    caches = []
    for li, lp in enumerate(params["layers"]):
        h, c = R.bert_layer_fwd(h, add_mask, cast(lp), nh, masks.get(("layer", li)))
        caches.append(c)
    e, hc = R.ner_head_fwd(h, cast(params["head"]), activation, masks.get("head"))
    return e, dict(emb=ec, layers=caches, head=hc, hidden=h)


def loss_and_grads(params, ids, mask, tt, tags, nh, masks=None, activation="swish", lens=None, dtype=np.float64):
    e, c = forward(params, ids, mask, tt, nh, masks, activation, dtype)
    B, S, K = e.shape
    lens = np.full(B, S, np.int64) if lens is None else lens
    trans = np.asarray(params["trans"], dtype)
    nll, loss, ge, gtrans = R.crf_nll_with_grads(e, tags, lens, trans)
    cast = lambda d: {k: np.asarray(v, dtype) for k, v in d.items()}
    dh, ghead = R.ner_head_bwd(ge, c["head"], cast(params["head"]))
    glayers = []
    for lp, lc in zip(reversed(params["layers"]), reversed(c["layers"])):
        dh, g = R.bert_layer_bwd(dh, lc, cast(lp))
        glayers.append(g)
    glayers.reverse()
    gemb = R.bert_embeddings_bwd(dh, c["emb"], cast(params["emb"]))
    return loss, e, dict(emb=gemb, layers=glayers, head=ghead, trans=gtrans)


def flatten(tree, prefix=""):
    out = {}
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.update(flatten(v, f"{prefix}{k}/"))
    elif isinstance(tree, list):
        for i, v in enumerate(tree):
            out.update(flatten(v, f"{prefix}{i}/"))
    else:
        out[prefix[:-1]] = tree
    return out


def adam_train(params, batches, nh, lr=1e-3, steps=1, activation="swish", dtype=np.float64):
    """`steps` optimisation steps (Keras Adam) cycling through `batches`; returns losses and final params."""
    import copy
    p = copy.deepcopy(params)
    flat_m = {k: np.zeros_like(np.asarray(v, dtype)) for k, v in flatten(p).items()}
    flat_v = {k: np.zeros_like(np.asarray(v, dtype)) for k, v in flatten(p).items()}
    losses = []
    for t in range(1, steps + 1):
        ids, mask, tt, tags = batches[(t - 1) % len(batches)]
        loss, _, g = loss_and_grads(p, ids, mask, tt, tags, nh, activation=activation, dtype=dtype)
        losses.append(float(loss))
        fg = flatten(g)

        def upd(tree, prefix=""):
            if isinstance(tree, dict):
                for k in tree:
                    if isinstance(tree[k], (dict, list)):
                        upd(tree[k], f"{prefix}{k}/")
                    else:
                        key = f"{prefix}{k}"
                        newp, flat_m[key], flat_v[key] = R.adam_step(np.asarray(tree[k], dtype), fg[key], flat_m[key], flat_v[key], t, lr)
                        tree[k] = newp
            else:
                for i, v in enumerate(tree):
                    upd(v, f"{prefix}{i}/")
        upd(p)
    return losses, p
