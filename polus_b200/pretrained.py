"""Import of HuggingFace BERT checkpoints into polus_b200.nn.BertModel.

The reference gets pretrained encoders from `TFAutoModel.from_pretrained(checkpoint)` (polus/data.py:526-530,
polus/models.py:225-229).  Here the weights come from a HuggingFace *state dict* -- a `.safetensors` file, a `.npz`, or a
mapping name -> array -- with the torch naming (`bert.` prefix optional):

    embeddings.word_embeddings.weight            [V, H]          encoder.layer.N.attention.output.dense.weight  [H, H]
    embeddings.position_embeddings.weight        [P, H]          encoder.layer.N.attention.output.LayerNorm.*
    embeddings.token_type_embeddings.weight      [T, H]          encoder.layer.N.intermediate.dense.weight      [I, H]
    embeddings.LayerNorm.{weight,bias}                           encoder.layer.N.output.dense.weight            [H, I]
    encoder.layer.N.attention.self.{query,key,value}.{weight [H, H], bias}    encoder.layer.N.output.LayerNorm.*
    pooler.dense.{weight,bias}

torch `Linear.weight` is [out, in]; the kernels here keep the Keras layout [in, out] and fuse q|k|v into one [H, 3H]
variable, so every dense weight is transposed and the three projections are concatenated.
"""
import numpy as np


def _load_state(source):
    if isinstance(source, dict):
        return {k: np.asarray(v) for k, v in source.items()}
    path = str(source)
    if path.endswith(".safetensors"):
        from safetensors.numpy import load_file
        return load_file(path)
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    raise ValueError(f"unsupported checkpoint {path!r}: expected .safetensors, .npz or a dict")


def load_hf_bert_weights(model, source, strict=True):
    """Copy a HuggingFace BERT state dict into `model` (a polus_b200.nn.BertModel or its BertMainLayer).
    Returns the list of state-dict keys that were not used (heads such as `cls.*`)."""
    state = _load_state(source)
    state = {(k[5:] if k.startswith("bert.") else k): v for k, v in state.items()}
    used = set()

    def get(name):
        if name not in state:
            raise KeyError(f"checkpoint has no tensor {name!r}")
        used.add(name)
        return np.asarray(state[name], dtype=np.float32)

    main = getattr(model, "bert", model)
    cfg = main.config
    emb = main.embeddings
    for param, name in ((emb.word, "embeddings.word_embeddings.weight"), (emb.position, "embeddings.position_embeddings.weight"),
                        (emb.token_type, "embeddings.token_type_embeddings.weight"), (emb.ln_gamma, "embeddings.LayerNorm.weight"),
                        (emb.ln_beta, "embeddings.LayerNorm.bias")):
        value = get(name)
        if value.shape != tuple(param.shape):
            raise ValueError(f"{name}: checkpoint shape {value.shape} != model shape {tuple(param.shape)} "
                             f"(vocab/positions/hidden of the BertConfig must match the checkpoint)")
        param.assign(value)
    for i, layer in enumerate(main.encoder.layer):
        p = f"encoder.layer.{i}."
        qkv_w = np.concatenate([get(p + f"attention.self.{n}.weight").T for n in ("query", "key", "value")], axis=1)
        qkv_b = np.concatenate([get(p + f"attention.self.{n}.bias") for n in ("query", "key", "value")])
        pairs = ((layer.Wqkv, qkv_w), (layer.bqkv, qkv_b),
                 (layer.Wo, get(p + "attention.output.dense.weight").T), (layer.bo, get(p + "attention.output.dense.bias")),
                 (layer.ln1_g, get(p + "attention.output.LayerNorm.weight")), (layer.ln1_b, get(p + "attention.output.LayerNorm.bias")),
                 (layer.W1, get(p + "intermediate.dense.weight").T), (layer.b1, get(p + "intermediate.dense.bias")),
                 (layer.W2, get(p + "output.dense.weight").T), (layer.b2, get(p + "output.dense.bias")),
                 (layer.ln2_g, get(p + "output.LayerNorm.weight")), (layer.ln2_b, get(p + "output.LayerNorm.bias")))
        for param, value in pairs:
            if value.shape != tuple(param.shape):
                raise ValueError(f"layer {i}: checkpoint shape {value.shape} != model shape {tuple(param.shape)}")
            param.assign(np.ascontiguousarray(value))
    if main.pooler is not None and "pooler.dense.weight" in state:
        if main.pooler.dense.kernel is None:
            main.pooler.dense.build((None, main.config.hidden_size))
            main.pooler.dense.built = True
        main.pooler.dense.kernel.assign(np.ascontiguousarray(get("pooler.dense.weight").T))
        main.pooler.dense.bias.assign(get("pooler.dense.bias"))
    unused = sorted(k for k in state if k not in used and not k.endswith("position_ids"))
    if strict and any(k.startswith(("embeddings.", "encoder.")) for k in unused):
        raise ValueError(f"checkpoint tensors without a destination (more layers than the model?): {unused[:5]}")
    del cfg
    return unused


def export_hf_bert_weights(model):
    """The inverse mapping: a HuggingFace-named state dict (numpy) of `model` -- what `safetensors.numpy.save_file` or
    `torch.load_state_dict` expect."""
    main = getattr(model, "bert", model)
    emb = main.embeddings
    H = main.config.hidden_size
    out = {"embeddings.word_embeddings.weight": emb.word.numpy(), "embeddings.position_embeddings.weight": emb.position.numpy(),
           "embeddings.token_type_embeddings.weight": emb.token_type.numpy(), "embeddings.LayerNorm.weight": emb.ln_gamma.numpy(),
           "embeddings.LayerNorm.bias": emb.ln_beta.numpy()}
    for i, layer in enumerate(main.encoder.layer):
        p = f"encoder.layer.{i}."
        w, b = layer.Wqkv.numpy(), layer.bqkv.numpy()
        for j, n in enumerate(("query", "key", "value")):
            out[p + f"attention.self.{n}.weight"] = np.ascontiguousarray(w[:, j * H:(j + 1) * H].T)
            out[p + f"attention.self.{n}.bias"] = b[j * H:(j + 1) * H].copy()
        out[p + "attention.output.dense.weight"] = np.ascontiguousarray(layer.Wo.numpy().T)
        out[p + "attention.output.dense.bias"] = layer.bo.numpy()
        out[p + "attention.output.LayerNorm.weight"] = layer.ln1_g.numpy()
        out[p + "attention.output.LayerNorm.bias"] = layer.ln1_b.numpy()
        out[p + "intermediate.dense.weight"] = np.ascontiguousarray(layer.W1.numpy().T)
        out[p + "intermediate.dense.bias"] = layer.b1.numpy()
        out[p + "output.dense.weight"] = np.ascontiguousarray(layer.W2.numpy().T)
        out[p + "output.dense.bias"] = layer.b2.numpy()
        out[p + "output.LayerNorm.weight"] = layer.ln2_g.numpy()
        out[p + "output.LayerNorm.bias"] = layer.ln2_b.numpy()
    if main.pooler is not None:
        out["pooler.dense.weight"] = np.ascontiguousarray(main.pooler.dense.kernel.numpy().T)
        out["pooler.dense.bias"] = main.pooler.dense.bias.numpy()
    return out
