"""CPU restatement (numpy) of the math on polus's data-parallel training step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (polus_b200/) imports this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and
only as the checker / the CPU baseline.

Why a restatement: the reference (/root/reference, bioinformatics-ua/polus v0.2.1) is pure Python on
TensorFlow + HuggingFace transformers (TF classes) + tensorflow_addons + Horovod, none of which is
installed here or on the GPU box -- `import polus` fails at polus/__init__.py:93.
The arithmetic lives in those un-vendored, un-pinned dependencies (requirements.txt:1-9), so each
function below restates the published algorithm of the dependency and cites the reference call site
that reaches it.

PARITY PINNING.  The reference's own tests hold no golden vectors for this path (the
only numeric check, tests/utils.py:3-5, is a one-sided cosine test that cannot fail).  This oracle is
therefore pinned against independent implementations that ARE importable in the build container:
  * BERT layer / embeddings forward+backward: HuggingFace `transformers` torch `BertModel` (eager
    attention, explicit additive -10000 mask) + torch autograd  -> tests/golden/bert_*.npz
    (generator: tests/golden/make_golden.py), checked in tests/test_oracle_cpu.py.
  * CRF log-likelihood / Viterbi: brute-force enumeration over all K^T paths (exact definition of
    tfa.text.crf_log_likelihood / crf_decode) for small T, and torch autograd for the gradient.
  * Keras Adam / HF WarmUp+PolynomialDecay: closed-form restatement checked against hand-computed
    steps; no TF available => "parity unpinned" against TF itself for these two.
Anything not pinned by one of the above is marked `# unpinned` at its definition.
"""
import math

import numpy as np

# --------------------------------------------------------------------------------------------------
# activations
# --------------------------------------------------------------------------------------------------
_erf = np.vectorize(math.erf, otypes=[np.float64])


def gelu(x):
    """Exact erf GELU -- HF BertIntermediate with hidden_act="gelu" (reached via polus/models.py:205-213)."""
    return (0.5 * x * (1.0 + _erf(np.asarray(x, np.float64) / math.sqrt(2.0)))).astype(x.dtype)


def gelu_grad(x):
    x64 = np.asarray(x, np.float64)
    cdf = 0.5 * (1.0 + _erf(x64 / math.sqrt(2.0)))
    pdf = np.exp(-0.5 * x64 * x64) / math.sqrt(2.0 * math.pi)
    return (cdf + x64 * pdf).astype(x.dtype)


def swish(x):
    """tf.keras.activations.swish = x*sigmoid(x) (polus/ner/models.py:30)."""
    return x / (1.0 + np.exp(-x))


def swish_grad(x):
    s = 1.0 / (1.0 + np.exp(-x))
    return s * (1.0 + x * (1.0 - s))


def mish(x):
    """tfa.activations.mish = x*tanh(softplus(x)) (polus/models.py:53-57)."""
    return x * np.tanh(np.log1p(np.exp(x)))


def mish_grad(x):
    sp = np.log1p(np.exp(x))
    t = np.tanh(sp)
    s = 1.0 / (1.0 + np.exp(-x))
    return t + x * (1.0 - t * t) * s


def relu(x):
    return np.maximum(x, 0)


ACT = {None: (lambda x: x, lambda x: np.ones_like(x)), "linear": (lambda x: x, lambda x: np.ones_like(x)),
       "gelu": (gelu, gelu_grad), "swish": (swish, swish_grad), "mish": (mish, mish_grad),
       "relu": (relu, lambda x: (x > 0).astype(x.dtype)),
       "tanh": (np.tanh, lambda x: 1.0 - np.tanh(x) ** 2)}


# --------------------------------------------------------------------------------------------------
# LayerNorm / softmax / dense
# --------------------------------------------------------------------------------------------------
def layer_norm(x, gamma, beta, eps=1e-12):
    """tf.keras LayerNormalization, biased variance, eps 1e-12 (HF BertConfig.layer_norm_eps)."""
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def layer_norm_bwd(dy, cache, gamma):
    xhat, rstd = cache
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(-1, keepdims=True)


def attention_mask_additive(mask):
    """polus/models.py:175-195: (1 - mask) * -10000.0, broadcast as [B,1,1,S]."""
    return (1.0 - mask.astype(np.float32))[:, None, None, :] * -10000.0


# --------------------------------------------------------------------------------------------------
# BERT encoder layer (HF TFBertLayer, post-LN) forward / backward.  Dropout masks are optional
# inputs (pre-scaled keep masks: 0 or 1/(1-p)) so device runs with identical Philox masks can be checked.
# params: dict with Wqkv [H,3H] (q|k|v column blocks), bqkv [3H], Wo [H,H], bo, ln1_g, ln1_b, W1 [H,I], b1,
#         W2 [I,H], b2, ln2_g, ln2_b   (Keras kernel layout [in,out])
# --------------------------------------------------------------------------------------------------
def bert_layer_fwd(x, add_mask, p, nh, masks=None, eps=1e-12):
    B, S, H = x.shape
    dh = H // nh
    masks = masks or {}
    qkv = x @ p["Wqkv"] + p["bqkv"]
    q, k, v = [qkv[..., i * H:(i + 1) * H].reshape(B, S, nh, dh).transpose(0, 2, 1, 3) for i in range(3)]
    scores = q @ k.transpose(0, 1, 3, 2) / math.sqrt(dh)
    if add_mask is not None:
        scores = scores + add_mask
    P = softmax(scores)
    Pd = P * masks["attn"] if "attn" in masks else P
    ctx = (Pd @ v).transpose(0, 2, 1, 3).reshape(B, S, H)
    ao = ctx @ p["Wo"] + p["bo"]
    ao_d = ao * masks["hidden1"] if "hidden1" in masks else ao
    z1 = ao_d + x
    h1, ln1c = layer_norm(z1, p["ln1_g"], p["ln1_b"], eps)
    u = h1 @ p["W1"] + p["b1"]
    a = gelu(u)
    o = a @ p["W2"] + p["b2"]
    o_d = o * masks["hidden2"] if "hidden2" in masks else o
    z2 = o_d + h1
    y, ln2c = layer_norm(z2, p["ln2_g"], p["ln2_b"], eps)
    cache = dict(x=x, q=q, k=k, v=v, P=P, Pd=Pd, ctx=ctx, h1=h1, ln1c=ln1c, u=u, a=a, ln2c=ln2c, masks=masks, nh=nh)
    return y, cache


def bert_layer_bwd(dy, c, p):
    x, q, k, v, P, Pd = c["x"], c["q"], c["k"], c["v"], c["P"], c["Pd"]
    masks, nh = c["masks"], c["nh"]
    B, S, H = x.shape
    dh = H // nh
    g = {}
    dz2, g["ln2_g"], g["ln2_b"] = layer_norm_bwd(dy, c["ln2c"], p["ln2_g"])
    do = dz2 * masks["hidden2"] if "hidden2" in masks else dz2
    dh1 = dz2.copy()
    g["W2"] = c["a"].reshape(-1, c["a"].shape[-1]).T @ do.reshape(-1, H)
    g["b2"] = do.reshape(-1, H).sum(0)
    da = do @ p["W2"].T
    du = da * gelu_grad(c["u"])
    g["W1"] = c["h1"].reshape(-1, H).T @ du.reshape(-1, du.shape[-1])
    g["b1"] = du.reshape(-1, du.shape[-1]).sum(0)
    dh1 += du @ p["W1"].T
    dz1, g["ln1_g"], g["ln1_b"] = layer_norm_bwd(dh1, c["ln1c"], p["ln1_g"])
    dao = dz1 * masks["hidden1"] if "hidden1" in masks else dz1
    dx = dz1.copy()
    g["Wo"] = c["ctx"].reshape(-1, H).T @ dao.reshape(-1, H)
    g["bo"] = dao.reshape(-1, H).sum(0)
    dctx = (dao @ p["Wo"].T).reshape(B, S, nh, dh).transpose(0, 2, 1, 3)
    dPd = dctx @ v.transpose(0, 1, 3, 2)
    dv = Pd.transpose(0, 1, 3, 2) @ dctx
    dP = dPd * masks["attn"] if "attn" in masks else dPd
    dS = P * (dP - (P * dP).sum(-1, keepdims=True)) / math.sqrt(dh)
    dq = dS @ k
    dk = dS.transpose(0, 1, 3, 2) @ q
    dqkv = np.concatenate([t.transpose(0, 2, 1, 3).reshape(B, S, H) for t in (dq, dk, dv)], -1)
    g["Wqkv"] = x.reshape(-1, H).T @ dqkv.reshape(-1, 3 * H)
    g["bqkv"] = dqkv.reshape(-1, 3 * H).sum(0)
    dx += dqkv @ p["Wqkv"].T
    return dx, g


def bert_embeddings_fwd(ids, tt, p, drop_mask=None, eps=1e-12):
    """HF TFBertEmbeddings: LN(word[ids] + pos[0..S) + type[tt]) then dropout (polus/data.py:526-545)."""
    B, S = ids.shape
    z = p["word"][ids] + p["pos"][None, :S] + p["type"][tt]
    y, c = layer_norm(z, p["emb_ln_g"], p["emb_ln_b"], eps)
    if drop_mask is not None:
        y = y * drop_mask
    return y, dict(ids=ids, tt=tt, ln=c, drop_mask=drop_mask)


def bert_embeddings_bwd(dy, c, p):
    if c["drop_mask"] is not None:
        dy = dy * c["drop_mask"]
    dz, gg, gb = layer_norm_bwd(dy, c["ln"], p["emb_ln_g"])
    B, S, H = dz.shape
    g = {"emb_ln_g": gg, "emb_ln_b": gb, "word": np.zeros_like(p["word"]), "pos": np.zeros_like(p["pos"]),
         "type": np.zeros_like(p["type"])}
    np.add.at(g["word"], c["ids"].reshape(-1), dz.reshape(-1, H))
    g["pos"][:S] += dz.sum(0)
    np.add.at(g["type"], c["tt"].reshape(-1), dz.reshape(-1, H))
    return g


# --------------------------------------------------------------------------------------------------
# NER head: [Dropout] -> Dense(H->hidden, act) -> Dense(hidden->K)  (polus/ner/models.py:26-67)
# --------------------------------------------------------------------------------------------------
def ner_head_fwd(h, p, activation="swish", drop_mask=None):
    f, _ = ACT[activation]
    hd = h * drop_mask if drop_mask is not None else h
    u = hd @ p["Wa"] + p["ba"]
    a = f(u)
    e = a @ p["Wb"] + p["bb"]
    return e, dict(hd=hd, u=u, a=a, activation=activation, drop_mask=drop_mask)


def ner_head_bwd(de, c, p):
    _, fg = ACT[c["activation"]]
    H = c["hd"].shape[-1]
    g = {"Wb": c["a"].reshape(-1, c["a"].shape[-1]).T @ de.reshape(-1, de.shape[-1]), "bb": de.reshape(-1, de.shape[-1]).sum(0)}
    da = de @ p["Wb"].T
    du = da * fg(c["u"])
    g["Wa"] = c["hd"].reshape(-1, H).T @ du.reshape(-1, du.shape[-1])
    g["ba"] = du.reshape(-1, du.shape[-1]).sum(0)
    dh = du @ p["Wa"].T
    if c["drop_mask"] is not None:
        dh = dh * c["drop_mask"]
    return dh, g


# --------------------------------------------------------------------------------------------------
# CRF (tensorflow_addons.text) -- polus/layers.py:56-126
# --------------------------------------------------------------------------------------------------
def crf_masked_transitions(trans, mask):
    """CRF.get_transitions (polus/layers.py:56-63)."""
    if mask is None:
        return trans
    return trans * mask + ((1 - mask).astype(np.int32) * -10000).astype(np.float32)


def _lse(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def crf_log_norm(x, length, trans):
    if length <= 0:
        return x.dtype.type(0)
    alpha = x[0]
    for t in range(1, length):
        alpha = x[t] + _lse(alpha[:, None] + trans, 0)
    return _lse(alpha, 0)


def crf_sequence_score(x, tags, length, trans):
    s = x.dtype.type(0)
    for t in range(length):
        s += x[t, tags[t]]
        if t + 1 < length:
            s += trans[tags[t], tags[t + 1]]
    return s


def crf_log_likelihood(emis, tags, lens, trans):
    """tfa.text.crf_log_likelihood -> [B] (called at polus/layers.py:91-96,109-114)."""
    return np.array([crf_sequence_score(emis[b], tags[b], int(lens[b]), trans) - crf_log_norm(emis[b], int(lens[b]), trans)
                     for b in range(emis.shape[0])], dtype=emis.dtype)


def crf_nll_with_grads(emis, tags, lens, trans, weights=None):
    """loss = mean_b(w_b * -ll_b) and its gradients (forward-backward).  polus/layers.py:86-126."""
    B, T, K = emis.shape
    w = np.ones(B, emis.dtype) if weights is None else np.asarray(weights, emis.dtype)
    nll = np.zeros(B, emis.dtype)
    gem = np.zeros_like(emis)
    gtr = np.zeros_like(trans)
    for b in range(B):
        L = int(lens[b])
        if L <= 0:
            continue
        x = emis[b]
        alpha = np.zeros((L, K), emis.dtype)
        alpha[0] = x[0]
        for t in range(1, L):
            alpha[t] = x[t] + _lse(alpha[t - 1][:, None] + trans, 0)
        logZ = _lse(alpha[L - 1], 0)
        beta = np.zeros((L, K), emis.dtype)
        for t in range(L - 2, -1, -1):
            beta[t] = _lse(trans + (x[t + 1] + beta[t + 1])[None, :], 1)
        nll[b] = logZ - crf_sequence_score(x, tags[b], L, trans)
        s = w[b] / B
        marg = np.exp(alpha + beta - logZ)
        marg[np.arange(L), tags[b, :L]] -= 1.0
        gem[b, :L] = s * marg
        for t in range(1, L):
            pair = np.exp(alpha[t - 1][:, None] + trans + (x[t] + beta[t])[None, :] - logZ)
            pair[tags[b, t - 1], tags[b, t]] -= 1.0
            gtr += s * pair
    return nll, (w * nll).mean(), gem, gtr


def crf_decode(emis, lens, trans):
    """tfa.text.crf_decode (polus/layers.py:78-80): Viterbi; ties -> lowest index; t >= len -> 0."""
    B, T, K = emis.shape
    tags = np.zeros((B, T), np.int32)
    score = np.zeros(B, emis.dtype)
    for b in range(B):
        L = int(lens[b])
        if L <= 0:
            continue
        delta = emis[b, 0].copy()
        bp = np.zeros((L, K), np.int64)
        for t in range(1, L):
            cand = delta[:, None] + trans  # [i, j]
            bp[t] = cand.argmax(0)
            delta = emis[b, t] + cand.max(0)
        cur = int(delta.argmax())
        score[b] = delta[cur]
        tags[b, L - 1] = cur
        for t in range(L - 1, 0, -1):
            cur = int(bp[t, cur])
            tags[b, t - 1] = cur
    return tags, score


def crf_sample_weights(y_true_onehot, mask_positive_classes, negative_weight):
    """polus/layers.py:116-121."""
    pos = y_true_onehot * mask_positive_classes
    neg = np.all(pos == 0, axis=(-2, -1)).astype(np.float32) * negative_weight
    return np.any(pos == 1, axis=(-2, -1)).astype(np.float32) + neg


# --------------------------------------------------------------------------------------------------
# losses  (polus/losses.py, tutorials/classifier_example.py:55)
# --------------------------------------------------------------------------------------------------
def sparse_softmax_xent(logits, labels):
    lse = _lse(logits, -1)
    rows = logits.shape[0]
    loss = (lse - logits[np.arange(rows), labels]).mean()
    g = np.exp(logits - lse[:, None])
    g[np.arange(rows), labels] -= 1.0
    return loss, g / rows


def weighted_softmax_xent(logits, y, class_weights):
    """polus/losses.py:5-18."""
    w = (np.asarray(class_weights, np.float32) * y).sum(-1)
    lse = _lse(logits, -1)
    ce = (y * (lse[:, None] - logits)).sum(-1)
    rows = logits.shape[0]
    g = (np.exp(logits - lse[:, None]) * y.sum(-1, keepdims=True) - y) * w[:, None] / rows
    return (ce * w).mean(), g


def weighted_sigmoid_xent(logits, y, class_weights, negative_weight):
    """polus/losses.py:21-42."""
    w = (np.asarray(class_weights, np.float32) * y).sum(-1) + np.all(y == 0, -1).astype(np.float32) * negative_weight
    l = (np.maximum(logits, 0) - logits * y + np.log1p(np.exp(-np.abs(logits)))).sum(-1)
    rows = logits.shape[0]
    g = (1.0 / (1.0 + np.exp(-logits)) - y) * w[:, None] / rows
    return (l * w).mean(), g


# --------------------------------------------------------------------------------------------------
# optimizer + schedule  (polus/training.py:191, polus/schedulers.py:5-23)  # unpinned vs TF
# --------------------------------------------------------------------------------------------------
def warmup_schedule_lr(step, num_train_steps, max_lr, warmup_percentage=0.1):
    """HF WarmUp(PolynomialDecay(power=1, end=1e-7)); note polus ignores its end_lr arg (schedulers.py:15)."""
    warm = int(num_train_steps * warmup_percentage)
    decay_steps = num_train_steps - warm
    if step < warm:
        return max_lr * (step / warm)
    d = min(step - warm, decay_steps)
    return (max_lr - 1e-7) * (1.0 - d / decay_steps) + 1e-7


def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7, weight_decay=0.0, decay=True):
    """Keras Adam (t = iterations+1): epsilon outside the bias-corrected sqrt.  weight_decay>0: HF AdamWeightDecay."""
    if weight_decay > 0 and decay:
        p = p - lr * weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


# --------------------------------------------------------------------------------------------------
# metrics / labels  (polus/metrics.py:55-91, polus/ner/utils.py:9-15)
# --------------------------------------------------------------------------------------------------
TAG2INT = {"PAD": 0, "O": 1, "B-Chemical": 2, "I-Chemical": 3}


def confusion_matrix(y_true, y_pred, num_classes):
    cm = np.zeros((num_classes, num_classes), np.int32)
    np.add.at(cm, (y_true, y_pred), 1)
    return cm


def macro_f1(cm):
    """polus/metrics.py:74-91 (float64, divide_no_nan); note its 'precision' divides by row sums."""
    tp = np.diag(cm).astype(np.float64)
    fp_tp = cm.sum(-1).astype(np.float64)
    fn_tp = cm.sum(-2).astype(np.float64)

    def dnn(a, b):
        b = np.asarray(b, np.float64)
        out = np.zeros_like(b)
        nz = b != 0
        out[nz] = (np.broadcast_to(a, b.shape)[nz]) / b[nz]
        return out

    precision, recall = dnn(tp, fp_tp), dnn(tp, fn_tp)
    inv_p, inv_r = dnn(1.0, precision), dnn(1.0, recall)
    return float(dnn(2.0, inv_p + inv_r).mean())
