"""Pins the CPU oracle (oracle/numpy_ref.py) against independent implementations:
 * HF `transformers` torch BERT blocks + autograd, via the committed fixture tests/golden/bert_hf_torch.npz
   (generator: tests/golden/make_golden.py);
 * brute-force enumeration of all K^T paths for the CRF (the definition of tfa.text.crf_log_likelihood/crf_decode);
 * Random123 known-answer vectors for Philox4x32-10;
 * hand-evaluated closed forms for the LR schedule and Keras Adam."""
import itertools
import os

import numpy as np

from oracle import ner_model as O
from oracle import numpy_ref as R
from oracle import philox

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    z = np.load(os.path.join(HERE, "golden", "bert_hf_torch.npz"))
    tree = lambda tag: {"emb": {k: z[f"{tag}/emb/{k}"] for k in ("word", "pos", "type", "emb_ln_g", "emb_ln_b")},
                        "layers": [{k: z[f"{tag}/layers/{i}/{k}"] for k in
                                    ("Wqkv", "bqkv", "Wo", "bo", "ln1_g", "ln1_b", "W1", "b1", "W2", "b2", "ln2_g", "ln2_b")}
                                   for i in range(2)]}
    return z, tree("w"), tree("g")


def test_bert_forward_matches_hf_torch():
    z, w, _ = _golden()
    h, _ = R.bert_embeddings_fwd(z["ids"], z["tt"], w["emb"])
    np.testing.assert_allclose(h, z["h0"], rtol=1e-9, atol=1e-10)
    add = R.attention_mask_additive(z["mask"]).astype(np.float64)
    for i, lp in enumerate(w["layers"]):
        h, _ = R.bert_layer_fwd(h, add, lp, nh=4)
        np.testing.assert_allclose(h, z[f"h{i + 1}"], rtol=1e-8, atol=1e-9)


def test_bert_backward_matches_torch_autograd():
    z, w, g = _golden()
    h, ec = R.bert_embeddings_fwd(z["ids"], z["tt"], w["emb"])
    add = R.attention_mask_additive(z["mask"]).astype(np.float64)
    caches = []
    for lp in w["layers"]:
        h, c = R.bert_layer_fwd(h, add, lp, nh=4)
        caches.append(c)
    assert abs((h * z["R"]).sum() - z["loss"]) < 1e-8
    dh = z["R"]
    for i in (1, 0):
        dh, gl = R.bert_layer_bwd(dh, caches[i], w["layers"][i])
        for k, v in gl.items():
            np.testing.assert_allclose(v, g["layers"][i][k], rtol=1e-7, atol=1e-9, err_msg=f"layer {i} {k}")
    ge = R.bert_embeddings_bwd(dh, ec, w["emb"])
    for k, v in ge.items():
        np.testing.assert_allclose(v, g["emb"][k], rtol=1e-7, atol=1e-9, err_msg=k)


def test_crf_against_brute_force_enumeration():
    rng = np.random.default_rng(0)
    for T, K in [(1, 3), (4, 3), (5, 4)]:
        x = rng.standard_normal((2, T, K))
        A = rng.standard_normal((K, K))
        tags = rng.integers(0, K, (2, T))
        lens = np.array([T, max(T - 2, 1)])
        ll = R.crf_log_likelihood(x, tags, lens, A)
        dec, dscore = R.crf_decode(x, lens, A)
        for b in range(2):
            L = int(lens[b])
            scores = {p: R.crf_sequence_score(x[b], np.array(p), L, A) for p in itertools.product(range(K), repeat=L)}
            logZ = np.log(np.sum(np.exp(list(scores.values()))))
            gold = scores[tuple(tags[b, :L])]
            assert abs(ll[b] - (gold - logZ)) < 1e-10
            best = max(scores, key=lambda p: (scores[p], tuple(-q for q in p)))
            assert tuple(dec[b, :L]) == best and np.all(dec[b, L:] == 0)
            assert abs(dscore[b] - scores[best]) < 1e-10


def test_crf_gradients_by_finite_differences():
    rng = np.random.default_rng(1)
    B, T, K = 3, 6, 4
    x, A = rng.standard_normal((B, T, K)), rng.standard_normal((K, K))
    tags, lens, w = rng.integers(0, K, (B, T)), np.array([6, 3, 1]), np.array([1.0, 0.5, 2.0])
    _, loss, gx, gA = R.crf_nll_with_grads(x, tags, lens, A, w)
    f = lambda x_, A_: R.crf_nll_with_grads(x_, tags, lens, A_, w)[1]
    eps = 1e-6
    for _ in range(10):
        i = tuple(rng.integers(0, s) for s in x.shape)
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        assert abs((f(xp, A) - f(xm, A)) / (2 * eps) - gx[i]) < 1e-7
        j = tuple(rng.integers(0, K, 2))
        Ap, Am = A.copy(), A.copy()
        Ap[j] += eps
        Am[j] -= eps
        assert abs((f(x, Ap) - f(x, Am)) / (2 * eps) - gA[j]) < 1e-7


def test_crf_decode_tie_break_lowest_index():
    x = np.zeros((1, 4, 3), np.float32)
    tags, _ = R.crf_decode(x, np.array([4]), np.zeros((3, 3), np.float32))
    assert tags.tolist() == [[0, 0, 0, 0]]
    m = np.ones((4, 4), np.float32)
    m[1, 3] = 0  # the reference's only concrete mask example (tests/test_utils.py:77)
    t = R.crf_masked_transitions(np.full((4, 4), 0.5, np.float32), m)
    assert t[1, 3] == -10000.0 and t[0, 0] == 0.5


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    z = np.zeros(1, np.uint32)
    assert [int(v[0]) for v in philox.philox4x32_10(z, z, z, z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.full(1, 0xFFFFFFFF, np.uint32)
    assert [int(v[0]) for v in philox.philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    pi = [np.full(1, v, np.uint32) for v in (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]
    assert [int(v[0]) for v in philox.philox4x32_10(*pi, 0xa4093822, 0x299f31d0)] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    keep = philox.dropout_keep_mask(8 * 50000, 0.1, 42, 1, 0)
    assert abs(keep.mean() - 0.9) < 5e-3


def test_schedule_and_adam_closed_forms():
    # polus/schedulers.py: warm = int(100*0.1) = 10; decay over 90 steps to 1e-7
    assert R.warmup_schedule_lr(0, 100, 1e-3) == 0.0
    assert abs(R.warmup_schedule_lr(5, 100, 1e-3) - 5e-4) < 1e-12
    assert abs(R.warmup_schedule_lr(10, 100, 1e-3) - 1e-3) < 1e-12
    assert abs(R.warmup_schedule_lr(55, 100, 1e-3) - ((1e-3 - 1e-7) * 0.5 + 1e-7)) < 1e-12
    assert abs(R.warmup_schedule_lr(1000, 100, 1e-3) - 1e-7) < 1e-15
    # first Keras-Adam step: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1) => p - lr*g/(|g| + eps*sqrt(1-b2)... )
    p, m, v = R.adam_step(np.array([1.0]), np.array([0.5]), np.zeros(1), np.zeros(1), 1, 0.01)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(p[0] - (1.0 - lr_t * 0.05 / (np.sqrt(0.001 * 0.25) + 1e-7))) < 1e-12


def test_full_model_oracle_gradients_by_finite_differences():
    rng = np.random.default_rng(0)
    H, nh, I, V, S, B, K = 16, 2, 32, 50, 8, 2, 4
    lp = lambda: {"Wqkv": rng.normal(0, .2, (H, 3 * H)), "bqkv": rng.normal(0, .1, 3 * H), "Wo": rng.normal(0, .2, (H, H)),
                  "bo": rng.normal(0, .1, H), "ln1_g": 1 + rng.normal(0, .1, H), "ln1_b": rng.normal(0, .1, H),
                  "W1": rng.normal(0, .2, (H, I)), "b1": rng.normal(0, .1, I), "W2": rng.normal(0, .2, (I, H)),
                  "b2": rng.normal(0, .1, H), "ln2_g": 1 + rng.normal(0, .1, H), "ln2_b": rng.normal(0, .1, H)}
    params = {"emb": {"word": rng.normal(0, .5, (V, H)), "pos": rng.normal(0, .5, (S, H)), "type": rng.normal(0, .5, (2, H)),
                      "emb_ln_g": 1 + rng.normal(0, .1, H), "emb_ln_b": rng.normal(0, .1, H)},
              "layers": [lp()], "head": {"Wa": rng.normal(0, .3, (H, 8)), "ba": rng.normal(0, .1, 8),
                                         "Wb": rng.normal(0, .3, (8, K)), "bb": rng.normal(0, .1, K)},
              "trans": rng.normal(0, .3, (K, K))}
    ids, tt, tags = rng.integers(0, V, (B, S)), rng.integers(0, 2, (B, S)), rng.integers(0, K, (B, S))
    mask = np.ones((B, S), int)
    mask[1, 5:] = 0
    _, _, g = O.loss_and_grads(params, ids, mask, tt, tags, nh)
    fg = O.flatten(g)
    for key, arr in (("layers/0/Wqkv", params["layers"][0]["Wqkv"]), ("emb/pos", params["emb"]["pos"]),
                     ("head/Wa", params["head"]["Wa"]), ("trans", params["trans"])):
        idx = tuple(rng.integers(0, s) for s in arr.shape)
        old, eps = arr[idx], 1e-5
        arr[idx] = old + eps
        lp_ = O.loss_and_grads(params, ids, mask, tt, tags, nh)[0]
        arr[idx] = old - eps
        lm_ = O.loss_and_grads(params, ids, mask, tt, tags, nh)[0]
        arr[idx] = old
        assert abs((lp_ - lm_) / (2 * eps) - fg[key][idx]) < 1e-6 * max(1, abs(fg[key][idx])), key


def test_macro_f1_and_labels():
    cm = R.confusion_matrix(np.array([0, 0, 1, 1, 2]), np.array([0, 1, 1, 1, 0]), 3)
    assert cm.tolist() == [[1, 1, 0], [0, 2, 0], [1, 0, 0]]
    # class 0: p=1/2 r=1/2 -> .5 ; class 1: p=1, r=2/3 -> .8 ; class 2: 0
    assert abs(R.macro_f1(cm) - (0.5 + 0.8 + 0.0) / 3) < 1e-12
    assert R.TAG2INT == {"PAD": 0, "O": 1, "B-Chemical": 2, "I-Chemical": 3}  # polus/ner/utils.py:9-15
