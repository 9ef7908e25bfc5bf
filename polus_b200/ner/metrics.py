"""Validation metrics of polus.ner (reference polus/ner/metrics.py:8-72, polus/ner/utils.py:231-308).

* ISequentialConfusionMatrixTF / MacroF1Score / Accuracy: token-level metrics over [B, S] label / prediction tensors;
  the confusion matrix is built on the device (polus_confusion_matrix), the K x K arithmetic stays on the host.
* EntityF1: strict entity-level precision / recall / F1 (span and type must match exactly,
  polus/ner/utils.py:269-308).  The reference decodes through its BioC corpus objects (polus/ner/elements.py, out of
  scope here); this class takes the same per-sample fields its decoder reads -- `spans`,
  `tags_int_pred`, `is_prediction` plus the gold `tags_int` (or a gold entity list) keyed by a document identifier --
  and applies the same BIO decoding (polus_b200.ner.bio.decode_bio).
"""
import numpy as np

from ..metrics import IConfusionMatrixTF, IMetric, _divide_no_nan
from ..tensor import Tensor
from .bio import decode_bio
from .utils import INT2TAG, eval_list_of_entity_sets, precision_recall_f1  # noqa: F401  (re-exported)


class ISequentialConfusionMatrixTF(IConfusionMatrixTF):
    """Confusion matrix over every position of [B, S] integer tensors (reference ner/metrics.py:23-37)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.__class__.__name__ == "ISequentialConfusionMatrixTF":
            raise Exception("This is an interface that cannot be instantiated")


class MacroF1Score(ISequentialConfusionMatrixTF):
    """mean_c 2 / (1/recall_c + 1/precision_c) -- plain divisions as in the reference (ner/metrics.py:39-57):
    a class that never occurs or is never predicted yields nan, exactly like the TF expression."""

    def _evaluate(self):
        m = self.confusion_matrix.astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            tp = np.diag(m)
            precision = tp / m.sum(axis=-1)
            recall = tp / m.sum(axis=-2)
            return float(np.mean(2.0 / ((1.0 / recall) + (1.0 / precision))))


class Accuracy(ISequentialConfusionMatrixTF):
    def _evaluate(self):
        m = self.confusion_matrix.astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            return float(np.trace(m) / m.sum())


def _host(x):
    if isinstance(x, Tensor):
        return x.numpy()
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        return x.numpy()
    return np.asarray(x)


class EntityF1(IMetric):
    """Entity-level F1 accumulated batch by batch (reference ner/metrics.py:8-21 over BioCSequenceDecoder).

    Each batch is a dict (or list of dicts) of arrays with leading batch dimension:
        identifier   [B]        document key (windows of one document are concatenated in arrival order)
        spans        [B, S, 2]  character span of every token
        tags_int_pred[B, S]     predicted label indices (TAG2INT of polus/ner/utils.py:9-15)
        is_prediction[B, S]     1 where the token's prediction counts (window overlap / padding excluded)
        tags_int     [B, S]     gold label indices   -- or pass gold={identifier: [(start, end, type)]} to __init__
    """

    def __init__(self, gold=None, int2tag=None):
        super().__init__()
        self.gold = {str(k): set(map(tuple, v)) for k, v in gold.items()} if gold is not None else None
        self.int2tag = dict(int2tag) if int2tag is not None else dict(INT2TAG)
        self.reset()

    def reset(self):
        self.docs = {}
        self.counts = {"tags": 0, "inside_tag_after_other_tag": 0, "inside_tag_with_different_entity_type": 0}

    def _samples_from_batch(self, samples):
        if isinstance(samples, dict):
            samples = [samples]
        for batch in samples:
            ident = _host(batch["identifier"])
            spans = _host(batch["spans"])
            pred = _host(batch["tags_int_pred"])
            keep = _host(batch["is_prediction"])
            true = _host(batch["tags_int"]) if "tags_int" in batch else None
            for i in range(len(ident)):
                key = ident[i].decode() if isinstance(ident[i], bytes) else str(ident[i])
                doc = self.docs.setdefault(key, {"spans": [], "pred": [], "true": []})
                sel = np.nonzero(keep[i] == 1)[0]
                doc["spans"].extend(map(tuple, spans[i][sel].tolist()))
                doc["pred"].extend(self.int2tag[int(v)] for v in pred[i][sel])
                if true is not None:
                    doc["true"].extend(self.int2tag[int(v)] for v in true[i][sel])

    def evaluate_ner(self):
        true_list, pred_list = [], []
        for key, doc in self.docs.items():
            es, c = decode_bio(doc["pred"], doc["spans"], allow_errors=True)
            for k in self.counts:
                self.counts[k] += c[k]
            if self.gold is not None:
                gold = self.gold.get(key, set())
            else:
                gold, _ = decode_bio(doc["true"], doc["spans"], allow_errors=True)
            true_list.append(gold)
            pred_list.append(es)
        return eval_list_of_entity_sets(true_list, pred_list)

    def _evaluate(self):
        return self.evaluate_ner()["f1"]
