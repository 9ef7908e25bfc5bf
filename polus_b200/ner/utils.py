"""polus.ner.utils (reference polus/ner/utils.py): the label contract, strict entity-level evaluation and the decoder that
turns batches of predicted BIO tags back into per-document entity sets.

The reference's decoder works on its BioC corpus objects (polus/ner/elements.py: Corpus / Collection / Document /
EntitySet, out of scope here -- SURVEY.md §2).  `BioCSequenceDecoder` below keeps its interface and data flow
(samples_from_batch -> decode -> evaluate_ner, keyed corpus / group / identifier, `is_prediction` filtering, BIO error
counters) over plain data: an entity set is a Python set of (start, end, type) tuples; the gold side is either a nested
dict {corpus: {group: {identifier: {"text": str, "es": iterable of (start, end, type)}}}} or any corpus objects that
iterate like the reference's ((group, collection) pairs, (identifier, document) pairs, document.text() /
document.get_entity_set()).  `get_collections*` would have to rebuild BioC Collection objects and is not provided.
"""
import math

import numpy as np

from ..core import BaseLogger
from .bio import decode_bio

TAG2INT = {"PAD": 0, "O": 1, "B-Chemical": 2, "I-Chemical": 3}
INT2TAG = {v: k for k, v in TAG2INT.items()}


def precision_recall_f1(tp, fp, fn, return_nan=True):
    """polus/ner/utils.py:231-248: each ratio is nan (or 0) when its denominator is empty; f1 = tp / (tp + (fp + fn)/2)."""
    bad = math.nan if return_nan else 0.0
    precision = tp / (tp + fp) if tp + fp != 0 else bad
    recall = tp / (tp + fn) if tp + fn != 0 else bad
    f1 = tp / (tp + 0.5 * (fp + fn)) if tp + 0.5 * (fp + fn) != 0 else bad
    return precision, recall, f1


def empty_results(counts=True):
    """polus/ner/utils.py:251-266."""
    results = {"tp": 0, "fp": 0, "fn": 0} if counts else {}
    results.update({"precision": 0.0, "recall": 0.0, "f1": 0.0})
    return results


def _as_entity_set(es):
    if hasattr(es, "get") and not isinstance(es, (set, frozenset, dict)):  # reference EntitySet: .get() -> entities
        es = es.get()
    out = set()
    for e in es:
        if isinstance(e, (tuple, list)):
            out.add(tuple(e))
        else:  # reference Entity objects: .span (start, end) and .typ
            out.add((e.span[0], e.span[1], e.typ))
    return out


def eval_list_of_entity_sets(true, pred, return_nan=True):
    """Strict evaluation (polus/ner/utils.py:269-308): a predicted entity is a true positive only if span and type match a
    gold entity exactly; TP = |true ∩ pred| per document, counts summed over documents, ratios taken at the end."""
    assert isinstance(true, list) and isinstance(pred, list) and len(true) == len(pred)
    results = empty_results(counts=True)
    for t_es, p_es in zip(true, pred):
        t_es, p_es = _as_entity_set(t_es), _as_entity_set(p_es)
        tp = len(t_es & p_es)
        results["tp"] += tp
        results["fp"] += len(p_es) - tp
        results["fn"] += len(t_es) - tp
    results["precision"], results["recall"], results["f1"] = precision_recall_f1(results["tp"], results["fp"], results["fn"], return_nan)
    return results


def _scalar(v):
    """One field of one unbatched sample -> python value (tf tensors, device tensors, numpy, bytes)."""
    if hasattr(v, "numpy") and not isinstance(v, np.ndarray):
        v = v.numpy()
    if isinstance(v, np.ndarray):
        v = v.tolist() if v.ndim else v.item()
    if isinstance(v, bytes):
        v = v.decode()
    return v


class BioCSequenceDecoder(BaseLogger):
    """Accumulates predicted tag sequences batch by batch and evaluates them against the gold entity sets of the corpora
    (reference polus/ner/utils.py:8-174)."""
    TAG2INT = TAG2INT
    INT2TAG = INT2TAG

    def __init__(self, corpora):
        super().__init__()
        self.documents_dict = {}
        self.documents = {}
        if isinstance(corpora, dict):
            for corpus, groups in corpora.items():
                self.documents[str(corpus)] = {
                    str(g): {str(i): {"text": d.get("text"), "es": _as_entity_set(d["es"])} for i, d in docs.items()}
                    for g, docs in groups.items()}
        else:
            for corpus in corpora:
                per_group = self.documents.setdefault(str(corpus), {})
                for group, collection in corpus:
                    per_group[str(group)] = {str(i): {"text": d.text(), "es": _as_entity_set(d.get_entity_set())}
                                             for i, d in collection}

    def clear_state(self):
        self.documents_dict = {}

    def samples_from_batch(self, samples):
        """samples: a dict (or list of dicts) of batched fields `corpus`, `group`, `identifier`, `spans` [B,S,2],
        `tags_int_pred` [B,S], `is_prediction` [B,S]; only positions with is_prediction == 1 are kept, in arrival order."""
        if isinstance(samples, dict):
            samples = [samples]
        for batch in samples:
            n = len(batch[next(iter(batch))])
            for i in range(n):
                corpus, group, identifier = (str(_scalar(batch[k][i])) for k in ("corpus", "group", "identifier"))
                spans = _scalar(batch["spans"][i])
                tags = [self.INT2TAG[int(t)] for t in _scalar(batch["tags_int_pred"][i])]
                keep = _scalar(batch["is_prediction"][i])
                doc = self.documents_dict.setdefault(corpus, {}).setdefault(group, {}).setdefault(identifier, {"spans": [], "tags": []})
                for s, t, k in zip(spans, tags, keep):
                    if k == 1:
                        doc["spans"].append(tuple(s))
                        doc["tags"].append(t)

    def decode(self):
        counts = {"tags": 0, "inside_tag_after_other_tag": 0, "inside_tag_with_different_entity_type": 0}
        for corpus, groups in self.documents_dict.items():
            for group, docs in groups.items():
                for identifier, doc in docs.items():
                    text = self.documents[corpus][group][identifier]["text"]
                    doc["es"], c = decode_bio(doc["tags"], doc["spans"], text, allow_errors=True)
                    for k in counts:
                        counts[k] += c[k]
        self.logger.info("Statistics about the BIO decoding process: tags={}, inside_tag_after_other_tag={}, "
                         "inside_tag_with_different_entity_type={}.".format(counts["tags"], counts["inside_tag_after_other_tag"],
                                                                            counts["inside_tag_with_different_entity_type"]))
        return counts

    def decode_from_samples(self, samples):
        self.samples_from_batch(samples)
        self.decode()

    def evaluate_ner_from_sample(self, samples):
        self.decode_from_samples(samples)
        return self._evaluate_ner()

    def evaluate_ner(self):
        self.decode()
        return self._evaluate_ner()

    def _evaluate_ner(self):
        true_list, pred_list = [], []
        for corpus, groups in self.documents_dict.items():
            for group, docs in groups.items():
                for identifier, doc in docs.items():
                    true_list.append(self.documents[corpus][group][identifier]["es"])
                    pred_list.append(doc["es"])
        results = eval_list_of_entity_sets(true_list, pred_list)
        self.clear_state()
        return results

    def get_collections(self):
        raise NotImplementedError("BioC Collection objects (polus/ner/elements.py) are outside this package; "
                                  "use decode() and read documents_dict[corpus][group][identifier]['es']")

    get_collections_from_samples = get_collections
