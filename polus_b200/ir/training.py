"""polus.ir.training (reference polus/ir/training.py:5-117): bi-encoder dense-retrieval trainer with
in-batch or k explicit negatives.  The frozen encoders run without recording gradients
(`forward_without_grads`), the projections + user `compute_scores` + user loss are differentiated."""
from .. import nn, ops
from ..tensor import I32
from ..training import BaseTrainer


class EfficientDenseRetrievalTrainer(BaseTrainer):
    def __init__(self, model, compute_scores, k_negatives=0, trainable_weights=None, *args, **kwargs):
        self.compute_scores = compute_scores
        self.k_negatives = k_negatives
        self.trainable_weights = model.trainable_weights if trainable_weights is None else trainable_weights
        super().__init__(model, *args, **kwargs)

    def __str__(self):
        return 'SimilarityTrainer'

    def forward_without_grads(self, question, positive_doc, negative_doc=None):
        q = self.model.encode_query(question, training=True)
        d = self.model.encode_document(positive_doc, training=True)
        if negative_doc is None:
            return q, d
        ids, mask = nn.as_tensor(negative_doc["input_ids"], I32), nn.as_tensor(negative_doc["attention_mask"], I32)
        B, k, S = ids.shape
        self.k_negatives = k
        # negative_docs[...][:, i, :] (ir/training.py:64) on the device: rows i, i+k, i+2k ... of the [B*k, S] view --
        # no host round trip, so the slicing is part of the captured step and follows every new batch
        ids2, mask2 = ids.view((B * k, S)), mask.view((B * k, S))
        negs = [self.model.encode_document({"input_ids": ops.gather_rows(ids2, i, k, B),
                                            "attention_mask": ops.gather_rows(mask2, i, k, B)}, training=True)
                for i in range(k)]
        return (q, d, negs)

    def forward_with_grads(self, question_rep, positive_doc_rep, negative_docs_rep=None):
        q = self.model.query_projection(question_rep, training=True)
        d = self.model.document_projection(positive_doc_rep, training=True)
        negs = []
        if negative_docs_rep is not None:
            negs = [self.model.document_projection(n, training=True) for n in negative_docs_rep]
        if self.post_process_logits is not None:
            q, d = self.post_process_logits(q), self.post_process_logits(d)
            negs = [self.post_process_logits(n) for n in negs]
        pos_scores, neg_scores = self.compute_scores(q, d, *negs)
        return pos_scores, neg_scores
