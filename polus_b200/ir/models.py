"""Ranking models for polus.ir.  The reference ships only the trainer (polus/ir/training.py) and expects a user model
exposing encode_query / encode_document / query_projection / document_projection (ir/training.py:50-51,82-83);
these two classes are the assembled configurations BASELINE.json names."""
from .. import nn, ops
from ..models import BertConfig, BertModel, PolusModel
from ..tensor import F32


class BertBiEncoder(PolusModel):
    """Bi-encoder dense retrieval: a (frozen) BERT encodes queries and documents to h[:,0,:]; trainable linear
    projections map them to the scoring space (what EfficientDenseRetrievalTrainer differentiates)."""

    def __init__(self, config=None, projection_dim=128, share_encoder=True, name="bert_bi_encoder", **kwargs):
        super().__init__(name=name)
        self.config = config or BertConfig(**kwargs)
        self.query_encoder = BertModel(self.config, add_pooling_layer=False)
        self.doc_encoder = self.query_encoder if share_encoder else BertModel(self.config, add_pooling_layer=False)
        self.query_projection = nn.Dense(projection_dim, name="query_projection")
        self.document_projection = nn.Dense(projection_dim, name="document_projection")

    def sublayers(self):
        return [self.query_projection, self.document_projection]  # encoders are frozen: not trainable

    def encode_query(self, q, training=False):
        with ops.no_grad():
            return self.query_encoder(**q, training=False)["pooler_output"]

    def encode_document(self, d, training=False):
        with ops.no_grad():
            return self.doc_encoder(**d, training=False)["pooler_output"]

    def call(self, question, document, training=False):
        q = self.query_projection(self.encode_query(question), training=training)
        d = self.document_projection(self.encode_document(document), training=training)
        return ops.reduce_sum(ops.mul(ops.cast(q, F32), ops.cast(d, F32)), axis=-1)


def in_batch_scores(q, d_pos, *d_negs):
    """compute_scores for EfficientDenseRetrievalTrainer: positives = diag(q d^T); negatives = the other
    documents of the batch (+ explicit negatives)."""
    q, d_pos = ops.cast(q, F32), ops.cast(d_pos, F32)
    pos = ops.reduce_sum(ops.mul(q, d_pos), axis=-1)          # [B]
    neg = ops.matmul(q, d_pos, transpose_b=True)              # [B,B] (diagonal = positives; the loss masks it)
    return pos, neg


def softmax_ranking_loss(pos_scores, neg_scores):
    """-log softmax over [in-batch documents]: cross entropy of row i against column i."""
    B = neg_scores.shape[0]
    return ops.cross_entropy(0, neg_scores, ops.constant_arange(B))


class BertCrossEncoder(PolusModel):
    """Cross-encoder ranker (BASELINE.json config 4): [CLS] q [SEP] d [SEP] -> BERT -> Dense(1) on h[:,0,:].
    Called on {"input_ids", "attention_mask", "token_type_ids"} holding POSITIVE pairs in the first half of the
    batch and NEGATIVE pairs in the second half; pair i and i + B/2 share the query."""

    def __init__(self, config=None, name="bert_cross_encoder", **kwargs):
        super().__init__(name=name)
        self.config = config or BertConfig(**kwargs)
        self.bert = BertModel(self.config, add_pooling_layer=False)
        self.score = nn.Dense(8, name="score")  # 8 outputs keep the GEMM TMA-aligned; column 0 is the score

    def sublayers(self):
        return [self.bert, self.score]

    def call(self, input_ids=None, attention_mask=None, token_type_ids=None, training=False, **unused):
        h = self.bert(input_ids=input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                      training=training)["pooler_output"]
        return self.score(h, training=training)


def pairwise_softplus_loss(y_unused, scores):
    """mean softplus(s_neg - s_pos) over pairs (RankNet / pairwise logistic) on column 0 of the score head."""
    s = ops.cast(scores, F32)
    B2, C = s.shape
    half = B2 // 2
    s0 = ops.gather_cols0(s)                                   # [B2]
    pos = ops.slice_rows(s0, 0, half)
    neg = ops.slice_rows(s0, half, half)
    return ops.reduce_mean(ops.unary("softplus", ops.sub(neg, pos)))


def explicit_negative_scores(q, d_pos, *d_negs):
    """compute_scores for k explicit negatives per query (polus/ir/training.py:94-107 hands compute_scores the
    projected query, positive and the k negative representations): pos [B], neg = list of k [B] dot products."""
    q = ops.cast(q, F32)
    pos = ops.reduce_sum(ops.mul(q, ops.cast(d_pos, F32)), axis=-1)
    negs = [ops.reduce_sum(ops.mul(q, ops.cast(n, F32)), axis=-1) for n in d_negs]
    return pos, negs


def pairwise_softplus_ranking_loss(pos_scores, neg_scores):
    """mean over the batch and the k negatives of softplus(s_neg - s_pos)."""
    negs = neg_scores if isinstance(neg_scores, (list, tuple)) else [neg_scores]
    total = None
    for n in negs:
        term = ops.reduce_mean(ops.unary("softplus", ops.sub(n, pos_scores)))
        total = term if total is None else ops.add(total, term)
    return ops.unary("scale", total, 1.0 / len(negs))
