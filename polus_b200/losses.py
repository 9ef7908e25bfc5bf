"""polus.losses (reference polus/losses.py:5-42) + the Keras loss the tutorial uses."""
import numpy as np

from . import nn, ops
from .tensor import F32, I32, Tensor


def weighted_softmax_cross_entropy_from_logits(class_weights):
    cw = Tensor.from_numpy(np.asarray(class_weights, np.float32), F32)

    def weighted_softmax_cross_entropy_from_logits_loss(y_true, y_pred):
        return ops.cross_entropy(1, nn.as_tensor(y_pred), ops.cast(nn.as_tensor(y_true), F32), cw)
    return weighted_softmax_cross_entropy_from_logits_loss


def weighted_sigmoid_cross_entropy_from_logits(class_weights, negative_weight):
    cw = Tensor.from_numpy(np.asarray(class_weights, np.float32), F32)

    def weighted_sigmoid_cross_entropy_from_logits_loss(y_true, y_pred):
        return ops.cross_entropy(2, nn.as_tensor(y_pred), ops.cast(nn.as_tensor(y_true), F32), cw, negative_weight)
    return weighted_sigmoid_cross_entropy_from_logits_loss


class SparseCategoricalCrossentropy:
    """tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True) (tutorials/classifier_example.py:55)."""

    def __init__(self, from_logits=True, **kwargs):
        if not from_logits:
            raise ValueError("only from_logits=True is supported")

    def __call__(self, y_true, y_pred):
        y = nn.as_tensor(y_true, I32)
        return ops.cross_entropy(0, nn.as_tensor(y_pred), y if y.dtype == I32 else ops.cast(y, I32))
