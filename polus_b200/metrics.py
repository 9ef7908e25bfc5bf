"""polus.metrics (reference polus/metrics.py:7-91).  The confusion matrix is built on the device
(polus_confusion_matrix); macro-F1 keeps the reference's float64 divide_no_nan arithmetic on the host
(K x K integers, nothing to accelerate)."""
import numpy as np

from . import _lib, device, nn
from .tensor import I32, Tensor


class IMetric:
    def __init__(self, reduce_f=None):
        super().__init__()
        if self.__class__.__name__ == "IMetric":
            raise Exception("This is an interface that cannot be instantiated")
        self.name = self.__class__.__name__
        self.reduce_f = reduce_f

    def samples_from_batch(self, samples):
        if self.reduce_f is not None:
            samples = self.reduce_f(samples)
        self._samples_from_batch(samples)

    def _samples_from_batch(self, samples):
        raise Exception("_samples_from_batch was internally called, but is not implemented")

    def reset(self):
        raise Exception("clear was called, but is not implemented")

    def _evaluate(self):
        raise Exception("_evaluate was internally called, but is not implemented")

    def evaluate(self):
        measure = self._evaluate()
        if isinstance(measure, Tensor):
            measure = measure.numpy()
        self.reset()
        return measure


class IConfusionMatrixTF(IMetric):
    def __init__(self, num_classes, reduce_f=None):
        super().__init__(reduce_f=reduce_f)
        if self.__class__.__name__ == "IConfusionMatrixTF":
            raise Exception("This is an interface that cannot be instantiated")
        self.num_classes = num_classes
        self.reset()

    def _samples_from_batch(self, samples):
        self.confusion_matrix += self._build_confusion_matrix(*samples)

    def _build_confusion_matrix(self, y_true, y_pred):
        """tf.math.confusion_matrix(y_true, y_pred): rows = labels, cols = predictions.
        (ValidationDataCallback hands (prediction, label) in that order, callbacks.py:236 -- kept as is.)"""
        yt = nn.as_tensor(np.asarray(y_true).reshape(-1) if not isinstance(y_true, Tensor) else y_true, I32)
        yp = nn.as_tensor(np.asarray(y_pred).reshape(-1) if not isinstance(y_pred, Tensor) else y_pred, I32)
        n = yt.size
        assert yp.size == n
        cm = Tensor((self.num_classes, self.num_classes), I32, zero=True)
        _lib.call("polus_confusion_matrix", yt.ptr, yp.ptr, n, self.num_classes, cm.ptr, device.stream())
        return cm.numpy()

    def reset(self):
        self.confusion_matrix = np.zeros((self.num_classes, self.num_classes), dtype=np.int32)


def _divide_no_nan(a, b):
    a = np.broadcast_to(np.asarray(a, np.float64), np.shape(b)).astype(np.float64)
    b = np.asarray(b, np.float64)
    out = np.zeros_like(b)
    nz = b != 0
    out[nz] = a[nz] / b[nz]
    return out


class MacroF1Score(IConfusionMatrixTF):
    def _evaluate(self):
        m = self.confusion_matrix
        tp = np.diag(m).astype(np.float64)
        fp_tp = m.sum(axis=-1).astype(np.float64)
        fn_tp = m.sum(axis=-2).astype(np.float64)
        precision = _divide_no_nan(tp, fp_tp)
        recall = _divide_no_nan(tp, fn_tp)
        inv_precision = _divide_no_nan(1.0, precision)
        inv_recall = _divide_no_nan(1.0, recall)
        return float(np.mean(_divide_no_nan(2.0, inv_precision + inv_recall)))
