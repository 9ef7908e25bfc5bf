"""polus.data (reference polus/data.py:49-136): DataLoader and the batch format that reaches train_step.

Only the part on the hot path's boundary is provided: `DataLoader(generator).to_tfDataset()` with the
reference's sharding rule -- sample i goes to rank `i mod size`, applied BEFORE user-side batching
(data.py:94-96) -- and a small `Dataset` carrying the tf.data verbs the reference's scripts chain
after it (map / cache / shuffle / batch / prefetch, tutorials/classifier_example.py:29-41).  Batches
are numpy arrays; the trainer stages them through pinned memory.  The cached loaders
(CachedDataLoader*, data.py:138-494: pickle chunk files) are a §8f "next" row.
"""
import types

import numpy as np

from . import PolusContext, hvd as _hvd, logger
from .core import find_dtype_and_shapes

AUTOTUNE = -1


class Dataset:
    """Minimal re-iterable pipeline with tf.data's semantics for the verbs polus scripts use."""

    def __init__(self, source, n=None):
        self._source = source  # callable returning an iterator
        self._n = n

    @staticmethod
    def from_generator(gen_fn, output_types=None, output_shapes=None):
        return Dataset(lambda: iter(gen_fn()))

    @staticmethod
    def from_tensor_slices(arrays):
        if isinstance(arrays, dict):
            n = len(next(iter(arrays.values())))
            return Dataset(lambda: ({k: v[i] for k, v in arrays.items()} for i in range(n)), n)
        if isinstance(arrays, (tuple, list)):
            n = len(arrays[0])
            return Dataset(lambda: (tuple(a[i] for a in arrays) for i in range(n)), n)
        return Dataset(lambda: iter(arrays), len(arrays))

    def __iter__(self):
        return self._source()

    def cardinality(self):
        return self._n if self._n is not None else -2  # tf.data.UNKNOWN_CARDINALITY

    def __len__(self):
        if self._n is None:
            raise TypeError("dataset length is unknown")
        return self._n

    def shard(self, num_shards, index):
        src = self._source
        n = None if self._n is None else (self._n - index + num_shards - 1) // num_shards
        return Dataset(lambda: (x for i, x in enumerate(src()) if i % num_shards == index), n)

    def map(self, f, num_parallel_calls=None):
        src = self._source
        return Dataset(lambda: (f(x) for x in src()), self._n)

    def cache(self):
        src, store = self._source, {}

        def it():
            if "data" not in store:
                store["data"] = list(src())
            return iter(store["data"])
        return Dataset(it, self._n)

    def shuffle(self, buffer_size, seed=None):
        src = self._source
        rng = np.random.default_rng(seed)

        def it():
            buf = []
            for x in src():
                buf.append(x)
                if len(buf) >= buffer_size:
                    j = int(rng.integers(len(buf)))
                    buf[j], buf[-1] = buf[-1], buf[j]
                    yield buf.pop()
            rng.shuffle(buf)
            yield from buf
        return Dataset(it, self._n)

    def take(self, count):
        src = self._source
        n = count if self._n is None else min(count, self._n)
        return Dataset(lambda: (x for i, x in zip(range(count), src())), n)

    def repeat(self, count=None):
        src = self._source

        def it():
            k = 0
            while count is None or k < count:
                yield from src()
                k += 1
        return Dataset(it, None if (count is None or self._n is None) else self._n * count)

    def prefetch(self, buffer_size=None):
        return self

    def batch(self, batch_size, drop_remainder=False):
        src = self._source

        def stack(items):
            first = items[0]
            if isinstance(first, dict):
                return {k: np.stack([np.asarray(i[k]) for i in items]) for k in first}
            if isinstance(first, (tuple, list)):
                return tuple(stack([i[j] for i in items]) for j in range(len(first)))
            return np.stack([np.asarray(i) for i in items])

        def it():
            buf = []
            for x in src():
                buf.append(x)
                if len(buf) == batch_size:
                    yield stack(buf)
                    buf = []
            if buf and not drop_remainder:
                yield stack(buf)
        n = None
        if self._n is not None:
            n = self._n // batch_size if drop_remainder else -(-self._n // batch_size)
        return Dataset(it, n)


class DataLoader:
    """Wraps a python generator of dict samples (reference data.py:49-136)."""

    def __init__(self, sample_generator, magic_k=10):
        super().__init__()
        self.name = sample_generator.__name__ if sample_generator is not None else "None"
        self.sample_generator = sample_generator
        self.magic_k = magic_k

    def to_tfDataset(self):
        self.dtypes, self.shapes = find_dtype_and_shapes(self, k=self.magic_k)
        ds = Dataset.from_generator(lambda: self, output_types=self.dtypes, output_shapes=self.shapes)
        if PolusContext().is_horovod_enabled():
            hvd = _hvd()
            # NB: the reference shards by local_rank, not rank (data.py:96) -- single node only
            ds = ds.shard(num_shards=hvd.size(), index=hvd.local_rank())
        return ds

    def set_name(self, _name):
        self.name = _name

    @property
    def __name__(self):
        return f"{self.__class__.__name__}_{self.name}"

    def __iter__(self):
        if not isinstance(self.sample_generator, types.GeneratorType):
            gen = self.sample_generator()
            if isinstance(gen, types.GeneratorType):
                return gen
            raise ValueError("The sample_generator that was set in the DataLoader was a function that did not return "
                             "an generator, it must return a generator")
        return iter(self.sample_generator)

    def get_n_samples(self):
        if not hasattr(self, "n_samples"):
            logger.info("this dataset does not have the number of samples in cache so it will take some time to counting")
            self.n_samples = sum(1 for _ in self)
        return self.n_samples


class IAccelerated_Map:
    def __init__(self):
        super().__init__()
        self.__name__ = self.__class__.__name__
        if self.__class__.__name__ == "IAccelerated_Map":
            raise Exception("This is an interface that cannot be instantiated")

    def build(self):
        raise NotImplementedError("build method was not implemented")


def access_embeddings(func):
    """Call `func(**kwargs["embeddings"])` when the config nests the arguments under "embeddings" (data.py:510-521)."""
    def function_wrapper(*args, **kwargs):
        if isinstance(kwargs, dict) and "embeddings" in kwargs:
            return func(**kwargs["embeddings"])
        return func(*args, **kwargs)
    return function_wrapper


@access_embeddings
def build_bert_embeddings(checkpoint, bert_layer_index=None, **kwargs):
    """Frozen encoder used as `train_map_f` "GPU preprocessing" (reference data.py:523-545): returns
    `embeddings(input_ids=..., attention_mask=..., token_type_ids=...)` -> {last_hidden_state, pooler_output},
    computed without recording gradients.  With `bert_layer_index` the model is cut there first (split_bert_model)
    and only the lower layers run.  `checkpoint` is resolved offline (BertConfig / dict / directory / BertModel)."""
    from . import ops
    from .models import BertModel, split_bert_model, split_bert_model_from_checkpoint
    if isinstance(checkpoint, BertModel):
        bert_model = checkpoint
        if bert_layer_index is not None:
            bert_model = split_bert_model(bert_model, bert_layer_index, return_post_bert_model=False)
    elif bert_layer_index is not None:
        bert_model = split_bert_model_from_checkpoint(checkpoint, bert_layer_index, return_post_bert_model=False)
    else:
        from .models import BertConfig
        import json
        import os
        if isinstance(checkpoint, BertConfig):
            bert_model = BertModel(checkpoint)
        elif isinstance(checkpoint, dict):
            bert_model = BertModel(BertConfig(**checkpoint))
        else:
            with open(os.path.join(checkpoint, "config.json")) as f:
                bert_model = BertModel(BertConfig(**json.load(f)))

    def embeddings(**kw):
        with ops.no_grad():
            return bert_model(**kw, training=False)
    embeddings.model = bert_model
    return embeddings
