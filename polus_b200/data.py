"""polus.data (reference polus/data.py:49-136): DataLoader and the batch format that reaches train_step.

Only the part on the hot path's boundary is provided: `DataLoader(generator).to_tfDataset()` with the
reference's sharding rule -- sample i goes to rank `i mod size`, applied BEFORE user-side batching
(data.py:94-96) -- and a small `Dataset` carrying the tf.data verbs the reference's scripts chain
after it (map / cache / shuffle / batch / prefetch, tutorials/classifier_example.py:29-41).  Batches
are numpy arrays; the trainer stages them through pinned memory.

CachedDataLoader / CachedDataLoaderwLookup (data.py:138-494) keep the reference's on-disk chunk
format, so caches written by either implementation are read by the other:
  `<folder>/<ident>__chunk<N>_<generator name>.index`   JSON {"files": [...], "cache_chunk_size": N, "n_samples": M
                                                          [, "lookup_file": path]}
  `<base>_0000.part`, `<base>_0001.part`, ...            pickle.dump(list of <= N samples)
  `<base>.lookup`                                        pickle.dump(lookup object)
"""
import json
import os
import pickle
import random
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


class CachedDataLoader(DataLoader):
    """DataLoader whose samples are materialised once into pickled chunk files and streamed back from them
    (reference data.py:138-421).  `pre_shuffle()` permutes the CHUNK order on every pass (samples inside a chunk keep
    their order), `merge()` concatenates the indices of several cached loaders, `from_cached_index()` reopens one.
    If building the cache raises, every file written so far is removed before the exception propagates."""

    def __init__(self, sample_generator=None, clean_up_function=None, cache_additional_identifier="", cache_chunk_size=8192,
                 cache_folder=os.path.join(".polus_cache", "data"), cache_index=None, **kwargs):
        assert sample_generator is not None or cache_index is not None
        self.cache_folder = cache_folder
        self.cache_chunk_size = cache_chunk_size
        self.cache_additional_identifier = cache_additional_identifier
        self.shuffle_blocks = False
        self.clean_up_function = clean_up_function
        self.cache_index = cache_index
        # a loader assembled by merge() has an index but no index file of its own
        self.cache_index_path = cache_index.get("cache_index_path") if cache_index is not None else None
        try:
            super().__init__(sample_generator=self._build_sample_generator(sample_generator), **kwargs)
        except Exception:
            if cache_index is None:
                logger.info("An error has occured so all the created files will be deleted")
                self.clean()
            raise

    # ------------------------------------------------------------------ constructors from existing caches
    @classmethod
    def from_cached_index(cls, index_path):
        index = cls.read_index(index_path)
        index["cache_index_path"] = index_path
        return cls(cache_index=index)

    @staticmethod
    def _merged_index(loaders):
        assert len(loaders) > 1
        merged = {"files": [], "cache_chunk_size": 0, "n_samples": 0}
        indices = [CachedDataLoader.read_index(dl.cache_index_path) for dl in loaders]
        for index in indices:
            merged["files"].extend(index["files"])
            merged["n_samples"] += index["n_samples"]
            # loaders may have been cached with different chunk sizes: the union reports the largest
            merged["cache_chunk_size"] = max(merged["cache_chunk_size"], index["cache_chunk_size"])
        return merged, indices

    @classmethod
    def merge(cls, *cache_dataloaders):
        return cls(cache_index=CachedDataLoader._merged_index(cache_dataloaders)[0])

    @staticmethod
    def read_index(file_path):
        with open(file_path, "r") as f:
            return json.load(f)

    def write_index_file(self, index_info):
        with open(self.cache_index_path, "w") as f:
            json.dump(index_info, f)

    # ------------------------------------------------------------------ cache construction
    def _cache_base_name(self, sample_generator):
        ident = f"{self.cache_additional_identifier}_" if self.cache_additional_identifier != "" else ""
        return f"{ident}_chunk{self.cache_chunk_size}_{sample_generator.__name__}"

    def _build_sample_generator(self, sample_generator):
        if self.cache_index is not None:
            return self._generator_from_index()
        os.makedirs(self.cache_folder, exist_ok=True)
        self.cache_base_name = self._cache_base_name(sample_generator)
        self.cache_base_path = os.path.join(self.cache_folder, self.cache_base_name)
        self.cache_index_path = f"{self.cache_base_path}.index"
        if os.path.exists(self.cache_index_path):
            logger.info("We found a compatible cache file for this DataLoader")
            sample_generator = None
        else:
            logger.info(f"DataLoader will store the samples in {self.cache_base_path}, with a max_sample per file of "
                        f"{self.cache_chunk_size}, this may take a while")
        return self._build_cache_generator(sample_generator)

    def _build_cache_generator(self, generator=None):
        if generator is not None:
            index_info = {"files": [], "cache_chunk_size": self.cache_chunk_size}
            self.cache_index = index_info  # so that clean() sees the chunk files already written if the generator raises
            chunk, n_samples = [], 0

            def flush():
                path = f"{self.cache_base_path}_{len(index_info['files']):04}.part"
                with open(path, "wb") as f:
                    pickle.dump(chunk, f)
                index_info["files"].append(path)

            logger.info("Starting to cache the dataset, this may take a while")
            for sample in generator():
                n_samples += 1
                chunk.append(sample)
                if len(chunk) >= self.cache_chunk_size:
                    flush()
                    chunk = []
            if chunk:
                flush()
            index_info["n_samples"] = n_samples
            self.write_index_file(index_info)
        if self.clean_up_function is not None:
            logger.info("Executing the clean up function after the cached dataset was created")
            self.clean_up_function()
        self.cache_index = self.__class__.read_index(self.cache_index_path)
        return self._generator_from_index()

    def _generator_from_index(self):
        self.n_samples = self.cache_index["n_samples"]
        self.cache_chunk_size = self.cache_index["cache_chunk_size"]
        logger.info(f"Total number of samples in dataset: {self.n_samples}")

        def generator():
            order = list(range(len(self.cache_index["files"])))
            if self.shuffle_blocks:
                random.shuffle(order)
            for k in order:
                with open(self.cache_index["files"][k], "rb") as f:
                    yield from pickle.load(f)
        return generator

    # ------------------------------------------------------------------ maintenance
    def clean(self):
        index = getattr(self, "cache_index", None)
        for path in (index or {}).get("files", []):
            if os.path.exists(path):
                os.remove(path)
        path = getattr(self, "cache_index_path", None)
        if path is not None and os.path.exists(path):
            os.remove(path)

    def pre_shuffle(self):
        """Chunk files are visited in a fresh random order on every pass."""
        self.shuffle_blocks = True
        return self

    def add_lookup_data(self, lookup_object):
        """Writes `<base>.lookup` and returns this cache reopened as a CachedDataLoaderwLookup."""
        lookup_file = f"{os.path.splitext(self.cache_index_path)[0]}.lookup"
        with open(lookup_file, "wb") as f:
            pickle.dump(lookup_object, f)
        return self.add_lookup_data_path(lookup_file)

    def add_lookup_data_path(self, lookup_data_path):
        assert os.path.exists(lookup_data_path)
        self.cache_index["lookup_file"] = lookup_data_path
        with open(self.cache_index_path, "w") as f:
            json.dump(self.cache_index, f)
        return CachedDataLoaderwLookup.from_cached_index(self.cache_index_path)

    def deep_copy(self, path=None, suffix=None):
        """Writes a second index file over the same chunk files and points this loader at it.  (The reference builds the
        default name from `suffix` even when it is None -> "..._None.index", data.py:403-406; here None means "copy".)"""
        if path is None:
            path = f"{os.path.splitext(self.cache_index_path)[0]}_{'copy' if suffix is None else suffix}.index"
        with open(path, "w") as f:
            json.dump(self.cache_index, f)
        self.cache_index_path = path


class CachedDataLoaderwLookup(CachedDataLoader):
    """CachedDataLoader that also persists one pickled lookup object next to the chunks (reference data.py:424-494)."""

    def __init__(self, *args, lookup_data=None, cache_index=None, **kwargs):
        if cache_index is not None and "lookup_file" in cache_index:
            lookup_data = self._load_lookup_data(cache_index["lookup_file"])
        if lookup_data is None:
            raise ValueError("Do not use CachedDataLoaderwLookup without setting a lookup_data, instead use CachedDataLoader")
        self.lookup_data = lookup_data
        super().__init__(*args, cache_index=cache_index, **kwargs)

    def get_lookup_data(self):
        return self.lookup_data

    @staticmethod
    def _load_lookup_data(lookup_file):
        with open(lookup_file, "rb") as f:
            return pickle.load(f)

    @classmethod
    def merge(cls, *cache_dataloaders):
        """Union of the chunk lists; the lookup objects (sequences) are concatenated in the same order."""
        merged, indices = CachedDataLoader._merged_index(cache_dataloaders)
        lookup = []
        for index in indices:
            lookup.extend(cls._load_lookup_data(index["lookup_file"]))
        return cls(cache_index=merged, lookup_data=lookup)

    def clean(self):
        index = getattr(self, "cache_index", None)
        lookup_file = (index or {}).get("lookup_file")
        if lookup_file is not None and os.path.exists(lookup_file):
            os.remove(lookup_file)
        super().clean()

    def write_index_file(self, index_info):
        lookup_file = f"{self.cache_base_path}.lookup"
        with open(lookup_file, "wb") as f:
            pickle.dump(self.lookup_data, f)
        index_info["lookup_file"] = lookup_file
        super().write_index_file(index_info)

    def _generator_from_index(self):
        if "lookup_file" in self.cache_index:  # (a merged loader carries its lookup object in memory only)
            self.lookup_data = self._load_lookup_data(self.cache_index["lookup_file"])
        return super()._generator_from_index()


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
    from .models import bert_model_from_checkpoint, split_bert_model
    bert_model = bert_model_from_checkpoint(checkpoint)
    if bert_layer_index is not None:
        bert_model = split_bert_model(bert_model, bert_layer_index, return_post_bert_model=False)

    def embeddings(**kw):
        with ops.no_grad():
            return bert_model(**kw, training=False)
    embeddings.model = bert_model
    return embeddings
