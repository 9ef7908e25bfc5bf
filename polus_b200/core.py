"""polus.core (reference polus/core.py:34-142), TF-free."""
import logging
import os

import numpy as np


def set_jit_compile(mode: bool):
    """Kept for API compatibility: the XLA switch has no meaning here (every step is already one
    captured CUDA graph); the flag round-trips through POLUS_JIT exactly like the reference."""
    os.environ["POLUS_JIT"] = str(mode)


def get_jit_compile():
    if os.environ.get("POLUS_JIT") is None:
        set_jit_compile(False)
    return os.environ.get("POLUS_JIT") == "True"


def find_dtype_and_shapes(data_generator, k=10):
    """Infer dtypes and (possibly dynamic = None) shapes from the first k dict samples (core.py:58-111)."""
    if k == -1:
        samples = [s for s in data_generator]
    else:
        gen = iter(data_generator)
        samples = []
        for _ in range(k):
            try:
                samples.append(next(gen))
            except StopIteration:
                break
    if not samples or not isinstance(samples[0], dict):
        raise ValueError(f"The find_dtype_and_shapes only supports when the sample came from generator are dict but found {type(samples[0]) if samples else None}")
    dtypes = {k_: np.asarray(v).dtype for k_, v in samples[0].items()}
    shapes = {k_: tuple(np.asarray(v).shape) for k_, v in samples[0].items()}
    for s in samples[1:]:
        assert set(s.keys()) == set(samples[0].keys())
        for k_, v in s.items():
            shp = np.asarray(v).shape
            assert len(shp) == len(shapes[k_])
            if tuple(shp) != shapes[k_]:
                shapes[k_] = tuple(a if (a is not None and a == b) else None for a, b in zip(shapes[k_], shp))
    return dtypes, shapes


def execute_if(condition_var, error_message="", on=True):
    def decorator(func):
        def function_wrapper(self, *args, **kwargs):
            if getattr(self, condition_var) == on:
                return func(self, *args, **kwargs)
            if error_message != "":
                print(error_message)
        return function_wrapper
    return decorator


class BaseLogger:
    """Base of the classes that log through the package logger.  polus/ner/utils.py:1 imports this name from polus.core,
    which does not define it at the surveyed HEAD (the import fails there); this is the evident intent."""

    def __init__(self):
        self.logger = logging.getLogger("polus")
