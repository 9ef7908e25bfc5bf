"""polus.utils (reference polus/utils.py:6-95), TF-free."""
import json
import random

import numpy as np


def set_random_seed(seed_value=42):
    """Seeds python, numpy, the host weight initialisers and the device dropout streams
    (reference seeds tf/random/numpy: utils.py:6-9)."""
    from . import nn, ops
    random.seed(seed_value)
    np.random.seed(seed_value)
    nn.set_initializer_seed(seed_value)
    ops.set_seed(seed_value)


def merge_dicts(*list_of_dicts):
    out = dict(list_of_dicts[0], **list_of_dicts[1])
    for d in list_of_dicts[2:]:
        out.update(d)
    return out


def flatten_dict(d):
    """Nested dict -> flat dict; on duplicate keys the LAST occurrence wins (utils.py:21-35)."""
    items = []
    for k, v in d.items():
        if isinstance(v, dict):
            items.extend(flatten_dict(v).items())
        else:
            items.append((k, v))
    return dict(items)


def unique(iterable, key=lambda x: x):
    return list({key(x): x for x in iterable}.values())


def is_jsonable(x):
    try:
        json.dumps(x)
        return True
    except (TypeError, OverflowError):
        return False


def complex_json_serializer(data):
    from .tensor import Tensor, _NAMES
    out = {}
    for k, v in data.items():
        if isinstance(v, dict):
            out[k] = complex_json_serializer(v)
        elif is_jsonable(v):
            out[k] = v
        elif isinstance(v, Tensor):
            out[k] = {"_class": "tensor", "dtype": _NAMES[v.dtype], "values": v.numpy().tolist()}
        elif isinstance(v, np.ndarray):
            out[k] = {"_class": "tensor", "dtype": v.dtype.name, "values": v.tolist()}
        else:
            raise ValueError(f"Cannot serialize {type(v)} please add a json serializer to this type of data")
    return out


def complex_json_deserializer(data):
    out = {}
    for k, v in data.items():
        if isinstance(v, dict):
            if "_class" not in v:
                out[k] = complex_json_deserializer(v)
            elif v["_class"] == "tensor":
                dt = {"bfloat16": "float32"}.get(v["dtype"], v["dtype"])
                out[k] = np.asarray(v["values"], dtype=dt)
            else:
                raise ValueError(f"Cannot deserialize {v['_class']} please add a json deserializer to this type of data")
        else:
            out[k] = v
    return out


class Singleton(type):
    def __init__(cls, *args, **kwargs):
        cls.__instance = None
        super().__init__(*args, **kwargs)

    def __call__(cls, *args, **kwargs):
        if cls.__instance is None:
            cls.__instance = super().__call__(*args, **kwargs)
        return cls.__instance
