"""Load tests/golden/reference_vectors.json and materialise the inputs of each case."""
import json
import math
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")
DT = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64}


def load(kind=None):
    with open(_PATH) as f:
        cases = json.load(f)["cases"]
    return [c for c in cases if kind is None or c["kind"] == kind]


def linspace(a, b, n, dtype=np.float64):
    """ndarray::Array::linspace: start + step * i with step = (end - start) / (n - 1)
    (reproduces tests/interp2d.rs:104-236 bit-exactly)."""
    dt = np.dtype(dtype).type
    a, b = dt(a), dt(b)
    step = (b - a) / dt(n - 1) if n > 1 else dt(0)
    return np.array([a + step * dt(i) for i in range(n)], dtype=dtype)


def materialise(spec, dtype):
    if isinstance(spec, dict) and "linspace" in spec:
        a, b, n = spec["linspace"]
        v = linspace(a, b, n, dtype)
        if "repeat_each" in spec:
            v = np.repeat(v, spec["repeat_each"])
        if "tile" in spec:
            v = np.tile(v, spec["tile"])
        if "shape" in spec:
            v = v.reshape(spec["shape"])
        return v
    return np.array(spec, dtype=dtype)


def index_grid(name):
    if name == "linspace":
        return linspace(0.0, 10.0, 11)
    if name == "exp2":
        return np.array([2.0 ** i for i in range(11)])
    if name == "ln_1p":
        return np.array([math.log1p(float(i)) for i in range(11)])
    raise KeyError(name)


def index_query(v):
    if isinstance(v, dict):
        if "exp2" in v:
            return 2.0 ** v["exp2"]          # 2f64.powf(x)
        if "ln_1p" in v:
            return math.log1p(v["ln_1p"])
    if v == "inf":
        return math.inf
    if v == "-inf":
        return -math.inf
    return float(v)


def default_axis(n, dtype):
    """Interp1DBuilder::new: x = 0..len cast to the element type (interp1d/mod.rs:399-410)"""
    return np.arange(n).astype(dtype)
