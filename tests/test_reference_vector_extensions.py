"""src/vector_extensions.rs:200-403 transcribed test-for-test; the computation runs on the GPU."""
import math

import numpy as np
import pytest

import golden_util as G
from ndarray_interp_b200 import Panic
from ndarray_interp_b200.vector_extensions import Monotonic, get_lower_index, monotonic_prop

pytestmark = pytest.mark.gpu


def lin():
    return G.linspace(0.0, 10.0, 11)


def exp():
    return np.array([2.0 ** i for i in range(11)])


def ln():
    return np.array([math.log1p(float(i)) for i in range(11)])


def test_outside_left():
    assert get_lower_index(lin(), -1.0) == 0


def test_outside_right():
    assert get_lower_index(lin(), 25.0) == 9


def test_left_border():
    assert get_lower_index(lin(), 0.0) == 0


def test_right_border():
    assert get_lower_index(lin(), 10.0) == 9


def test_exact_index():
    for i in range(10):
        assert get_lower_index(lin(), float(i)) == i


def test_index():
    q = np.array([i / 10.0 for i in range(100)])
    assert np.array_equal(get_lower_index(lin(), q), np.arange(100) // 10)


def test_pos_inf_index():
    assert get_lower_index(lin(), math.inf) == 9


def test_neg_inf_index():
    assert get_lower_index(lin(), -math.inf) == 0


def test_nan():
    with pytest.raises(Panic, match="not implemented: failed to convert NaN to usize"):
        get_lower_index(lin(), math.nan)


def test_exponential_exact_index():
    for i in range(10):
        assert get_lower_index(exp(), 2.0 ** i) == i


def test_exponential_index():
    q = np.array([2.0 ** (x / 10.0) for x in range(100)])
    assert np.array_equal(get_lower_index(exp(), q), np.arange(100) // 10)


def test_exponential_right_border():
    assert get_lower_index(exp(), 1024.0) == 9


def test_exponential_left_border():
    assert get_lower_index(exp(), 1.0) == 0


def test_log():
    q = np.array([math.log1p(x / 10.0) for x in range(100)])
    assert np.array_equal(get_lower_index(ln(), q), np.arange(100) // 10)


@pytest.mark.parametrize("case", G.load("monotonic"), ids=lambda c: c["name"])
def test_monotonic(case):
    x = np.array(case["x"], dtype=G.DT[case["dtype"]])
    view = x[::case["stride"]]
    expect = {"NotMonotonic": Monotonic.NotMonotonic, "RisingStrict": Monotonic.Rising(True),
              "Rising": Monotonic.Rising(False), "FallingStrict": Monotonic.Falling(True),
              "Falling": Monotonic.Falling(False)}[case["expect"]]
    assert monotonic_prop(view) == expect
    assert monotonic_prop(view[::1]) == expect


def test_monotonic_nan_semantics_match_the_state_machine():
    # vector_extensions.rs:136-170 (see tests/test_oracle_golden.py::test_monotonic_nan_semantics)
    nan = np.nan
    assert monotonic_prop(np.array([nan, 1.0])) == Monotonic.Falling(True)
    assert monotonic_prop(np.array([1.0, 1.0, nan])) == Monotonic.Falling(False)
    assert monotonic_prop(np.array([1.0, 2.0, nan])) == Monotonic.NotMonotonic
    assert monotonic_prop(np.array([3.0, 2.0, nan])) == Monotonic.NotMonotonic
