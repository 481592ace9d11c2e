"""Pins oracle/lsap.c against the installed SciPy (the solver the reference calls, reference
src/matcher.py:136) and against the known answers in SURVEY.md §8(a.1)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import matcher_oracle as mo


def _same(cost):
    r0, c0 = linear_sum_assignment(cost)
    r1, c1 = mo.lsap(cost)
    assert r0.tolist() == r1.tolist() and c0.tolist() == c1.tolist()


def test_known_answers():
    assert mo.lsap(np.array([[4, 1, 3], [2, 0, 5], [3, 2, 2]], np.float32))[1].tolist() == [1, 0, 2]
    for shape in ((3, 3), (5, 3), (3, 5), (4, 2)):
        r, c = mo.lsap(np.zeros(shape, np.float32))
        k = min(shape)
        assert r.tolist() == list(range(k)) and c.tolist() == list(range(k))
    r, c = mo.lsap(np.array([[1, 1, 2], [1, 1, 2], [2, 2, 0], [1, 1, 2]], np.float32))
    assert r.tolist() == [0, 1, 2] and c.tolist() == [0, 1, 2]


@pytest.mark.parametrize("shape", [(576, 10), (576, 50), (576, 100), (100, 576), (37, 37), (1, 9), (9, 1)])
def test_random_float_vs_scipy(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for _ in range(8):
        _same(rng.standard_normal(shape).astype(np.float32))


@pytest.mark.parametrize("shape", [(40, 7), (7, 40), (25, 25), (576, 20)])
def test_tie_heavy_integer_vs_scipy(shape):
    rng = np.random.default_rng(7 + shape[0])
    for hi in (2, 3, 5):
        for _ in range(10):
            _same(rng.integers(0, hi, size=shape).astype(np.float32))


def test_empty():
    r, c = mo.lsap(np.zeros((576, 0), np.float32))
    assert r.size == 0 and c.size == 0
