"""The canonical exp and the synthetic-input recipe (CPU only)."""
import numpy as np

from oracle import canon, synth


def test_exp_c_accuracy():
    x = np.concatenate([np.linspace(-700, 700, 400001), -np.logspace(-14, 2.8, 50000), [0.0, -0.0]])
    y, ref = canon.exp_c(x), np.exp(x)
    ulp = np.abs(y - ref) / np.spacing(ref)
    assert ulp.max() <= 1.5


def test_exp_c_special_values():
    assert canon.exp_c(0.0) == 1.0
    assert canon.exp_c(-800.0) == 0.0
    assert canon.exp_c(800.0) == np.inf
    assert np.isnan(canon.exp_c(np.nan))
    assert canon.exp_c(-np.inf) == 0.0


def test_lame_matches_reference_formula():
    lam, mu = canon.lame(70e3, 0.3)
    assert lam == 70e3 * 0.3 / (1 + 0.3) / (1 - 2 * 0.3)  # python_materials/elasticity.py:12-13
    assert mu == 70e3 / 2 / (1 + 0.3)


def test_synth_is_counter_based():
    a = synth.strain(1000, 7, 1e-2, 3, 4)
    b = synth.strain(400, 7, 1e-2, 3, 4, start=600)
    assert np.array_equal(a[600:], b)  # shards reproduce the global stream
    u = synth.uniform(7, np.arange(200000, dtype=np.uint64), 3)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    F = synth.defgrad(10, 1, 3e-2, 0, 4)
    assert np.array_equal(F, np.tile([1, 1, 1, 0, 0, 0, 0, 0, 0.0], (10, 1)))
