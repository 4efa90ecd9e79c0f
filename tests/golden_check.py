"""How the tests hold results to the committed history fixtures (tests/golden/*_history.npz).

The fixtures were written in round 1 by the un-fused canonical arithmetic (every product rounded before it is added)
driven through the reference's own Material / _vmap / DataManager machinery (tests/golden/make_golden.py).  The
canonical arithmetic now places fused multiply-adds by hand (oracle/canon.py), so:

* the oracle evaluated with ``oracle.canon.unfused()`` must still reproduce the fixtures BIT FOR BIT -- they are
  today's pin of "the round-1 oracle", untouched;
* the fused arithmetic (oracle and CUDA kernels) must agree with them to the north star's tolerance: identical
  active sets / local iteration counts and stress, state, tangent within rtol 1e-10.
"""

import numpy as np

RTOL = 1e-10  # BASELINE.json north_star: "stress, Ct and state within rtol 1e-10"


def close(a, b, what=""):
    """rtol 1e-10 on every entry, with an absolute floor of 1e-10 x the field's magnitude (entries that are zero up to
    round-off in a tensor whose other entries are O(scale))."""
    a, b = np.asarray(a), np.asarray(b)
    scale = float(np.max(np.abs(b))) if b.size else 0.0
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=RTOL * scale, err_msg=what)


def same_active_set(out, ref, borderline=0.0):
    """Identical active sets and local iteration counts.  ``borderline``: fraction of points whose count may differ by
    ONE -- a local Newton whose residual lands within rounding of its tolerance stops one step earlier or later when the
    last bits of the arithmetic change (fused vs un-fused); only the Hosford solve, with its 4 residuals, meets such
    points (about 1 in 60 000)."""
    assert np.array_equal(out["flag"], ref["flag"]), "active sets differ"
    diff = np.abs(out["n_iter"].astype(int) - ref["n_iter"].astype(int))
    assert diff.max(initial=0) <= (1 if borderline else 0), "local iteration counts differ"
    assert (diff != 0).mean() <= borderline, "local iteration counts differ"
    assert np.array_equal(out["fail"], ref["fail"])
