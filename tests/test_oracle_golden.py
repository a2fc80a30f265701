"""The oracle against its own frozen outputs (tests/golden/oracle_v1.npz, generator: tests/golden/make_golden.py).  The reference
ships no golden vectors and cannot run here; these fixtures pin the oracle -- the arbiter of every GPU parity test --
to the state that passed the reference's known-answer tests, so that it cannot drift silently."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_reproduces_its_golden_vectors():
    import make_golden

    ref = np.load(os.path.join(HERE, "golden", "oracle_v1.npz"))
    got = make_golden.compute()
    assert sorted(got) == sorted(ref.files)
    for k in ref.files:
        a, b = np.asarray(got[k]), ref[k]
        assert a.dtype == b.dtype and a.shape == b.shape, k
        assert np.array_equal(a, b), k  # same machine arithmetic, serial kernels: bit for bit
