"""Oracle (and, on a GPU box, the CUDA path) against numbers produced by the REAL reference + Ceres
(oracle/ceres_ref/README.md).  Skipped while tests/golden/ceres/ holds no dumps: neither Ceres nor Eigen is installed in
this image, so the files cannot be produced here -- this is the hook that closes "parity unpinned" once they can."""
import os

import numpy as np
import pytest

from oracle import ba_oracle as bo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ceres")
CASES = ["cfg2", "cfg2_full", "cfg3", "cfg3_full"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_against_ceres_iteration_history(name):
    path = os.path.join(GOLD, name + ".bin")
    if not os.path.exists(path):
        pytest.skip("no Ceres dump for this case (Ceres / Eigen are not installed here); see oracle/ceres_ref/README.md")
    from tests.golden.export_windows import CASES as MAKE
    w, st, _ = MAKE[name]()
    raw = np.fromfile(path, dtype="<f8")
    per = 1 + 16 * w.N + w.M
    K = len(raw) // per
    for k in range(1, K + 1):
        rec = raw[(k - 1) * per:k * per]
        frames, rho = rec[1:1 + 16 * w.N].reshape(w.N, 16), rec[1 + 16 * w.N:]
        s, summ = bo.solve(w, st, max_iter=k)
        assert np.allclose(s.p, frames[:, 4:7], rtol=0, atol=1e-7), (name, k)
        assert np.allclose(s.rho, rho, rtol=1e-6, atol=0), (name, k)
