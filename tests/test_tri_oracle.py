"""CPU known-answer check of oracle/tri_oracle.py: noise-free projections triangulate back exactly; a point
behind a camera or farther than 100 depth units is rejected (stereo.h:112-117)."""
import numpy as np

from oracle import lie, tri_oracle


def _cams(rng, n):
    qs = [lie.expmap(rng.normal(0, 0.05, 3)) for _ in range(n)]
    ps = [np.array([0.3 * i, 0.02 * i, 0.0]) + rng.normal(0, 0.01, 3) for i in range(n)]
    return [tri_oracle.projection_matrix(q, p) for q, p in zip(qs, ps)]


def test_exact_recovery_and_rejection():
    rng = np.random.default_rng(0)
    Ps = _cams(rng, 5)
    X = np.array([0.4, -0.3, 6.0])
    zs = [(P @ np.r_[X, 1.0])[:2] / (P @ np.r_[X, 1.0])[2] for P in Ps]
    ok, p, score = tri_oracle.triangulate_scored(Ps, zs)
    assert ok and np.allclose(p, X, atol=1e-9) and score < 1e-20
    far = np.array([0.0, 0.0, 500.0])
    zs = [(P @ np.r_[far, 1.0])[:2] / (P @ np.r_[far, 1.0])[2] for P in Ps]
    ok, p, _ = tri_oracle.triangulate_scored(Ps, zs)
    assert not ok and abs(np.linalg.norm(p) - 1.0) < 1e-12
