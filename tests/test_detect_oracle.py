"""Pins oracle/detect_oracle.py (CPU): the Poisson-disk restatement against a brute-force statement of what the filter
is for, and the quirk of the reference's cell walk (poisson_disk_filter.h:92-107)."""
import numpy as np

from oracle import detect_oracle as D


def test_poisson_filter_equals_brute_force_when_cells_are_single():
    rng = np.random.default_rng(5)
    r = 25.0
    for trial in range(20):
        cand = rng.uniform(0, 400, size=(600, 2))
        f = D.PoissonDiskFilter(r)
        acc = []
        for p in cand:
            ok = all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 >= r * r for q in acc)
            assert f.insert(p) == ok          # accepted points are >= r apart, so no cell ever holds two of them
            if ok:
                acc.append(p)
        assert len(acc) > 50


def test_cell_walk_skips_the_first_cell_and_looks_one_past_the_end():
    f = D.PoissonDiskFilter(10.0)
    visited = []

    class Spy(dict):
        def get(self, k, d=None):
            visited.append(k)
            return d
    f.grid = Spy()
    f.test((100.0, 100.0))
    ix, iy = f.index((100.0, 100.0))
    assert (ix - 2, iy - 2) not in visited and (ix - 2, iy + 3) in visited and len(visited) == 25


def test_border_is_applied_after_the_filter():
    img = np.zeros((100, 200), dtype=np.uint8)
    corners = np.array([[10.0, 50.0], [22.0, 50.0], [100.0, 50.0]])       # the first one lies in the 20-px border
    out = D.detect_keypoints(img, [], 25.0, corners=corners)
    assert out.tolist() == [[100.0, 50.0]]                                # (22, 50) is blocked by the rejected (10, 50)
