"""The claim behind cl2_points_sweep_hint (include/cl2.h, DESIGN.md section 4 "Sweeps"), checked by brute force on the
CPU: with both grids anchored at the same (minX, minY), removing the points of the noise cells of a run (E, m) does
not change which points a run (e, m') with m' >= m keeps, provided E >= 3e - gcd(e, E)."""
import math

import numpy as np

from cloops2_b200.callCisLoops import _rows_carry_over


def noise_keep(xy, eps, minPts, ox, oy):
    """removeNoiseGrids (blockDBSCAN.py:101-122) with an explicit grid origin; keep flag per point."""
    gx, gy = (xy[:, 0] - ox) // eps, (xy[:, 1] - oy) // eps
    cells = {}
    for i, c in enumerate(zip(gx.tolist(), gy.tolist())):
        cells.setdefault(c, []).append(i)
    s3 = {(a, b): sum(len(cells.get((a + da, b + db), ())) for da in (-1, 0, 1) for db in (-1, 0, 1)) for a, b in cells}
    keep = np.zeros(len(xy), dtype=bool)
    for (a, b), rows in cells.items():
        if s3[(a, b)] >= minPts or any(s3.get((a + da, b + db), 0) >= minPts
                                       for da in (-1, 0, 1) for db in (-1, 0, 1) if (da or db)):
            keep[rows] = True
    return keep


def random_points(rng):
    n, span = int(rng.integers(50, 500)), int(rng.choice([2000, 6000, 20000]))
    x = rng.integers(0, span, n)
    y = x + rng.integers(0, span // 2, n)
    for _ in range(int(rng.integers(0, 5))):
        cx, cy = rng.integers(0, span, 2)
        m = int(rng.integers(3, 20))
        x = np.concatenate([x, (cx + rng.normal(0, 150, m)).astype(np.int64)])
        y = np.concatenate([y, (cy + rng.normal(0, 150, m)).astype(np.int64)])
    return np.stack([np.abs(x), np.abs(y)], axis=1)


def test_rows_of_removed_cells_never_matter_for_safe_pairs():
    rng = np.random.default_rng(1)
    safe_cases = 0
    for _ in range(300):
        xy = random_points(rng)
        ox, oy = int(xy[:, 0].min()), int(xy[:, 1].min())
        E = int(rng.choice([500, 1000, 700, 450, 999]))
        e = int(rng.choice([100, 200, 250, 300, 333, 500]))
        m = int(rng.integers(2, 8))
        m2 = m + int(rng.integers(0, 3))
        if not _rows_carry_over(E, m, e, m2):
            continue
        assert e < E and E >= 3 * e - math.gcd(e, E)
        safe_cases += 1
        kE = noise_keep(xy, E, m, ox, oy)
        full = noise_keep(xy, e, m2, ox, oy)
        sub = np.zeros(len(xy), dtype=bool)
        sub[kE] = noise_keep(xy[kE], e, m2, ox, oy)
        assert np.array_equal(full, sub), (E, e, m, m2)
    assert safe_cases > 100


def test_carry_over_condition():
    assert _rows_carry_over(1000, 5, 500, 5) and _rows_carry_over(500, 5, 200, 5) and _rows_carry_over(1000, 5, 200, 9)
    assert not _rows_carry_over(500, 5, 400, 5)        # 500 < 3*400 - 100
    assert not _rows_carry_over(450, 5, 333, 5)
    assert not _rows_carry_over(1000, 5, 200, 4)       # smaller minPts
    assert not _rows_carry_over(200, 5, 500, 5)        # larger eps
