"""CPU check of the claim the angular visibility cache rests on, against the reference-pinned C
oracle: outside the tie bands, the tiles one camera ray marks do not depend on where in the gap its angle lies."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import heist_oracle as ho  # noqa: E402
from oracle.vis_bands import tie_bands  # noqa: E402


def ray_tiles(rows, cols, walls, row, col, vision_range, angle):
    """Tiles one ray at `angle` marks, through the oracle: a camera with fov 0 casts all its rays at its heading."""
    e = ho.OracleEnv(rows, cols, budget=1000)
    e.set_layout(walls, [{"row": row, "col": col, "fov_angle": 0.0, "heading": float(angle), "rotation_speed": 0.0,
                          "vision_range": vision_range}], [])
    e.reset()
    return ho.pack_bits(e.visibility).tobytes()


@pytest.mark.parametrize("rows,cols,row,col,vision_range,seed", [(20, 20, 10, 10, 6, 1), (20, 20, 2, 17, 6, 2),
                                                                  (33, 47, 16, 5, 7, 3), (12, 12, 6, 6, 3, 4)])
def test_rays_inside_a_gap_mark_the_same_tiles(rows, cols, row, col, vision_range, seed):
    rng = np.random.default_rng(seed)
    walls = [(int(rng.integers(1, rows - 1)), int(rng.integers(1, cols - 1))) for _ in range(25)]
    walls = [w for w in walls if w != (row, col)]
    bands = tie_bands(vision_range, 90.0)
    assert 50 < len(bands) <= 512
    n_gaps, distinct = 0, set()
    for (_, e0), (s1, _) in zip(bands[:-1], bands[1:]):
        assert s1 > e0
        ref = ray_tiles(rows, cols, walls, row, col, vision_range, 0.5 * (e0 + s1))
        # the gap's ends (just inside) and random interior angles, as fp64 values the reference would be handed
        for a in [np.nextafter(e0, s1), np.nextafter(s1, e0)] + list(rng.uniform(e0, s1, 3)):
            assert ray_tiles(rows, cols, walls, row, col, vision_range, a) == ref, (e0, s1, a)
        n_gaps += 1
        distinct.add(ref)
    assert n_gaps == len(bands) - 1 and len(distinct) > 10


def test_bands_are_thin():
    """Sanity of the construction itself: the bands are tiny slivers (a generic ray never falls into one)."""
    bands = tie_bands(6, 120.0)
    widths = np.array([e - s for s, e in bands[1:-1]])
    assert widths.max() < 1e-3 and np.median(widths) < 1e-8     # the widest are the tangent crossings at d = k + 0.5
    assert widths.sum() < 0.01                                  # of a 480-degree domain
