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


@pytest.mark.parametrize("seed", range(12))
def test_guard_cone_pairs_cover_every_state_a_patrol_reaches(seed):
    """The guards' cones are tabulated only for the (waypoint, heading) pairs of k_build_cache's fixed point
    (oracle.vis_bands.guard_reach restates it).  Property: a guard driven by the reference's rules -- Guard.update
    (security.py:145-159: index advance with Python %, heading from the move unless it is no move) and reset
    (environment.py:205-208: back to waypoint 0, heading kept) at arbitrary ticks -- never leaves that set; and the set
    is small (a ring: about its length plus its headings)."""
    import math
    import struct
    import heist_b200
    from oracle.vis_bands import guard_reach
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(1, 13))
    if seed % 3 == 0:   # the Architect's patrol ring (networks.py:324-335), clamped near a border
        r0, c0 = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        path = [(max(1, r0 + dr - 1), max(1, c0 + dc - 1)) for dr, dc in ((0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0))]
    else:               # arbitrary paths with repeated waypoints (no-move steps)
        path = [(int(rng.integers(0, 4)), int(rng.integers(0, 4))) for _ in range(n)]
    speed = int(rng.choice([1, 1, 2, 3, 5, -1, -2, 0, len(path)]))
    vals, reach = guard_reach(heist_b200.guard_heading_table(path, speed), speed)
    slot = {struct.pack("<d", v): s for s, v in enumerate(vals)}
    idx, heading = 0, 0.0
    seen = set()
    for t in range(4000):
        if rng.random() < 0.05:
            idx = 0                                           # reset(): current_idx = 0, heading kept
        elif len(path) >= 2:                                  # Guard.update
            old = idx
            idx = (idx + speed) % len(path)
            dr, dc = path[idx][0] - path[old][0], path[idx][1] - path[old][1]
            if dr != 0 or dc != 0:
                heading = math.degrees(math.atan2(-dr, dc)) % 360.0
        s = slot[struct.pack("<d", heading)]                   # (a heading that is no slot would be a KeyError)
        assert s in reach[idx], (path, speed, t, idx, heading)
        seen.add((idx, s))
    total = sum(len(r) for r in reach)
    assert len(seen) <= total <= len(path) * len(vals)
    if seed % 3 == 0 and speed == 1:
        assert total <= len(path) + 2 * len(vals)             # a ring: ~11 of its 32 pairs
