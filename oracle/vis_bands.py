"""Test infrastructure: numpy restatement of the tie-band construction of csrc/heist_cache.cuh (k_build_cache
step 1-2), used by tests/test_cache_bands_cpu.py to check the cache's claim against the C oracle on the CPU:
between two bands, every ray of a camera marks the same tiles.  Not used by the product path.

Reference semantics being tabulated: Camera.get_vision_cone_tiles (security.py:53-101): sample k of a ray at
angle a lands on tile (round(row - sin(a) * d), round(col + cos(a) * d)), d = 0.5 * k, k = 1 .. 2 * vision_range."""
import numpy as np

MU2 = 8e-12    # VC_MU2
PAD = 1e-9     # VC_PAD


def tie_bands(vision_range, fov):
    """Merged (start, end) angle bands in degrees that enclose every rounding-tie crossing of every sample, over the
    domain [-fov/2, 360 + fov/2] a ray angle can take; the first and last band are the outside of the domain."""
    dom_lo, dom_hi = -0.5 * fov - 1e-6, 360.0 + 0.5 * fov + 1e-6
    raw = [(-1e300, dom_lo), (dom_hi, 1e300)]
    for j in range(1, 2 * vision_range + 1):
        d = 0.5 * j
        for axis in (0, 1):
            for m in range(-8, 8):
                tie = m + 0.5
                if abs(tie) > d + MU2:
                    continue
                t, mu = tie / d, MU2 / d
                c_lo, c_hi = max(-1.0, t - mu), min(1.0, t + mu)
                if axis:                      # dy = -sin(a) = -cos(a - 90)
                    c_lo, c_hi = -c_hi, -c_lo
                a_lo, a_hi = np.degrees(np.arccos(c_hi)), np.degrees(np.arccos(c_lo))
                off = 90.0 if axis else 0.0
                for b0, b1 in ((off + a_lo, off + a_hi), (off - a_hi, off - a_lo)):
                    for n in (-1, 0, 1, 2):
                        s, e = b0 + 360.0 * n - PAD, b1 + 360.0 * n + PAD
                        if e < dom_lo or s > dom_hi:
                            continue
                        raw.append((s, e))
    raw.sort()
    merged = [list(raw[0])]
    for s, e in raw[1:]:
        if s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([s, e])
    return [tuple(b) for b in merged]


def guard_reach(heading_table, speed):
    """Restatement of k_build_cache's reachable-pair fixed point (csrc/heist_cache.cuh): which heading slots a guard
    can carry on each waypoint.  heading_table[i] = heading after leaving waypoint i (NaN: the move is no move).
    Slot 0 is the default heading 0.0, further slots the distinct table headings in order of appearance (compared by
    bit pattern, like the kernel).  Returns (slot_values, reach) with reach[i] = set of slots on waypoint i."""
    import math
    import struct
    bits = lambda x: struct.pack("<d", x)
    n = len(heading_table)
    vals = [0.0]
    hslot = []
    for h in heading_table:
        if math.isnan(h):
            hslot.append(255)
            continue
        for s, v in enumerate(vals):
            if bits(v) == bits(h):
                break
        else:
            vals.append(h)
            s = len(vals) - 1
        hslot.append(s)
    stride = speed % n if n >= 2 else 0          # Python int %, as py_imod
    reach = [set() for _ in range(n)]
    reach[0] = set(range(len(vals)))             # a reset puts the guard on waypoint 0 with whatever heading it carries
    changed = True
    while changed:
        changed = False
        for k in range(n):
            if not reach[k]:
                continue
            nxt = (k + stride) % n
            add = reach[k] if hslot[k] == 255 else {hslot[k]}
            if not add <= reach[nxt]:
                reach[nxt] |= add
                changed = True
    return vals, reach
