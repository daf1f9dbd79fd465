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
