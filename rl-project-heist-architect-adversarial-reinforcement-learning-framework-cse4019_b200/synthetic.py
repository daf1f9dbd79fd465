"""Seeded synthetic workloads (host-side numpy): sampled Architect asset maps, camera parameters and
action streams of the shapes BASELINE.json names.  Used by bench.py and the parity tests so that the
CUDA path and its CPU checker see identical inputs."""
import numpy as np

BASE_SEED = 20261018


def sample_asset_maps(rng, n, rows, cols, p_wall=0.03, p_cam=0.012, p_guard=0.006):
    """[n,R,C] int8 in {0,1,2,3}: i.i.d. per interior cell."""
    u = rng.random((n, rows, cols))
    am = np.zeros((n, rows, cols), np.int8)
    am[u < p_wall] = 1
    am[(u >= p_wall) & (u < p_wall + p_cam)] = 2
    am[(u >= p_wall + p_cam) & (u < p_wall + p_cam + p_guard)] = 3
    am[:, 0, :] = 0
    am[:, -1, :] = 0
    am[:, :, 0] = 0
    am[:, :, -1] = 0
    return am


def sample_asset_maps_exact(rng, n, rows, cols, n_walls, n_cams, n_guards):
    """[n,R,C] int8 with exactly the given asset counts on distinct random interior cells (config 3)."""
    am = np.zeros((n, rows, cols), np.int8)
    interior = (rows - 2) * (cols - 2)
    k = n_walls + n_cams + n_guards
    codes = np.array([1] * n_walls + [2] * n_cams + [3] * n_guards, np.int8)
    for i in range(n):
        cells = rng.choice(interior, size=k, replace=False)
        r, c = cells // (cols - 2) + 1, cells % (cols - 2) + 1
        am[i, r, c] = codes
    return am


def sample_cam_params(rng, n, nice=False):
    """[n,3] float32 (fov, speed, heading) in the ranges of the Architect heads (networks.py:232-236)."""
    if nice:  # the tie-prone defaults people type: fov 60, speed 15, heading multiple of 15
        return np.stack([np.full(n, 60.0), np.full(n, 15.0), rng.integers(0, 24, n) * 15.0], 1).astype(np.float32)
    return np.stack([rng.uniform(30, 120, n), rng.uniform(5, 35, n), rng.uniform(0, 360, n)], 1).astype(np.float32)


def sample_actions(rng, T, n):
    """[T,n] int8 uniform over the 5 Solver actions."""
    return rng.integers(0, 5, size=(T, n), dtype=np.int8)


def make_valid_workload(env, seed, budget, exact_counts=None, nice=False, max_rounds=64):
    """Sample asset maps until every env's layout is BFS-valid ("random valid layouts"): validity comes
    from the device decode+BFS; invalid envs are resampled.  Returns host (asset_map, cam_params)."""
    rng = np.random.default_rng(seed)
    n, R, C = env.num_envs, env.R, env.C

    def draw(m):
        if exact_counts is not None:
            return sample_asset_maps_exact(rng, m, R, C, *exact_counts)
        return sample_asset_maps(rng, m, R, C)

    am = draw(n)
    cp = sample_cam_params(rng, n, nice)
    for _ in range(max_rounds):
        valid = env.set_layout_from_asset_map(am, cp, budget).cpu().numpy()
        bad = np.nonzero(~valid)[0]
        if bad.size == 0:
            return am, cp
        am[bad] = draw(bad.size)
    raise RuntimeError("could not sample valid layouts")
