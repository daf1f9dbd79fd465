"""ctypes front-end for the CPU oracle (oracle/heist_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity status: pinned against the Python reference through tests/golden/ (see
tests/golden/make_golden.py and tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libheist_oracle.so")

STATUS_NAMES = ["running", "detected", "vault_reached", "timeout", "already_done"]


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "heist_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.oenv_create.restype = vp
    L.oenv_create.argtypes = [i, i, i, i, i, i, i, i, d, d, d]
    L.oenv_free.argtypes = [vp]
    L.oenv_scale_budget.argtypes = [vp, i]
    L.oenv_is_valid.argtypes = [vp]
    L.oenv_is_valid.restype = i
    L.oenv_set_layout.restype = i
    L.oenv_set_layout.argtypes = [vp, i, vp, i, vp, vp, vp, i, vp, vp, i, vp, vp, vp]
    L.oenv_reset.argtypes = [vp]
    L.oenv_step.restype = i
    L.oenv_step.argtypes = [vp, i, C.POINTER(d), C.POINTER(i)]
    L.oenv_state_tensor.argtypes = [vp, vp]
    L.oenv_obs_vectors.argtypes = [vp, vp, vp, vp]
    L.oenv_get_grid.argtypes = [vp, vp]
    L.oenv_get_vis.argtypes = [vp, vp]
    L.oenv_get_info.argtypes = [vp, vp]
    L.oenv_get_cam_headings.argtypes = [vp, vp]
    L.oenv_get_guards.argtypes = [vp, vp, vp]
    L.oracle_bfs.restype = i
    L.oracle_bfs.argtypes = [vp, i, i, i, i, i, i]
    L.oracle_generate_patrol.argtypes = [i, i, i, i, vp]
    L.oracle_decode_layout.argtypes = [vp, i, i, i, vp, vp, vp, vp]
    L.oracle_gae.argtypes = [vp, vp, vp, i, i, d, d, vp, vp]
    L.oracle_normalize.argtypes = [vp, i, vp]
    L.oracle_architect_reward.restype = d
    L.oracle_architect_reward.argtypes = [i, d]
    L.oracle_num_threads.restype = i
    L.oracle_rollout.restype = C.c_long
    L.oracle_rollout.argtypes = [vp, i, vp, i, i, vp, vp, vp, vp, vp, i]
    L.oracle_reset_all.argtypes = [vp, i, i]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ia(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    return a if shape is None else a.reshape(shape)


class OracleEnv:
    """One environment; mirrors HeistEnvironment (heist_architect/environment.py:40-426)."""

    def __init__(self, rows=20, cols=20, max_steps=200, start=(1, 1), vault=None, budget=15,
                 reward_vault=10.0, reward_detection=-1.0, reward_step=-0.01):
        self.R, self.C, self.max_steps = rows, cols, max_steps
        self.start = tuple(start)
        self.vault = tuple(vault) if vault is not None else (rows - 2, cols - 2)
        self._L = lib()
        self._h = self._L.oenv_create(rows, cols, max_steps, self.start[0], self.start[1],
                                      self.vault[0], self.vault[1], budget,
                                      reward_vault, reward_detection, reward_step)

    def __del__(self):
        try:
            if self._h:
                self._L.oenv_free(self._h)
                self._h = None
        except Exception:
            pass

    def scale_budget(self, b):
        self._L.oenv_scale_budget(self._h, int(b))

    def set_layout(self, walls, cameras, guards):
        """Same argument shapes as HeistEnvironment.set_layout (environment.py:102-113)."""
        wall_rc = _ia(walls, (-1, 2)) if len(walls) else np.zeros((0, 2), np.int32)
        ncam = len(cameras)
        cam_rc = np.zeros((max(ncam, 1), 2), np.int32)
        cam_f = np.zeros((max(ncam, 1), 3), np.float64)
        cam_rng = np.zeros(max(ncam, 1), np.int32)
        for k, cd in enumerate(cameras):
            cam_rc[k] = (cd["row"], cd["col"])
            cam_f[k] = (cd.get("fov_angle", 60.0), cd.get("heading", 0.0), cd.get("rotation_speed", 15.0))
            cam_rng[k] = cd.get("vision_range", 6)
        ng = len(guards)
        stride = max([len(g["patrol_path"]) for g in guards] + [1])
        g_len = np.zeros(max(ng, 1), np.int32)
        g_path = np.zeros((max(ng, 1), stride, 2), np.int32)
        g_speed = np.zeros(max(ng, 1), np.int32)
        g_rng = np.zeros(max(ng, 1), np.int32)
        g_fov = np.zeros(max(ng, 1), np.float64)
        for k, gd in enumerate(guards):
            p = gd["patrol_path"]
            g_len[k] = len(p)
            if len(p):
                g_path[k, :len(p)] = np.asarray(p, np.int32)
            g_speed[k] = gd.get("speed", 1)
            g_rng[k] = gd.get("vision_range", 4)
            g_fov[k] = gd.get("fov_angle", 90.0)
        rc = self._L.oenv_set_layout(self._h, len(wall_rc), _p(wall_rc), ncam, _p(cam_rc), _p(cam_f),
                                     _p(cam_rng), ng, _p(g_len), _p(g_path), stride, _p(g_speed),
                                     _p(g_rng), _p(g_fov))
        if rc < 0:
            raise ValueError("guard waypoint outside the grid")
        return bool(rc)

    def is_level_valid(self):
        return bool(self._L.oenv_is_valid(self._h))

    def reset(self):
        self._L.oenv_reset(self._h)

    def step(self, action):
        r, dn = C.c_double(), C.c_int()
        st = self._L.oenv_step(self._h, int(action), C.byref(r), C.byref(dn))
        return r.value, bool(dn.value), st

    @property
    def grid(self):
        g = np.zeros((self.R, self.C), np.int32)
        self._L.oenv_get_grid(self._h, _p(g))
        return g

    @property
    def visibility(self):
        v = np.zeros((self.R, self.C), np.float32)
        self._L.oenv_get_vis(self._h, _p(v))
        return v

    def info(self):
        a = np.zeros(10, np.int32)
        self._L.oenv_get_info(self._h, _p(a))
        keys = ["solver_r", "solver_c", "tick", "done", "detected", "vault_reached", "n_walls", "n_cams",
                "n_guards", "spent"]
        return dict(zip(keys, a.tolist()))

    def cam_headings(self):
        n = self.info()["n_cams"]
        a = np.zeros(max(n, 1), np.float64)
        self._L.oenv_get_cam_headings(self._h, _p(a))
        return a[:n]

    def guards_state(self):
        n = self.info()["n_guards"]
        a = np.zeros((max(n, 1), 3), np.int32)
        h = np.zeros(max(n, 1), np.float64)
        self._L.oenv_get_guards(self._h, _p(a), _p(h))
        return a[:n], h[:n]

    def state_tensor(self):
        s = np.zeros((3, self.R, self.C), np.float32)
        self._L.oenv_state_tensor(self._h, _p(s))
        return s

    def obs_vectors(self):
        a, b, c = np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros(1, np.float32)
        self._L.oenv_obs_vectors(self._h, _p(a), _p(b), _p(c))
        return a, b, c


def pack_bits(vis):
    """(..., R, C) 0/1 array -> (..., R, W) uint32 row bitmaps, bit c%32 of word c//32 = column c."""
    vis = np.asarray(vis) > 0.5
    Cn = vis.shape[-1]
    W = (Cn + 31) // 32
    out = np.zeros(vis.shape[:-1] + (W,), np.uint32)
    for c in range(Cn):
        out[..., c // 32] |= vis[..., c].astype(np.uint32) << np.uint32(c % 32)
    return out


def bfs(grid, start, goal):
    g = np.ascontiguousarray(grid, dtype=np.int32)
    return bool(lib().oracle_bfs(_p(g), g.shape[0], g.shape[1], start[0], start[1], goal[0], goal[1]))


def decode_layout(asset_map, budget, fov, speed, heading):
    """networks.py:273-322 on one (H, W) asset map -> (walls, cameras, guards) like the reference."""
    am = np.ascontiguousarray(asset_map, dtype=np.int8)
    H, W = am.shape
    wall_rc = np.zeros((H * W, 2), np.int32)
    cam_rc = np.zeros((H * W, 2), np.int32)
    gpath = np.zeros((H * W, 8, 2), np.int32)
    counts = np.zeros(4, np.int32)
    lib().oracle_decode_layout(_p(am), H, W, int(budget), _p(wall_rc), _p(cam_rc), _p(gpath), _p(counts))
    walls = [tuple(x) for x in wall_rc[:counts[0]].tolist()]
    cams = [{"row": r, "col": c, "fov_angle": float(fov), "rotation_speed": float(speed),
             "heading": float(heading), "vision_range": 6} for r, c in cam_rc[:counts[1]].tolist()]
    guards = [{"patrol_path": [tuple(x) for x in gpath[k].tolist()], "speed": 1, "vision_range": 4,
               "fov_angle": 90.0} for k in range(counts[2])]
    return walls, cams, guards, int(counts[3])


def gae(rew, val, done, gamma=0.99, lam=0.95):
    """Column-wise GAE + returns on time-major (T,) or (T, N) float32 arrays (solver.py:141-143,228-244)."""
    rew = np.ascontiguousarray(rew, np.float32)
    val = np.ascontiguousarray(val, np.float32)
    dn = np.ascontiguousarray(done, np.float32)
    T = rew.shape[0]
    n = 1 if rew.ndim == 1 else rew.shape[1]
    adv, ret = np.zeros_like(rew), np.zeros_like(rew)
    lib().oracle_gae(_p(rew), _p(val), _p(dn), T, n, float(gamma), float(lam), _p(adv), _p(ret))
    return adv, ret


def normalize(adv):
    a = np.ascontiguousarray(adv, np.float32).ravel()
    out = np.zeros_like(a)
    lib().oracle_normalize(_p(a), a.size, _p(out))
    return out.reshape(np.shape(adv))


def architect_reward(valid, solve_rate):
    return lib().oracle_architect_reward(int(bool(valid)), float(solve_rate))


def num_threads():
    return lib().oracle_num_threads()


def rollout(envs, actions, autoreset=True, want_vis=False, n_threads=0):
    """Step every OracleEnv in `envs` through time-major int8 actions (T, n)."""
    acts = np.ascontiguousarray(actions, np.int8)
    T, n = acts.shape
    assert n == len(envs)
    hs = (C.c_void_p * n)(*[e._h for e in envs])
    rew32 = np.zeros((T, n), np.float32)
    rew64 = np.zeros((T, n), np.float64)
    done = np.zeros((T, n), np.uint8)
    status = np.zeros((T, n), np.uint8)
    vis = None
    if want_vis:
        R, Cn = envs[0].R, envs[0].C
        vis = np.zeros((T, n, R, (Cn + 31) // 32), np.uint32)
    live = lib().oracle_rollout(hs, n, _p(acts), T, int(autoreset), _p(rew32), _p(rew64), _p(done),
                                _p(status), _p(vis), int(n_threads))
    return {"reward": rew32, "reward64": rew64, "done": done, "status": status, "vis_bits": vis, "live": live}


def reset_all(envs, n_threads=0):
    n = len(envs)
    hs = (C.c_void_p * n)(*[e._h for e in envs])
    lib().oracle_reset_all(hs, n, int(n_threads))
