"""BatchedHeistEnv -- N independent Heist Architect environments stepped by sm_100a CUDA kernels.

Host side of the drop-in boundary: mirrors HeistEnvironment (heist_architect/environment.py:40-426)
for a batch, calling the C ABI in include/heist_b200.h.  All tensors are torch CUDA tensors; torch is
only the owner of device memory and streams here.  There is no CPU path.
"""
import ctypes as C
import math
import warnings
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _ffi

STATUS_NAMES = ("running", "detected", "vault_reached", "timeout", "already_done")


@dataclass
class EnvironmentConfig:
    """Same fields and defaults as the reference's EnvironmentConfig (environment.py:18-37)."""
    grid_rows: int = 20
    grid_cols: int = 20
    max_steps: int = 200
    start_pos: Tuple[int, int] = (1, 1)
    vault_pos: Tuple[int, int] = None
    architect_budget: int = 15
    reward_vault: float = 10.0
    reward_detection: float = -1.0
    reward_step: float = -0.01
    reward_architect_detect: float = 1.0
    reward_architect_invalid: float = -1.0

    def __post_init__(self):
        if self.vault_pos is None:
            self.vault_pos = (self.grid_rows - 2, self.grid_cols - 2)


class _DevView:
    """Zero-copy torch view of handle-owned device memory (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def guard_heading_table(path, speed):
    """heading after leaving waypoint i: degrees(atan2(-dr, dc)) % 360.0 (security.py:150-159), NaN if no move."""
    L = len(path)
    out = [float("nan")] * L
    if L < 2:
        return out
    for i in range(L):
        j = (i + speed) % L
        dr, dc = path[j][0] - path[i][0], path[j][1] - path[i][1]
        if dr != 0 or dc != 0:
            out[i] = math.degrees(math.atan2(-dr, dc)) % 360.0
    return out


class BatchedHeistEnv:
    def __init__(self, config: Optional[EnvironmentConfig] = None, num_envs: int = 1, device=None,
                 max_walls: int = 64, max_cams: int = 8, max_guards: int = 4, max_path: int = 8,
                 warn_uncached: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedHeistEnv needs a CUDA device (B200); there is no CPU fallback")
        self.config = config or EnvironmentConfig()
        cfg = self.config
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("BatchedHeistEnv: device must be a CUDA device")
        self.num_envs = int(num_envs)
        self.R, self.C = int(cfg.grid_rows), int(cfg.grid_cols)
        self.W = (self.C + 31) // 32
        self.max_walls, self.max_cams, self.max_guards, self.max_path = max_walls, max_cams, max_guards, max_path
        self._lib = _ffi.load()
        p = _ffi.HeistParams(
            grid_rows=self.R, grid_cols=self.C, max_steps=int(cfg.max_steps),
            start_row=int(cfg.start_pos[0]), start_col=int(cfg.start_pos[1]),
            vault_row=int(cfg.vault_pos[0]), vault_col=int(cfg.vault_pos[1]),
            architect_budget=int(cfg.architect_budget), max_walls=max_walls, max_cams=max_cams,
            max_guards=max_guards, max_path=max_path, reward_vault=float(cfg.reward_vault),
            reward_detection=float(cfg.reward_detection), reward_step=float(cfg.reward_step))
        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        torch.cuda.init()
        _ffi.check(self._lib.heist_create(C.byref(p), self.num_envs, dev_index, C.byref(h)), "heist_create")
        self._h = h
        # not an error, but a several-fold slowdown: say so (HEIST_REQUIRE_VIS_CACHE=1 makes heist_create fail instead)
        msg = self._lib.heist_last_warning()
        self.cache_warning = msg.decode() if msg else ""
        if self.cache_warning and warn_uncached:
            warnings.warn(self.cache_warning, RuntimeWarning, stacklevel=2)
        self._budget = None  # per-env int32 tensor or None (config default)
        self._make_views()

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.heist_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def close(self):
        self.__del__()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _make_views(self):
        v = _ffi.HeistStateView()
        _ffi.check(self._lib.heist_get_state(self._h, C.byref(v)), "heist_get_state")
        N, R, Cc, W = self.num_envs, self.R, self.C, self.W
        Kc, Kg, L = max(self.max_cams, 1), max(self.max_guards, 1), self.max_path

        def view(ptr, shape, typestr):
            return torch.as_tensor(_DevView(ptr, shape, typestr, self), device=self.device)

        self.tile_codes = view(v.tile, (N, R, Cc), "|u1")
        self.wall_bits = view(v.wall_bits, (N, R, W), "<i4")
        self.visibility_bits = view(v.vis_bits, (N, R, W), "<i4")
        self.env_static = view(v.env_static, (N, 4), "<i4")
        self.env_dyn = view(v.env_dyn, (N, 8), "<i4")
        self.cam_f = view(v.cam_f, (N, Kc, 2), "<f8")
        self.cam_i = view(v.cam_i, (N, Kc, 4), "<i2")
        self.cam_heading = view(v.cam_heading, (N, Kc), "<f8")
        self.guard_fov = view(v.guard_fov, (N, Kg), "<f8")
        self.guard_i = view(v.guard_i, (N, Kg, 4), "<i4")
        self.guard_path = view(v.guard_path, (N, Kg, L, 2), "|u1")
        self.guard_heading = view(v.guard_heading, (N, Kg), "<f8")
        self.guard_idx = view(v.guard_idx, (N, Kg), "<i4")
        self.wall_accepted = view(v.wall_accepted, (N, max(self.max_walls, 1)), "|u1")

    def _dev(self, x, dtype):
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=dtype)
        else:
            t = torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)
        return t.contiguous()

    MODE_DEFAULT, MODE_EXACT, MODE_MARCH, MODE_TABLES = 0, 1, 2, 3

    def set_mode(self, mode):
        """Verification knob (include/heist_b200.h): 0 = angular visibility cache + ray-march for what it does not
        cover, 1 = all-fp64 ray-march, 2 = filtered ray-march everywhere, 3 = tables only (the caller guarantees that the
        cache covers every layout -- Architect-decoded ones always are -- so a tick is exactly one kernel: CUDA graphs).
        All modes are bit-identical."""
        _ffi.check(self._lib.heist_set_mode(self._h, int(mode)), "heist_set_mode")

    def cache_stats(self):
        """(envs served by the angular visibility cache after the last set_layout, cache bytes on the device)."""
        import ctypes
        n, b = ctypes.c_int32(0), ctypes.c_int64(0)
        _ffi.check(self._lib.heist_cache_stats(self._h, ctypes.byref(n), ctypes.byref(b), self._stream()), "heist_cache_stats")
        return int(n.value), int(b.value)

    def launch_count(self):
        """Kernels launched by reset / step / step_many / step_observe on this handle so far."""
        import ctypes
        n = ctypes.c_int64(0)
        _ffi.check(self._lib.heist_launch_count(self._h, ctypes.byref(n)), "heist_launch_count")
        return int(n.value)

    def set_exact_only(self, flag):
        """Force the all-fp64 ray-march (True) or go back to the default mode (False)."""
        self.set_mode(self.MODE_EXACT if flag else self.MODE_DEFAULT)

    def check_errors(self):
        _ffi.check(self._lib.heist_check_errors(self._h, self._stream()), "heist_check_errors")

    # ------------------------------------------------------------------ layouts
    def scale_budget(self, budget):
        """env.budget.scale_budget (budget.py:64-67) for the batch: int or per-env array."""
        if budget is None:
            self._budget, self._budget_max = None, int(self.config.architect_budget)
        elif np.isscalar(budget):
            self._budget = torch.full((self.num_envs,), int(budget), dtype=torch.int32, device=self.device)
            self._budget_max = int(budget)
        else:
            self._budget = self._dev(budget, torch.int32)
            assert self._budget.shape == (self.num_envs,)
            self._budget_max = int(self._budget.max().item())

    def _check_decode_capacity(self):
        """The Architect decode buys at most budget // cost assets of a kind (networks.py:283-318); the per-env lists
        must hold them, or the layout would silently differ from the reference's."""
        b = getattr(self, "_budget_max", int(self.config.architect_budget))
        if b // 3 > self.max_cams or b // 5 > self.max_guards or b > self.max_walls:
            raise ValueError(f"budget {b} can buy more assets than the capacities hold (max_walls {self.max_walls}, "
                             f"max_cams {self.max_cams}, max_guards {self.max_guards}): create the env with larger capacities")

    def set_layout_from_asset_map(self, asset_map, cam_params, budget=None, allow_cameras=True, allow_guards=True):
        """Architect decode + curriculum filter + set_layout + BFS (networks.py:273-335, training.py:464-470).

        asset_map [N,R,C] int8 {0,1,2,3}; cam_params [N,3] float32 (fov, speed, heading) -> valid [N] bool."""
        if budget is not None:
            self.scale_budget(budget)
        self._check_decode_capacity()
        am = self._dev(asset_map, torch.int8)
        cp = self._dev(cam_params, torch.float32)
        assert am.shape == (self.num_envs, self.R, self.C) and cp.shape == (self.num_envs, 3)
        valid = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        _ffi.check(self._lib.heist_decode_validate(self._h, _ptr(am), _ptr(cp), _ptr(self._budget), int(allow_cameras),
                                                   int(allow_guards), _ptr(valid), self._stream()),
                   "heist_decode_validate")
        self._keep = (am, cp)
        return valid.bool()

    def pack_layouts(self, layouts):
        """Pack per-env (walls, cameras, guards) in the reference's set_layout format (environment.py:102-113)
        into the HeistLayoutArrays tensors (host numpy).  Capacity overflow raises."""
        N, Kw, Kc, Kg, L = self.num_envs, max(self.max_walls, 1), max(self.max_cams, 1), max(self.max_guards, 1), self.max_path
        assert len(layouts) == N
        a = {
            "n_walls": np.zeros(N, np.int32), "wall_rc": np.zeros((N, Kw, 2), np.int16),
            "n_cams": np.zeros(N, np.int32), "cam_rc": np.zeros((N, Kc, 2), np.int16),
            "cam_f": np.zeros((N, Kc, 3), np.float64), "cam_range": np.zeros((N, Kc), np.int32),
            "n_guards": np.zeros(N, np.int32), "guard_len": np.zeros((N, Kg), np.int32),
            "guard_path": np.zeros((N, Kg, L, 2), np.int16), "guard_head": np.full((N, Kg, L), np.nan, np.float64),
            "guard_speed": np.zeros((N, Kg), np.int32), "guard_range": np.zeros((N, Kg), np.int32),
            "guard_fov": np.zeros((N, Kg), np.float64),
        }
        for n, (walls, cams, guards) in enumerate(layouts):
            if len(walls) > self.max_walls or len(cams) > self.max_cams or len(guards) > self.max_guards:
                raise ValueError(f"env {n}: layout exceeds capacity (walls {len(walls)}/{self.max_walls}, cameras "
                                 f"{len(cams)}/{self.max_cams}, guards {len(guards)}/{self.max_guards})")
            a["n_walls"][n] = len(walls)
            for k, (r, c) in enumerate(walls):
                a["wall_rc"][n, k] = (max(-1, min(int(r), 32767)), max(-1, min(int(c), 32767)))
            a["n_cams"][n] = len(cams)
            for k, cd in enumerate(cams):
                a["cam_rc"][n, k] = (max(-1, min(int(cd["row"]), 32767)), max(-1, min(int(cd["col"]), 32767)))
                a["cam_f"][n, k] = (float(cd.get("fov_angle", 60.0)), float(cd.get("heading", 0.0)),
                                    float(cd.get("rotation_speed", 15.0)))
                a["cam_range"][n, k] = int(cd.get("vision_range", 6))
            a["n_guards"][n] = len(guards)
            for k, gd in enumerate(guards):
                path = [tuple(map(int, p)) for p in gd["patrol_path"]]
                if len(path) > L:
                    raise ValueError(f"env {n}: patrol path of {len(path)} waypoints exceeds max_path={L}")
                speed = int(gd.get("speed", 1))
                a["guard_len"][n, k] = len(path)
                for i, (r, c) in enumerate(path):
                    if not (0 <= r < self.R and 0 <= c < self.C):
                        raise ValueError(f"env {n}: guard waypoint {(r, c)} outside the grid")
                    a["guard_path"][n, k, i] = (r, c)
                a["guard_head"][n, k, :len(path)] = guard_heading_table(path, speed)
                a["guard_speed"][n, k] = speed
                a["guard_range"][n, k] = int(gd.get("vision_range", 4))
                a["guard_fov"][n, k] = float(gd.get("fov_angle", 90.0))
        return a

    def set_layout_explicit(self, layouts, budget=None):
        """HeistEnvironment.set_layout for every env; `layouts` is a list of (walls, cameras, guards) or the
        dict returned by pack_layouts -> valid [N] bool."""
        if budget is not None:
            self.scale_budget(budget)
        arrays = layouts if isinstance(layouts, dict) else self.pack_layouts(layouts)
        dev = {k: torch.as_tensor(v).to(self.device).contiguous() for k, v in arrays.items()}
        la = _ffi.HeistLayoutArrays(**{k: dev[k].data_ptr() for k in dev})
        valid = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        _ffi.check(self._lib.heist_set_layout_explicit(self._h, C.byref(la), _ptr(self._budget), _ptr(valid),
                                                       self._stream()), "heist_set_layout_explicit")
        self._keep = dev
        return valid.bool()

    # ------------------------------------------------------------------ dynamics
    def reset(self, mask=None):
        m = None if mask is None else self._dev(mask, torch.uint8)
        _ffi.check(self._lib.heist_reset(self._h, _ptr(m), self._stream()), "heist_reset")
        self._keep_mask = m

    def step(self, actions, want_reward64=False):
        """-> reward [N] float32, done [N] bool, status [N] uint8 (and reward64 [N] float64 if asked)."""
        a = self._dev(actions, torch.int8)
        assert a.shape == (self.num_envs,)
        N = self.num_envs
        reward = torch.empty(N, dtype=torch.float32, device=self.device)
        r64 = torch.empty(N, dtype=torch.float64, device=self.device) if want_reward64 else None
        done = torch.empty(N, dtype=torch.uint8, device=self.device)
        status = torch.empty(N, dtype=torch.uint8, device=self.device)
        _ffi.check(self._lib.heist_step(self._h, _ptr(a), _ptr(reward), _ptr(r64), _ptr(done), _ptr(status),
                                        self._stream()), "heist_step")
        self._keep_a = a
        if want_reward64:
            return reward, done.bool(), status, r64
        return reward, done.bool(), status

    def step_many(self, actions, autoreset=True, want_vis=False, out=None):
        """T steps in one launch on time-major actions [T,N] int8 -> dict(reward, done, status[, vis_bits])."""
        a = self._dev(actions, torch.int8)
        T, N = a.shape
        assert N == self.num_envs
        if out is None:
            out = {"reward": torch.empty((T, N), dtype=torch.float32, device=self.device),
                   "done": torch.empty((T, N), dtype=torch.uint8, device=self.device),
                   "status": torch.empty((T, N), dtype=torch.uint8, device=self.device)}
            if want_vis:
                out["vis_bits"] = torch.empty((T, N, self.R, self.W), dtype=torch.int32, device=self.device)
        _ffi.check(self._lib.heist_step_many(self._h, _ptr(a), T, int(autoreset), _ptr(out.get("reward")),
                                             _ptr(out.get("done")), _ptr(out.get("status")), _ptr(out.get("vis_bits")),
                                             self._stream()), "heist_step_many")
        self._keep_a = a
        return out

    def step_many_host(self, actions_host, out_host=None, autoreset=True, vis_out=None):
        """step_many for rollout buffers that live on the HOST: actions_host [T,N] int8 (pinned), outputs into pinned
        host tensors dict(reward f32, done u8, status u8) (allocated if not given); vis_out, if given, is a DEVICE
        tensor [T,N,R,W] int32.  The copies ride along the pipelined launch (include/heist_b200.h); the host tensors
        are valid after the current stream has been synchronised."""
        a = actions_host if isinstance(actions_host, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(actions_host))
        assert a.dtype == torch.int8 and a.device.type == "cpu" and a.is_contiguous()
        if not a.is_pinned():
            a = a.pin_memory()
        T, N = a.shape
        assert N == self.num_envs
        if out_host is None:
            out_host = {"reward": torch.empty((T, N), dtype=torch.float32).pin_memory(),
                        "done": torch.empty((T, N), dtype=torch.uint8).pin_memory(),
                        "status": torch.empty((T, N), dtype=torch.uint8).pin_memory()}
        for v in out_host.values():
            assert v.device.type == "cpu" and v.is_pinned() and v.is_contiguous()
        _ffi.check(self._lib.heist_step_many_host(self._h, _ptr(a), T, int(autoreset), _ptr(out_host.get("reward")),
                                                  _ptr(out_host.get("done")), _ptr(out_host.get("status")), _ptr(vis_out),
                                                  self._stream()), "heist_step_many_host")
        self._keep_h = (a, out_host, vis_out)
        return out_host

    def step_observe(self, actions, autoreset=True, state_out=None):
        """One tick + the dense state for the next policy forward (training.py:523-529 for the batch).
        -> reward [N] f32, done [N] bool, status [N] u8, state [N,3,R,C] f32 (after the auto-reset, if any)."""
        a = self._dev(actions, torch.int8)
        N = self.num_envs
        reward = torch.empty(N, dtype=torch.float32, device=self.device)
        done = torch.empty(N, dtype=torch.uint8, device=self.device)
        status = torch.empty(N, dtype=torch.uint8, device=self.device)
        if state_out is None:
            state_out = torch.empty((N, 3, self.R, self.C), dtype=torch.float32, device=self.device)
        _ffi.check(self._lib.heist_step_observe(self._h, _ptr(a), int(autoreset), _ptr(reward), _ptr(done), _ptr(status),
                                                _ptr(state_out), self._stream()), "heist_step_observe")
        return reward, done.bool(), status, state_out

    def expand_states(self, vis_bits, pos, env_idx, out=None):
        """Dense [M,3,R,C] states from packed transitions (visibility bitmap, row|col<<16, env index)."""
        vb = vis_bits.contiguous().view(-1, self.R, self.W)
        M = vb.shape[0]
        p = pos.contiguous().view(-1).to(torch.int32)
        e = env_idx.contiguous().view(-1).to(torch.int32)
        assert p.numel() == M and e.numel() == M and vb.dtype == torch.int32
        if out is None:
            out = torch.empty((M, 3, self.R, self.C), dtype=torch.float32, device=self.device)
        _ffi.check(self._lib.heist_expand_states(self._h, _ptr(vb), _ptr(p), _ptr(e), M, _ptr(out), self._stream()),
                   "heist_expand_states")
        return out

    # ------------------------------------------------------------------ observations
    def observe(self, out=None):
        """get_state_tensor for the batch: [N,3,R,C] float32 (environment.py:347-374)."""
        if out is None:
            out = torch.empty((self.num_envs, 3, self.R, self.C), dtype=torch.float32, device=self.device)
        _ffi.check(self._lib.heist_observe(self._h, _ptr(out), self._stream()), "heist_observe")
        return out

    def observation(self):
        """_get_observation for the batch (environment.py:305-345): the same five keys, batched."""
        st = self.observe()
        vec = torch.empty((self.num_envs, 5), dtype=torch.float32, device=self.device)
        _ffi.check(self._lib.heist_observation_vectors(self._h, _ptr(vec), self._stream()),
                   "heist_observation_vectors")
        return {"occupancy_grid": st[:, 0], "visibility_map": st[:, 1], "solver_position": vec[:, 0:2],
                "vault_direction": vec[:, 2:4], "time_feature": vec[:, 4:5]}

    def architect_reward(self):
        """calculate_architect_reward per env (rewards.py:43-73) -> (reward [N] f64, solve_rate [N] f64)."""
        rw = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        sr = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        _ffi.check(self._lib.heist_architect_reward(self._h, _ptr(rw), _ptr(sr), self._stream()),
                   "heist_architect_reward")
        return rw, sr

    # ------------------------------------------------------------------ state accessors
    @property
    def solver_pos(self):
        p = self.env_dyn[:, 0]
        return torch.stack([p & 0xFFFF, p >> 16], dim=1)

    @property
    def tick(self):
        return self.env_dyn[:, 1]

    @property
    def done(self):
        return (self.env_dyn[:, 4] & 1).bool()

    @property
    def solver_detected(self):
        return ((self.env_dyn[:, 4] >> 1) & 1).bool()

    @property
    def vault_reached(self):
        return ((self.env_dyn[:, 4] >> 2) & 1).bool()

    @property
    def valid(self):
        return self.env_static[:, 2].bool()

    @property
    def budget_spent(self):
        return self.env_static[:, 3]

    def visibility_dense(self):
        """[N,R,C] float32 0/1 from the packed bitmap (host-side convenience for tests / the facade)."""
        bits = self.visibility_bits
        cols = torch.arange(self.C, device=self.device)
        words = bits[:, :, cols // 32]
        return ((words >> (cols % 32)) & 1).to(torch.float32)
