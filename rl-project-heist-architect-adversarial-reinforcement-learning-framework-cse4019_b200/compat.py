"""Single-env facade with the reference's HeistEnvironment surface (environment.py:40-426) over the
batched CUDA engine (N = 1), so AdversarialTrainer / main.py demo code can use it unchanged:

    import heist_architect.training as tr
    tr.HeistEnvironment = heist_b200.HeistEnvironment      # swap-in point, training.py:28,152

Everything that is computed (layout validity, visibility, movement, rewards, observations) comes
from the kernels; this class only mirrors results into the Python attributes the reference exposes.
"""
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .batched_env import STATUS_NAMES, BatchedHeistEnv, EnvironmentConfig


class _Budget:
    """BudgetManager surface used by the trainer (budget.py:21-67): scale_budget, spent, remaining."""

    def __init__(self, env, total):
        self._env = env
        self.total_budget = int(total)
        self._scaled = False   # scale_budget zeroes `spent` until the next set_layout (budget.py:64-67)

    @property
    def spent(self):
        return 0 if self._scaled else int(self._env._b.budget_spent[0].item())

    @property
    def remaining(self):
        return self.total_budget - self.spent

    def scale_budget(self, new_budget):
        self.total_budget = int(new_budget)
        self._scaled = True
        self._env._b.scale_budget(int(new_budget))

    def reset(self):
        pass


class _Visibility:
    """DynamicVisibilityMap surface (visibility.py:12-69): .visibility, .is_visible()."""

    def __init__(self, env):
        self._env = env
        self.rows, self.cols = env.config.grid_rows, env.config.grid_cols

    @property
    def visibility(self):
        return self._env._b.visibility_dense()[0].cpu().numpy()

    def is_visible(self, row, col):
        return bool(self.visibility[row, col] > 0.5)


class HeistEnvironment:
    ACTIONS = {0: (0, 0), 1: (-1, 0), 2: (1, 0), 3: (0, -1), 4: (0, 1)}
    ACTION_NAMES = {0: "WAIT", 1: "UP", 2: "DOWN", 3: "LEFT", 4: "RIGHT"}
    NUM_SOLVER_ACTIONS = 5

    def __init__(self, config: Optional[EnvironmentConfig] = None, device=None, max_walls=512, max_cams=24,
                 max_guards=8, max_path=64):
        src = config or EnvironmentConfig()
        # accept the reference's own dataclass instance as well
        self.config = src if isinstance(src, EnvironmentConfig) else EnvironmentConfig(
            **{k: getattr(src, k) for k in EnvironmentConfig.__dataclass_fields__ if hasattr(src, k)})
        self._b = BatchedHeistEnv(self.config, 1, device, max_walls=max_walls, max_cams=max_cams,
                                  max_guards=max_guards, max_path=max_path, warn_uncached=False)
        self.budget = _Budget(self, self.config.architect_budget)
        self.visibility_map = _Visibility(self)
        self.walls: List[Any] = []
        self.solver_path: List[Tuple[int, int]] = [tuple(self.config.start_pos)]
        self.detection_events: List[Dict] = []
        self._layout_guards: List[Dict] = []

    # ------------------------------------------------------------------ layout
    def set_layout(self, walls, cameras, guards) -> bool:
        guards = [g for g in guards]
        for g in guards:
            if len(g["patrol_path"]) > self._b.max_path:
                raise ValueError("patrol path longer than the facade's max_path")
        walls = [(int(r), int(c)) for r, c in walls]
        valid = self._b.set_layout_explicit([(walls, cameras, guards)])
        self._b.check_errors()
        self.budget._scaled = False
        # HeistEnvironment.walls (environment.py:119-121): the walls of the request that were placed and paid for, in
        # request order -- the kernel reports which (a wall a guard later overwrites stays listed, as in the reference)
        ok = self._b.wall_accepted[0, :len(walls)].cpu().numpy()
        self.walls = [SimpleNamespace(row=r, col=c) for (r, c), a in zip(walls, ok) if a]
        return bool(valid[0].item())

    def is_level_valid(self) -> bool:
        return bool(self._b.valid[0].item())

    # ------------------------------------------------------------------ state mirrors
    @property
    def grid(self):
        return self._b.tile_codes[0].cpu().numpy().astype(np.int32)

    @property
    def cameras(self):
        n = int(self._b.env_static[0, 0].item())
        f = self._b.cam_f[0, :n].cpu().numpy()
        i = self._b.cam_i[0, :n].cpu().numpy()
        h = self._b.cam_heading[0, :n].cpu().numpy()
        return [SimpleNamespace(row=int(i[k, 0]), col=int(i[k, 1]), fov_angle=float(f[k, 0]), heading=float(h[k]),
                                rotation_speed=float(f[k, 1]), vision_range=int(i[k, 2])) for k in range(n)]

    @property
    def guards(self):
        n = int(self._b.env_static[0, 1].item())
        gi = self._b.guard_i[0, :n].cpu().numpy()
        gp = self._b.guard_path[0, :n].cpu().numpy()
        gh = self._b.guard_heading[0, :n].cpu().numpy()
        gx = self._b.guard_idx[0, :n].cpu().numpy()
        gf = self._b.guard_fov[0, :n].cpu().numpy()
        out = []
        for k in range(n):
            path = [tuple(map(int, p)) for p in gp[k, :gi[k, 0]]]
            idx = int(gx[k])
            out.append(SimpleNamespace(patrol_path=path, speed=int(gi[k, 1]), current_idx=idx,
                                       vision_range=int(gi[k, 2]), fov_angle=float(gf[k]), heading=float(gh[k]),
                                       row=path[idx][0], col=path[idx][1], position=path[idx]))
        return out

    @property
    def solver_pos(self):
        p = self._b.solver_pos[0].tolist()
        return (int(p[0]), int(p[1]))

    @property
    def tick(self):
        return int(self._b.tick[0].item())

    @property
    def done(self):
        return bool(self._b.done[0].item())

    @property
    def solver_detected(self):
        return bool(self._b.solver_detected[0].item())

    @property
    def vault_reached(self):
        return bool(self._b.vault_reached[0].item())

    # ------------------------------------------------------------------ solver phase
    def reset(self):
        self._b.reset()
        self.solver_path = [tuple(self.config.start_pos)]
        self.detection_events = []
        return self._get_observation()

    def step(self, action: int):
        if action not in self.ACTIONS:
            raise KeyError(action)
        tick_before = self.tick
        a = torch.tensor([int(action)], dtype=torch.int8)
        _, done, status, r64 = self._b.step(a, want_reward64=True)
        st = int(status[0].item())
        if st == 4:
            return self._get_observation(), 0.0, True, {"status": "already_done"}
        pos = self.solver_pos
        self.solver_path.append(pos)
        if self.solver_detected:  # a detected env is done, so the flag can only have been set by this step
            self.detection_events.append({"tick": tick_before, "position": pos})
        info = {"status": STATUS_NAMES[st], "tick": tick_before}
        return self._get_observation(), float(r64[0].item()), bool(done[0].item()), info

    # ------------------------------------------------------------------ observations
    def _get_observation(self) -> Dict[str, np.ndarray]:
        o = self._b.observation()
        return {k: v[0].cpu().numpy() for k, v in o.items()}

    def get_state_tensor(self) -> np.ndarray:
        return self._b.observe()[0].cpu().numpy()

    def get_architect_reward(self) -> float:
        if not self.is_level_valid():
            return self.config.reward_architect_invalid
        if self.solver_detected:
            return self.config.reward_architect_detect
        return 0.0

    def get_environment_state(self) -> Dict[str, Any]:
        return {
            "grid": self.grid.tolist(),
            "visibility": self.visibility_map.visibility.tolist(),
            "solver_pos": self.solver_pos, "solver_path": self.solver_path,
            "vault_pos": self.config.vault_pos, "start_pos": self.config.start_pos,
            "tick": self.tick, "done": self.done,
            "cameras": [{"row": c.row, "col": c.col, "heading": c.heading, "fov_angle": c.fov_angle,
                         "vision_range": c.vision_range} for c in self.cameras],
            "guards": [{"row": g.row, "col": g.col, "heading": g.heading, "patrol_path": g.patrol_path,
                        "current_idx": g.current_idx} for g in self.guards],
            "detection_events": self.detection_events,
        }

    def render_text(self) -> str:
        sym = {0: ".", 1: "#", 2: "S", 3: "V", 4: "C", 5: "G"}
        g, pos = self.grid, self.solver_pos
        return "\n".join("".join("@" if (r, c) == pos else sym.get(int(g[r, c]), "?") for c in range(g.shape[1]))
                         for r in range(g.shape[0]))

    def __repr__(self):
        return (f"HeistEnvironment(grid={self.config.grid_rows}x{self.config.grid_cols}, cameras={len(self.cameras)}, "
                f"guards={len(self.guards)}, walls={len(self.walls)}, tick={self.tick})")
