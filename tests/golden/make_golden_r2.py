"""Generate tests/golden/heist_golden_r2.npz by running the UNMODIFIED Python reference (second fixture file).

Run in the build container only (the reference tree is not present on the GPU box):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_r2.py

Three groups, all outputs of reference code:

* tie*      - cameras whose headings ACCUMULATE to within a few ulp of multiples of 30 degrees without being
              equal to them (rotation_speed 0.1 and 1/3 from heading 0: 0.1 * 300 = 30.000000000000156), fov
              60 / 90 / 120, walls on and around the tiles the tie rays decide, >= 400 ticks.  These are the
              rays whose tile depends on the last bit of the platform's cos/sin (security.py:69-75).
* big*      - 32x32 and 64x64 layouts at the top curriculum budget (4 cameras + 2 guards; 7 cameras fov 120).
* trainer*  - the exact call sequence AdversarialTrainer._run_one_episode (training.py:418-600) makes on its
              HeistEnvironment -- budget.scale_budget, set_layout, then per attempt reset / get_state_tensor /
              step ... / tick, is_level_valid (rewards.py:58), get_environment_state -- recorded from the
              unmodified trainer driving the unmodified env, with every return value.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (puts the reference on sys.path, imports it)

import torch  # noqa: E402
import heist_architect.training as tr  # noqa: E402
from heist_architect.environment import HeistEnvironment, EnvironmentConfig  # noqa: E402

OUT = os.path.join(HERE, "heist_golden_r2.npz")


def tie_layout(rng, fov, speed, variant):
    """Camera mid-grid whose rays sweep through the multiples of 30 degrees; walls sprinkled over the tiles within
    range (the tie rays decide between neighbouring tiles there), one slow guard for good measure."""
    r0, c0 = 10, 9 + variant
    walls = set()
    while len(walls) < 9:
        dr, dc = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
        if (dr, dc) != (0, 0) and max(abs(dr), abs(dc)) >= 2:
            walls.add((r0 + dr, c0 + dc))
    cams = [{"row": r0, "col": c0, "fov_angle": float(fov), "heading": 0.0, "rotation_speed": speed, "vision_range": 6}]
    if variant == 1:   # a second camera starting on a multiple of 30, negative non-dyadic speed
        cams.append({"row": 4, "col": 14, "fov_angle": float(fov), "heading": 90.0, "rotation_speed": -0.3, "vision_range": 6})
    guards = [{"patrol_path": mg.patrol(15, 4, 20, 20), "speed": 1, "vision_range": 4, "fov_angle": 90.0}]
    return sorted(walls), cams, guards, 22


def big_layout(rng, R, kind):
    walls = [(int(rng.integers(1, R - 1)), int(rng.integers(1, R - 1))) for _ in range(2 if kind == "4c2g" else 1)]
    cams = []
    for _ in range(4 if kind == "4c2g" else 7):
        fov = float(np.float32(rng.uniform(30, 120))) if kind == "4c2g" else 120.0
        cams.append({"row": int(rng.integers(1, R - 1)), "col": int(rng.integers(1, R - 1)), "fov_angle": fov,
                     "heading": float(np.float32(rng.uniform(0, 360))), "rotation_speed": float(np.float32(rng.uniform(5, 35))),
                     "vision_range": 6})
    guards = []
    if kind == "4c2g":
        for _ in range(2):
            guards.append({"patrol_path": mg.patrol(int(rng.integers(1, R - 1)), int(rng.integers(1, R - 1)), R, R),
                           "speed": 1, "vision_range": 4, "fov_angle": 90.0})
    return walls, cams, guards, 22


class Recorder(HeistEnvironment):
    """The reference env, unmodified, with a tape of every call the trainer makes and what came back."""
    tape = None   # shared list, set by the generator
    depth = 0     # only calls made from outside the env are taped (set_layout calls is_level_valid itself)

    @classmethod
    def _log(cls, rec):
        if cls.depth == 0:
            cls.tape.append(rec)

    class _BudgetTap:
        def __init__(self, env, inner):
            self._env, self._inner = env, inner

        def scale_budget(self, b):
            self._inner.scale_budget(b)
            Recorder._log({"call": "scale_budget", "arg": int(b), "spent": int(self._inner.spent),
                                  "remaining": int(self._inner.remaining)})

        def __getattr__(self, k):
            return getattr(self._inner, k)

    def __init__(self, config=None):
        super().__init__(config)
        self.budget = Recorder._BudgetTap(self, self.budget)

    def _obs_rec(self, obs):
        return {"vec": np.concatenate([obs["solver_position"], obs["vault_direction"], obs["time_feature"]]).astype(np.float32),
                "occ": obs["occupancy_grid"].astype(np.float32), "vis": obs["visibility_map"].astype(np.float32)}

    def set_layout(self, walls, cameras, guards):
        Recorder.depth += 1
        try:
            v = super().set_layout(walls, cameras, guards)
        finally:
            Recorder.depth -= 1
        Recorder._log({"call": "set_layout", "walls": [list(map(int, w)) for w in walls],
                              "cameras": [{k: (float(x) if isinstance(x, float) else int(x)) for k, x in c.items()} for c in cameras],
                              "guards": [{**g, "patrol_path": [list(map(int, p)) for p in g["patrol_path"]]} for g in guards],
                              "ret": bool(v), "spent": int(self.budget.spent),
                              "n_placed": [len(self.walls), len(self.cameras), len(self.guards)],
                              "walls_placed": [[w.row, w.col] for w in self.walls],
                              "repr": repr(self)})
        return v

    def is_level_valid(self):
        v = super().is_level_valid()
        Recorder._log({"call": "is_level_valid", "ret": bool(v)})
        return v

    def reset(self):
        Recorder.depth += 1
        try:
            obs = super().reset()
        finally:
            Recorder.depth -= 1
        Recorder._log({"call": "reset", "_obs": self._obs_rec(obs)})
        return obs

    def step(self, action):
        Recorder.depth += 1
        try:
            obs, r, d, info = super().step(action)
        finally:
            Recorder.depth -= 1
        Recorder._log({"call": "step", "arg": int(action), "reward": float(r), "done": bool(d),
                              "info": {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in info.items()},
                              "tick_after": int(self.tick), "_obs": self._obs_rec(obs)})
        return obs, r, d, info

    def get_state_tensor(self):
        s = super().get_state_tensor()
        Recorder._log({"call": "get_state_tensor", "_state": np.asarray(s, np.float32)})
        return s

    def get_environment_state(self):
        st = super().get_environment_state()
        Recorder._log({"call": "get_environment_state", "ret": json.loads(json.dumps(st, default=lambda o: o.tolist()))})
        return st


def record_trainer_episodes(store, meta):
    tr.HeistEnvironment = Recorder   # the swap-in point the facade uses too (training.py:28,152)
    torch.manual_seed(20261018)
    np.random.seed(20261018)
    tmp = tempfile.mkdtemp(prefix="heist_golden_")
    cfg = EnvironmentConfig(grid_rows=20, grid_cols=20, max_steps=40)
    Recorder.tape = []
    trainer = tr.AdversarialTrainer(config=cfg, solver_episodes_per_layout=3, total_episodes=500,
                                    save_dir=os.path.join(tmp, "ckpt"), log_dir=os.path.join(tmp, "logs"))
    runs = []
    for ep in range(240, 262):
        Recorder.tape = []
        metrics, entry = trainer._run_one_episode(ep)
        tape = Recorder.tape
        valid = bool(entry.data["level_valid"])
        has_assets = any(c["call"] == "set_layout" and (c["cameras"] or c["guards"]) for c in tape)
        runs.append((ep, valid, has_assets, tape, metrics, entry))
        print(f"trainer episode {ep}: valid={valid} assets={has_assets} calls={len(tape)}", flush=True)
    # keep the first invalid layout and the three longest valid episodes with cameras / guards in play, in order
    keep = [r for r in runs if not r[1]][:1]
    keep += sorted([r for r in runs if r[1] and r[2]], key=lambda r: -len(r[3]))[:3]
    keep.sort(key=lambda r: r[0])
    episodes = []
    for ep, valid, _, tape, metrics, entry in keep:
        idx = len(episodes)
        for i, c in enumerate(tape):
            for key in ("_obs", "_state"):
                if key in c:
                    v = c.pop(key)
                    if key == "_obs":
                        store[f"trainer{idx}/{i}/vec"] = v["vec"]
                        store[f"trainer{idx}/{i}/occ"] = v["occ"]
                        store[f"trainer{idx}/{i}/vis"] = v["vis"]
                    else:
                        store[f"trainer{idx}/{i}/state"] = v
        episodes.append({"episode": ep, "valid": valid, "tape": tape,
                         "log_entry": {k: v for k, v in entry.data.items() if k != "timestamp"},
                         "metrics": {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in metrics.items()}})
    meta["trainer"] = {"R": 20, "C": 20, "max_steps": 40, "episodes": episodes}


def main():
    store, meta = {}, {"traces": [], "numpy": np.__version__, "torch": torch.__version__}
    rng = np.random.default_rng(20261019)
    cases = []
    for vi, (fov, speed, T) in enumerate([(60, 0.1, 640), (90, 0.1, 420), (120, 0.1, 420),
                                          (60, 1.0 / 3.0, 420), (90, 1.0 / 3.0, 420), (120, 1.0 / 3.0, 420)]):
        cases.append((f"tie{fov}_{'tenth' if speed == 0.1 else 'third'}", 20, 20, 200, tie_layout(rng, fov, speed, vi % 2),
                      np.zeros(T, np.int8) if vi % 3 else mg.biased_actions(rng, T)))
    for k in range(4):
        cases.append((f"big32_{k}", 32, 32, 200, big_layout(rng, 32, "4c2g" if k % 2 == 0 else "7c"), mg.biased_actions(rng, 40)))
    for k in range(4):
        cases.append((f"big64_{k}", 64, 64, 200, big_layout(rng, 64, "4c2g" if k % 2 == 0 else "7c"), mg.biased_actions(rng, 36)))
    for name, R, C, ms, layout, actions in cases:
        print("trace", name, flush=True)
        rec = mg.run_trace(R, C, ms, layout, actions, want_state_every=29)
        walls, cams, guards, budget = layout
        meta["traces"].append({"name": name, "R": R, "C": C, "max_steps": ms, "budget": budget,
                               "walls": [list(map(int, w)) for w in walls], "cameras": cams,
                               "guards": [{**g, "patrol_path": [list(map(int, p)) for p in g["patrol_path"]]} for g in guards],
                               "valid": rec["valid"], "spent": int(rec["spent"]), "n_placed": list(map(int, rec["n_placed"]))})
        store[f"{name}/actions"] = actions
        for key in ["grid", "vis0", "reward", "done", "status", "pos", "tick", "vis", "vis_post", "cam_heading",
                    "guard_idx", "guard_heading", "state_t", "state", "obs_vec"]:
            store[f"{name}/{key}"] = rec[key]
        if name.startswith("tie"):   # how close to a multiple of 30 the headings came without hitting it
            h = rec["cam_heading"][:, 0]
            k30 = np.rint(h / 30.0) * 30.0
            near = (np.abs(h - k30) < 1e-9) & (h != k30)
            meta.setdefault("tie_near_ticks", {})[name] = int(near.sum())
    record_trainer_episodes(store, meta)
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(meta["traces"]), "traces")


if __name__ == "__main__":
    main()
