"""Generate tests/golden/heist_golden.npz by running the UNMODIFIED Python reference.

Run in the build container only (the reference tree is not present on the GPU box):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden.py

Everything stored is an output of reference code (heist_architect.*): HeistEnvironment
set_layout/reset/step/get_state_tensor/_get_observation, bfs_path_exists,
ArchitectNetwork.generate_layout's decode loop (driven with one-hot logits so the sampled
asset map is forced), SolverAgent._compute_gae + returns + normalisation, and
RewardCalculator.calculate_architect_reward.  Inputs (layouts, actions) are stored next to
the outputs so the tests never need the reference.
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("HEIST_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import torch  # noqa: E402
from heist_architect.environment import HeistEnvironment, EnvironmentConfig  # noqa: E402
from heist_architect.utils import bfs_path_exists  # noqa: E402
from heist_architect.networks import ArchitectNetwork  # noqa: E402
from heist_architect.agents.solver import SolverAgent  # noqa: E402
from heist_architect.rewards import RewardCalculator  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heist_golden.npz")
STATUS = {"running": 0, "detected": 1, "vault_reached": 2, "timeout": 3, "already_done": 4}


def pack_bits(vis):
    vis = np.asarray(vis) > 0.5
    R, Cn = vis.shape
    W = (Cn + 31) // 32
    out = np.zeros((R, W), np.uint32)
    for c in range(Cn):
        out[:, c // 32] |= vis[:, c].astype(np.uint32) << np.uint32(c % 32)
    return out


def patrol(r, c, H, W):
    offs = [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0)]
    return [(max(1, min(H - 2, r + dr - 1)), max(1, min(W - 2, c + dc - 1))) for dr, dc in offs]


def random_layout(rng, R, C, kind):
    """Random layout dicts in the reference's set_layout format."""
    n_walls = int(rng.integers(0, 10 if R <= 20 else 24))
    walls = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(n_walls)]
    if rng.random() < 0.3 and n_walls:
        walls.append(walls[0])  # duplicate wall: second one is an invalid placement
    n_cams = int(rng.integers(0, 4))
    cams = []
    for _ in range(n_cams):
        if kind == "nice":
            fov = float(rng.choice([30, 45, 60, 90, 120]))
            speed = float(rng.choice([-30, -15, 5, 15, 30, 45]))
            heading = float(rng.choice(np.arange(0, 360, 15)))
        else:
            fov = float(np.float32(rng.uniform(30, 120)))
            speed = float(np.float32(rng.uniform(5, 35))) * (1 if rng.random() < 0.8 else -1)
            heading = float(np.float32(rng.uniform(0, 360)))
        cams.append({"row": int(rng.integers(1, R - 1)), "col": int(rng.integers(1, C - 1)),
                     "fov_angle": fov, "heading": heading, "rotation_speed": speed,
                     "vision_range": int(rng.choice([6, 6, 6, 4, 3]))})
    n_guards = int(rng.integers(0, 3))
    guards = []
    for _ in range(n_guards):
        mode = rng.random()
        if mode < 0.6:
            p = patrol(int(rng.integers(1, R - 1)), int(rng.integers(1, C - 1)), R, C)
            g = {"patrol_path": p, "speed": 1, "vision_range": 4, "fov_angle": 90.0}
        elif mode < 0.85:  # custom path, odd speed: diagonal / long moves -> atan2 headings
            L = int(rng.integers(1, 7))
            p = [(int(rng.integers(1, R - 1)), int(rng.integers(1, C - 1))) for _ in range(L)]
            g = {"patrol_path": p, "speed": int(rng.choice([1, 2, 3, -1])),
                 "vision_range": int(rng.choice([3, 4, 5])), "fov_angle": float(rng.choice([90.0, 60.0, 75.5]))}
        else:  # guard starting on a wall / the start tile: overwrite quirk (environment.py:148)
            tgt = walls[0] if (walls and 0 < walls[0][0] < R - 1 and 0 < walls[0][1] < C - 1) else (1, 1)
            p = [tgt, (tgt[0], min(C - 2, tgt[1] + 1))]
            g = {"patrol_path": p, "speed": 1, "vision_range": 4, "fov_angle": 90.0}
        guards.append(g)
    budget = int(rng.choice([5, 8, 15, 22, 40]))
    return walls, cams, guards, budget


def biased_actions(rng, T):
    # DOWN/RIGHT-heavy so that some traces reach the vault
    return rng.choice(5, size=T, p=[0.1, 0.1, 0.35, 0.1, 0.35]).astype(np.int8)


def run_trace(R, C, max_steps, layout, actions, want_state_every=7):
    walls, cams, guards, budget = layout
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=max_steps, architect_budget=budget)
    env = HeistEnvironment(cfg)
    valid = env.set_layout(walls, cams, guards)
    rec = {"valid": bool(valid), "spent": env.budget.spent, "grid": env.grid.astype(np.int8).copy(),
           "n_placed": (len(env.walls), len(env.cameras), len(env.guards))}
    env.reset()
    T = len(actions)
    W = (C + 31) // 32
    K, G = len(env.cameras), len(env.guards)
    out = {
        "vis0": pack_bits(env.visibility_map.visibility),
        "reward": np.zeros(T, np.float64), "done": np.zeros(T, np.uint8), "status": np.zeros(T, np.uint8),
        "pos": np.zeros((T, 2), np.int16), "tick": np.zeros(T, np.int32),
        "vis": np.zeros((T, R, W), np.uint32),          # after step, before any auto-reset
        "vis_post": np.zeros((T, R, W), np.uint32),     # after the auto-reset (== vis when no reset)
        "cam_heading": np.zeros((T, max(K, 1)), np.float64),
        "guard_idx": np.zeros((T, max(G, 1)), np.int32),
        "guard_heading": np.zeros((T, max(G, 1)), np.float64),
        "state_t": [], "state": [], "obs_vec": [],
    }
    for t in range(T):
        obs, r, done, info = env.step(int(actions[t]))
        out["reward"][t] = r
        out["done"][t] = done
        out["status"][t] = STATUS[info["status"]]
        out["pos"][t] = env.solver_pos
        out["tick"][t] = env.tick
        out["vis"][t] = pack_bits(env.visibility_map.visibility)
        for k, cam in enumerate(env.cameras):
            out["cam_heading"][t, k] = cam.heading
        for k, g in enumerate(env.guards):
            out["guard_idx"][t, k] = g.current_idx
            out["guard_heading"][t, k] = g.heading
        if t % want_state_every == 0:
            out["state_t"].append(t)
            out["state"].append(env.get_state_tensor().astype(np.float32))
            out["obs_vec"].append(np.concatenate([obs["solver_position"], obs["vault_direction"],
                                                  obs["time_feature"]]).astype(np.float32))
            assert np.array_equal(obs["occupancy_grid"], out["state"][-1][0])
            assert np.array_equal(obs["visibility_map"], out["state"][-1][1])
        if done:
            env.reset()
        out["vis_post"][t] = pack_bits(env.visibility_map.visibility)
    out["state_t"] = np.asarray(out["state_t"], np.int32)
    out["state"] = np.stack(out["state"]) if out["state"] else np.zeros((0, 3, R, C), np.float32)
    out["obs_vec"] = np.stack(out["obs_vec"]) if out["obs_vec"] else np.zeros((0, 5), np.float32)
    rec.update(out)
    return rec


def main():
    store = {}
    meta = {"traces": [], "decode": [], "numpy": np.__version__, "torch": torch.__version__}
    rng = np.random.default_rng(20261018)

    # ---- A/B: the reference's own smoke scripts (test_sanity.py:21-41, test_fixes.py:13-37) ----
    fixed = [
        ("sanity", 10, 10, 200,
         ([(3, 3), (3, 4), (3, 5)],
          [{"row": 5, "col": 5, "fov_angle": 60, "heading": 0, "rotation_speed": 15, "vision_range": 4}],
          [{"patrol_path": [(7, 2), (7, 3), (7, 4), (7, 5)], "speed": 1, "vision_range": 3, "fov_angle": 90}],
          15), np.array([4] * 5, np.int8)),
        ("fixes", 10, 10, 200, ([], [], [], 15), np.array([2] * 7 + [4] * 7, np.int8)),
        # survey KAT: negative speed modulo, heading persistence, two timeouts
        ("kat20", 20, 20, 200,
         ([(8, 8), (8, 9), (8, 10), (12, 12), (5, 14)],
          [{"row": 10, "col": 10, "fov_angle": 73.25, "heading": 11.5, "rotation_speed": 17.125, "vision_range": 6},
           {"row": 4, "col": 15, "fov_angle": 120.0, "heading": 200.0, "rotation_speed": -7.5, "vision_range": 6}],
          [{"patrol_path": [(13, 4), (13, 5), (13, 6), (14, 6), (15, 6), (15, 5), (15, 4), (14, 4)],
            "speed": 1, "vision_range": 4, "fov_angle": 90.0}], 22), np.zeros(400, np.int8)),
        # all of detection + vault + timeout in one env: tiny max_steps
        ("short", 10, 10, 14, ([], [{"row": 7, "col": 7, "fov_angle": 120, "heading": 300, "rotation_speed": 0,
                                    "vision_range": 6}], [], 15), np.array([2] * 7 + [4] * 7 + [0] * 6, np.int8)),
        # unsolvable layout (start boxed in): the env still steps (trainer would skip it)
        ("blocked", 10, 10, 200, ([(1, 2), (2, 1), (2, 2)], [], [], 15),
         np.array([4, 2, 0, 1, 3, 2, 4, 4, 2, 2], np.int8)),
        # many timeouts with partial credit (environment.py:291-297)
        ("timeouts", 10, 10, 9, ([(4, 4)], [], [], 15), biased_actions(rng, 72)),
        # three vault runs back to back through auto-reset
        ("vaultrun", 10, 10, 200, ([], [], [], 15), np.array(([2] * 7 + [4] * 7) * 3, np.int8)),
        # guard parked on the vault: detection and vault fire in the same step (:273-288)
        ("vaultguard", 10, 10, 200,
         ([], [], [{"patrol_path": [(8, 8)], "speed": 1, "vision_range": 1, "fov_angle": 30.0}], 15),
         np.array([2] * 7 + [4] * 7 + [0, 4], np.int8)),
        # vault + timeout in the same step (max_steps == path length)
        ("vaulttimeout", 10, 10, 14, ([], [], [], 15), np.array([2] * 7 + [4] * 7 + [2, 2], np.int8)),
    ]
    cases = list(fixed)
    plan = [(10, 10, 60, 8, 40), (20, 20, 200, 20, 48), (32, 32, 200, 6, 40), (64, 64, 200, 2, 30),
            (12, 17, 50, 4, 60)]
    for (R, C, ms, count, T) in plan:
        for k in range(count):
            kind = "nice" if k % 3 == 0 else "f32"
            cases.append((f"rand{R}x{C}_{k}", R, C, ms, random_layout(rng, R, C, kind), biased_actions(rng, T)))

    for name, R, C, ms, layout, actions in cases:
        print("trace", name, flush=True)
        rec = run_trace(R, C, ms, layout, actions)
        walls, cams, guards, budget = layout
        meta["traces"].append({"name": name, "R": R, "C": C, "max_steps": ms, "budget": budget,
                               "walls": [list(map(int, w)) for w in walls], "cameras": cams,
                               "guards": [{**g, "patrol_path": [list(map(int, p)) for p in g["patrol_path"]]}
                                          for g in guards],
                               "valid": rec["valid"], "spent": int(rec["spent"]),
                               "n_placed": list(map(int, rec["n_placed"]))})
        store[f"{name}/actions"] = actions
        for key in ["grid", "vis0", "reward", "done", "status", "pos", "tick", "vis", "vis_post", "cam_heading",
                    "guard_idx", "guard_heading", "state_t", "state", "obs_vec"]:
            store[f"{name}/{key}"] = rec[key]
        if name == "kat20":
            # SURVEY 8c recipe: hash of packbits(vis > .5) before each step
            env = HeistEnvironment(EnvironmentConfig(grid_rows=20, grid_cols=20, architect_budget=22))
            env.set_layout(walls, cams, guards)
            o = env.reset()
            h = hashlib.sha256()
            tot = 0
            for _ in range(400):
                h.update(np.packbits(o["visibility_map"] > 0.5).tobytes())
                tot += int((o["visibility_map"] > 0.5).sum())
                o, r, d, info = env.step(0)
                if d:
                    o = env.reset()
            meta["kat20_sha256"] = h.hexdigest()
            meta["kat20_sum"] = tot
            meta["kat20_final_headings"] = [c.heading for c in env.cameras]

    # ---- E: Architect decode loop (networks.py:273-322) with forced samples ----
    class Forced(ArchitectNetwork):
        def forward(self, grid_state):
            am = torch.as_tensor(self._am, dtype=torch.long)
            H, W = am.shape
            logits = torch.full((1, 4, H, W), -1e4)
            logits.scatter_(1, am.view(1, 1, H, W), 1e4)
            p = {"fov": torch.tensor([[self._p[0]]], dtype=torch.float32),
                 "speed": torch.tensor([[self._p[1]]], dtype=torch.float32),
                 "heading": torch.tensor([[self._p[2]]], dtype=torch.float32)}
            return logits, torch.zeros(1, 1), p

    dec_maps = []
    for k in range(36):
        H, W = [(20, 20), (10, 10), (32, 32), (12, 17)][k % 4]
        dens = [0.75, 0.2, 0.05, 0.02][(k // 4) % 4]
        am = (rng.random((H, W)) < dens) * rng.integers(1, 4, size=(H, W))
        if k % 9 == 0:
            am[1, 1] = 1  # wall on START: decoded, then rejected by set_layout
        if k % 7 == 0:
            am[H - 2, W - 2] = 3  # guard on the vault corner: clamped patrol, duplicates
        budget = int([5, 8, 15, 22][k % 4]) if k % 11 else 0
        prm = (np.float32(rng.uniform(30, 120)), np.float32(rng.uniform(5, 35)), np.float32(rng.uniform(0, 360)))
        net = Forced(grid_rows=H, grid_cols=W)
        net._am, net._p = am, prm
        with torch.no_grad():
            walls, cams, guards, lp, val = net.generate_layout(torch.zeros(1, 1, H, W), budget, 1.0)
        allow_c, allow_g = bool(k % 2 == 0 or k % 3 == 0), bool(k % 3 != 1)
        cams_f = cams if allow_c else []      # curriculum filter, training.py:464-467
        guards_f = guards if allow_g else []
        env = HeistEnvironment(EnvironmentConfig(grid_rows=H, grid_cols=W, architect_budget=budget))
        env.budget.scale_budget(budget)
        valid = env.set_layout(walls, cams_f, guards_f)
        env.reset()
        meta["decode"].append({"H": H, "W": W, "budget": budget, "params": [float(x) for x in prm],
                               "allow_cameras": allow_c, "allow_guards": allow_g,
                               "walls": [list(map(int, w)) for w in walls],
                               "cameras": [{kk: (float(v) if isinstance(v, float) else int(v)) for kk, v in c.items()}
                                           for c in cams],
                               "guards": [[list(map(int, p)) for p in g["patrol_path"]] for g in guards],
                               "valid": bool(valid), "spent": int(env.budget.spent)})
        store[f"decode{k}/asset_map"] = am.astype(np.int8)
        store[f"decode{k}/grid"] = env.grid.astype(np.int8)
        store[f"decode{k}/vis0"] = pack_bits(env.visibility_map.visibility)
        dec_maps.append(k)

    # ---- H: bfs_path_exists (utils.py:52-85) ----
    bfs_grids, bfs_ans, bfs_sg = [], [], []
    for k in range(60):
        R, C = [(20, 20), (32, 32), (64, 64), (10, 10), (9, 33)][k % 5]
        g = np.zeros((R, C), np.int32)
        g[0, :] = g[-1, :] = g[:, 0] = g[:, -1] = 1
        if k % 4 == 3:  # serpentine: long shortest path
            for r in range(2, R - 1, 2):
                g[r, 1:C - 1] = 1
                g[r, (C - 2) if (r // 2) % 2 else 1] = 0
            if k % 8 == 7:
                g[2, 1:C - 1] = 1
        else:
            g[1:-1, 1:-1] = (rng.random((R - 2, C - 2)) < [0.1, 0.3, 0.42][k % 3]).astype(np.int32)
        s, t = (1, 1), (R - 2, C - 2)
        g[s] = 2
        g[t] = 3
        pad = np.zeros((64, 64), np.int8)
        pad[:R, :C] = g
        bfs_grids.append(pad)
        bfs_sg.append([R, C, *s, *t])
        bfs_ans.append(bfs_path_exists(g, s, t))
    store["bfs/grids"] = np.stack(bfs_grids)
    store["bfs/dims"] = np.asarray(bfs_sg, np.int32)
    store["bfs/answer"] = np.asarray(bfs_ans, np.uint8)

    # ---- F: GAE + returns + normalisation (solver.py:141-147, 228-244) ----
    agent = SolverAgent(grid_rows=10, grid_cols=10)
    for k, n in enumerate([1, 2, 7, 64, 500, 2000]):
        rew = (rng.normal(size=n) * [0.1, 1, 10][k % 3]).astype(np.float32)
        val = rng.normal(size=n).astype(np.float32)
        dn = (rng.random(n) < 0.05).astype(np.float32)
        if n > 2:
            dn[-1] = float(k % 2)
        r_t, v_t, d_t = torch.from_numpy(rew), torch.from_numpy(val), torch.from_numpy(dn)
        adv = agent._compute_gae(r_t, v_t, d_t)
        ret = adv + v_t
        norm = (adv - adv.mean()) / (adv.std() + 1e-8) if len(adv) > 1 else adv
        for key, a in [("rew", rew), ("val", val), ("done", dn), ("adv", adv.numpy()), ("ret", ret.numpy()),
                       ("norm", norm.numpy())]:
            store[f"gae{k}/{key}"] = np.asarray(a, np.float32)
    meta["gae_cases"] = 6

    # ---- G: architect reward (rewards.py:43-73) ----
    calc = RewardCalculator()
    env_ok = HeistEnvironment(EnvironmentConfig(grid_rows=10, grid_cols=10))
    env_ok.set_layout([], [], [])
    env_bad = HeistEnvironment(EnvironmentConfig(grid_rows=10, grid_cols=10, architect_budget=40))
    env_bad.set_layout([(1, 2), (2, 1), (2, 2)], [], [])
    assert not env_bad.is_level_valid()
    rates = [k / 20 for k in range(21)] + [1 / 3, 2 / 3, 0.2, 0.6, 0.8, 0.8000000000000002]
    store["arch/solve_rate"] = np.asarray(rates, np.float64)
    store["arch/reward_valid"] = np.asarray([calc.calculate_architect_reward(env_ok, s) for s in rates], np.float64)
    store["arch/reward_invalid"] = np.asarray([calc.calculate_architect_reward(env_bad, s) for s in rates], np.float64)

    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(meta["traces"]), "traces")


if __name__ == "__main__":
    main()
