"""GPU parity tests: the CUDA path (through the C ABI) against golden fixtures from the reference
and against the CPU oracle on seeded inputs.

Bars: visibility bitmaps, positions, ticks, done/status, grids, BFS validity, budgets: bit-exact.
Rewards: float64 == against golden (and float32 == against the oracle's float32 cast).
State tensors / observation vectors: float32 ==.  GAE advantages / returns: float32 == (the
north-star tolerance is 1e-5 relative; the op order is reproduced, so equality holds).
"""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import heist_b200  # noqa: E402
from heist_b200 import BatchedHeistEnv, EnvironmentConfig, synthetic  # noqa: E402
from oracle import heist_oracle as ho  # noqa: E402


def u32(t):
    return t.cpu().numpy().view(np.uint32)


# --------------------------------------------------------------------------------------------
# golden traces (outputs of the unmodified reference)
# --------------------------------------------------------------------------------------------
def _groups(golden):
    groups = {}
    for name, t in golden.traces.items():
        key = (t["R"], t["C"], t["max_steps"], len(golden.arr(name, "actions")))
        groups.setdefault(key, []).append(name)
    return groups


def test_golden_traces_bit_exact(golden):
    _single_tick_traces(golden, 0)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_round2_traces_bit_exact(golden2, mode):
    """tests/golden/heist_golden_r2.npz through step(): camera headings that accumulate to within a few ulp of
    multiples of 30 degrees (rotation_speed 0.1 and 1/3; these rays take their direction from the host-libm table,
    heist_common.cuh) and 32x32 / 64x64 budget-22 layouts -- in the cache, all-fp64 and filtered-march modes."""
    _single_tick_traces(golden2, mode, only=(lambda n: True) if mode == 0 else (lambda n: n.startswith("tie") or n.endswith("_0")))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_round2_traces_step_many(golden2, mode):
    _step_many_traces(golden2, mode)


@pytest.mark.parametrize("mode", [0, 2])
def test_round3_patrol_traces_bit_exact(golden3, mode):
    """tests/golden/heist_golden_r3.npz (recorded from the reference): guard patrols with strides, no-move steps, one- and
    two-waypoint paths, up to four overlapping guards and resets on every phase of a patrol -- the cached path serves
    them from the cones of the reachable (waypoint, heading) pairs only; single ticks and the rollout kernels."""
    _single_tick_traces(golden3, mode)
    _step_many_traces(golden3, mode)


def test_round3_shape_traces_fused_state_every_tick(golden3):
    """The shape* traces of heist_golden_r3.npz (5x4 ... 40x64, recorded from the reference with get_state_tensor after
    EVERY step): heist_step_observe -- the fused tick kernel writing the (3, R, C) state with its own index arithmetic --
    reproduces every state, reward and status; resets by mask as the trainer does."""
    for n in [n for n in golden3.traces if n.startswith("shape")]:
        t = golden3.traces[n]
        R, C = t["R"], t["C"]
        env = BatchedHeistEnv(EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=t["max_steps"]), 1, max_path=8)
        lay = golden3.layout(n)
        assert bool(env.set_layout_explicit([lay[:3]], budget=np.array([lay[3]])).item()) == t["valid"]
        env.check_errors()
        assert env.cache_stats()[0] == 1   # (table-driven: the fused kernel is what runs)
        env.reset()
        acts, states = golden3.arr(n, "actions"), golden3.arr(n, "state")
        state = torch.empty(1, 3, R, C, device="cuda")
        for k in range(len(acts)):
            rew, done, status, _ = env.step_observe(acts[k:k + 1], autoreset=False, state_out=state)
            assert rew.item() == np.float32(golden3.arr(n, "reward")[k]) and status.item() == golden3.arr(n, "status")[k], (n, k)
            assert np.array_equal(state[0].cpu().numpy(), states[k]), (n, k)
            env.reset(mask=done)
            assert np.array_equal(u32(env.visibility_bits)[0], golden3.arr(n, "vis_post")[k]), (n, k)
        env.check_errors()
        env.close()


def _single_tick_traces(golden, mode, only=lambda n: True):
    for (R, C, ms, T), names in _groups(golden).items():
        names = [n for n in names if only(n)]
        if not names:
            continue
        N = len(names)
        env = BatchedHeistEnv(EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=ms), N,
                              max_walls=64, max_cams=8, max_guards=4, max_path=8)
        env.set_mode(mode)
        lays = [golden.layout(n) for n in names]
        valid = env.set_layout_explicit([l[:3] for l in lays], budget=np.array([l[3] for l in lays]))
        env.check_errors()
        assert valid.cpu().tolist() == [golden.traces[n]["valid"] for n in names]
        assert env.budget_spent.cpu().tolist() == [golden.traces[n]["spent"] for n in names]
        ncam = env.env_static[:, 0].cpu().numpy()
        ngrd = env.env_static[:, 1].cpu().numpy()
        for j, n in enumerate(names):
            assert [ncam[j], ngrd[j]] == golden.traces[n]["n_placed"][1:], n
            assert np.array_equal(env.tile_codes[j].cpu().numpy().astype(np.int8), golden.arr(n, "grid")), n
        env.reset()
        vis = u32(env.visibility_bits)
        for j, n in enumerate(names):
            assert np.array_equal(vis[j], golden.arr(n, "vis0")), n
        acts = np.stack([golden.arr(n, "actions") for n in names], 1)
        G = {k: [golden.arr(n, k) for n in names] for k in
             ["reward", "done", "status", "pos", "tick", "vis", "vis_post", "cam_heading", "guard_idx",
              "guard_heading", "state_t", "state", "obs_vec"]}
        si = [0] * N
        for t in range(T):
            rew, done, status, r64 = env.step(acts[t], want_reward64=True)
            r64, rew = r64.cpu().numpy(), rew.cpu().numpy()
            done_h, status_h = done.cpu().numpy(), status.cpu().numpy()
            pos, tick = env.solver_pos.cpu().numpy(), env.tick.cpu().numpy()
            vis = u32(env.visibility_bits)
            ch, gh, gx = env.cam_heading.cpu().numpy(), env.guard_heading.cpu().numpy(), env.guard_idx.cpu().numpy()
            need_state = any(si[j] < len(G["state_t"][j]) and G["state_t"][j][si[j]] == t for j in range(N))
            if need_state:
                state = env.observe().cpu().numpy()
                o = env.observation()
                vec = torch.cat([o["solver_position"], o["vault_direction"], o["time_feature"]], 1).cpu().numpy()
                assert np.array_equal(o["occupancy_grid"].cpu().numpy(), state[:, 0])
                assert np.array_equal(o["visibility_map"].cpu().numpy(), state[:, 1])
            for j, n in enumerate(names):
                where = (n, t)
                assert r64[j] == G["reward"][j][t], where
                assert rew[j] == np.float32(G["reward"][j][t]), where
                assert done_h[j] == bool(G["done"][j][t]) and status_h[j] == G["status"][j][t], where
                assert tuple(pos[j]) == tuple(G["pos"][j][t]) and tick[j] == G["tick"][j][t], where
                assert np.array_equal(vis[j], G["vis"][j][t]), where
                assert np.array_equal(ch[j, :ncam[j]], G["cam_heading"][j][t, :ncam[j]]), where
                assert np.array_equal(gx[j, :ngrd[j]], G["guard_idx"][j][t, :ngrd[j]]), where
                assert np.array_equal(gh[j, :ngrd[j]], G["guard_heading"][j][t, :ngrd[j]]), where
                if si[j] < len(G["state_t"][j]) and G["state_t"][j][si[j]] == t:
                    assert np.array_equal(state[j], G["state"][j][si[j]]), where
                    assert np.array_equal(vec[j], G["obs_vec"][j][si[j]]), where
                    si[j] += 1
            env.reset(mask=done)  # the trainer's pattern: reset after a finished episode
            vis = u32(env.visibility_bits)
            for j, n in enumerate(names):
                assert np.array_equal(vis[j], G["vis_post"][j][t]), (n, t, "post")
        env.close()


def test_golden_traces_step_many(golden):
    """Same traces through the multi-step kernel with in-kernel auto-reset."""
    _step_many_traces(golden, 0)


def _step_many_traces(golden, mode):
    for (R, C, ms, T), names in _groups(golden).items():
        N = len(names)
        env = BatchedHeistEnv(EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=ms), N)
        env.set_mode(mode)
        lays = [golden.layout(n) for n in names]
        env.set_layout_explicit([l[:3] for l in lays], budget=np.array([l[3] for l in lays]))
        env.reset()
        acts = np.stack([golden.arr(n, "actions") for n in names], 1)
        out = env.step_many(acts, autoreset=True, want_vis=True)
        rew, done, status, vis = (out["reward"].cpu().numpy(), out["done"].cpu().numpy(),
                                  out["status"].cpu().numpy(), u32(out["vis_bits"]))
        for j, n in enumerate(names):
            assert np.array_equal(rew[:, j], golden.arr(n, "reward").astype(np.float32)), n
            assert np.array_equal(done[:, j], golden.arr(n, "done")), n
            assert np.array_equal(status[:, j], golden.arr(n, "status")), n
            assert np.array_equal(vis[:, j], golden.arr(n, "vis_post")), n
        env.close()


def test_survey_kat20_hash(golden):
    import hashlib
    walls, cams, guards, budget = golden.layout("kat20")
    env = BatchedHeistEnv(EnvironmentConfig(), 1)
    assert env.set_layout_explicit([(walls, cams, guards)], budget=budget).item()
    env.reset()
    v0 = u32(env.visibility_bits)[0]
    out = env.step_many(np.zeros((400, 1), np.int8), autoreset=True, want_vis=True)
    vis = np.concatenate([v0[None], u32(out["vis_bits"])[:-1, 0]])  # map seen before each of the 400 steps
    dense = ((vis[:, :, 0, None] >> np.arange(20)) & 1).astype(bool)
    h = hashlib.sha256()
    for t in range(400):
        h.update(np.packbits(dense[t]).tobytes())
    assert h.hexdigest() == "513d958ef974b21eade9aa03015912517dece7e4b253561dd48067fbad01e8b6"
    assert int(dense.sum()) == 28938 and int(out["done"].sum().item()) == 2
    assert env.cam_heading[0, :2].cpu().tolist() == [21.5, 80.0]


def test_decode_golden(golden):
    for k, d in enumerate(golden.meta["decode"]):
        am = golden.z[f"decode{k}/asset_map"]
        env = BatchedHeistEnv(EnvironmentConfig(grid_rows=d["H"], grid_cols=d["W"]), 1, max_walls=64, max_cams=8,
                              max_guards=8)
        valid = env.set_layout_from_asset_map(am[None], np.asarray([d["params"]], np.float32), budget=d["budget"],
                                              allow_cameras=d["allow_cameras"], allow_guards=d["allow_guards"])
        env.check_errors()
        assert bool(valid.item()) == d["valid"], k
        assert env.budget_spent.item() == d["spent"], k
        assert np.array_equal(env.tile_codes[0].cpu().numpy().astype(np.int8), golden.z[f"decode{k}/grid"]), k
        env.reset()
        assert np.array_equal(u32(env.visibility_bits)[0], golden.z[f"decode{k}/vis0"]), k
        env.close()


def test_bfs_golden(golden):
    grids, dims, ans = golden.z["bfs/grids"], golden.z["bfs/dims"], golden.z["bfs/answer"]
    for g, (R, C, sr, sc, gr, gc), a in zip(grids, dims, ans):
        walls = [(int(r), int(c)) for r, c in zip(*np.nonzero(g[1:R - 1, 1:C - 1] == 1))]
        walls = [(r + 1, c + 1) for r, c in walls]
        env = BatchedHeistEnv(EnvironmentConfig(grid_rows=int(R), grid_cols=int(C), start_pos=(int(sr), int(sc)),
                                                vault_pos=(int(gr), int(gc))), 1, max_walls=4096)
        valid = env.set_layout_explicit([(walls, [], [])], budget=100000)
        env.check_errors()
        assert bool(valid.item()) == bool(a)
        assert np.array_equal(env.tile_codes[0].cpu().numpy(), g[:R, :C].astype(np.uint8))
        env.close()


def test_gae_golden(golden):
    for k in range(golden.meta["gae_cases"]):
        g = {key: golden.z[f"gae{k}/{key}"] for key in ["rew", "val", "done", "adv", "ret", "norm"]}
        adv, ret = heist_b200.compute_gae(torch.from_numpy(g["rew"]).cuda(), torch.from_numpy(g["val"]).cuda(),
                                          torch.from_numpy(g["done"]).cuda().to(torch.uint8))
        assert np.array_equal(adv.cpu().numpy(), g["adv"]), k
        assert np.array_equal(ret.cpu().numpy(), g["ret"]), k
        norm = heist_b200.normalize_advantages(adv)
        np.testing.assert_allclose(norm.cpu().numpy(), g["norm"], rtol=1e-5, atol=1e-6)


def test_architect_reward_golden(golden):
    # drive the per-env outcome counters to the golden solve rates through the state view
    rates = [(k, 20) for k in range(21)] + [(1, 3), (2, 3)]
    env = BatchedHeistEnv(EnvironmentConfig(grid_rows=10, grid_cols=10), len(rates))
    for j, (k, n) in enumerate(rates):
        env.env_dyn[j, 5] = k
        env.env_dyn[j, 6] = n - k
    rw, sr = env.architect_reward()
    exp = {float(s): float(r) for s, r in zip(golden.z["arch/solve_rate"], golden.z["arch/reward_valid"])}
    for j, (k, n) in enumerate(rates):
        assert sr[j].item() == k / n and rw[j].item() == exp[k / n], (k, n)
    env.set_layout_explicit([([(1, 2), (2, 1), (2, 2)], [], [])] * len(rates))
    rw, _ = env.architect_reward()
    assert (rw == -1.0).all()


# --------------------------------------------------------------------------------------------
# seeded differential tests against the CPU oracle
# --------------------------------------------------------------------------------------------
def oracle_envs(am, cp, cfg, budget, allow_cameras=True, allow_guards=True):
    envs, valid = [], []
    for i in range(am.shape[0]):
        walls, cams, guards, _ = ho.decode_layout(am[i], budget, cp[i, 0], cp[i, 1], cp[i, 2])
        e = ho.OracleEnv(cfg.grid_rows, cfg.grid_cols, max_steps=cfg.max_steps, budget=budget)
        valid.append(e.set_layout(walls, cams if allow_cameras else [], guards if allow_guards else []))
        envs.append(e)
    return envs, np.asarray(valid)


@pytest.mark.parametrize("R,C,N,T,budget,nice", [
    (20, 20, 512, 200, 15, False),     # config 2 parity subset (>= 256 envs x 200 ticks)
    (20, 20, 256, 200, 15, True),      # tie-prone "nice" angles (fov 60, speed 15, heading k*15)
    (32, 32, 128, 120, 22, False),     # config 3 shape
    (64, 64, 48, 80, 22, False),       # config 4 shape
    (12, 17, 64, 50, 8, False),        # ragged grid, C not a multiple of 4
])
def test_rollout_matches_oracle(R, C, N, T, budget, nice):
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=min(200, T))
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(synthetic.BASE_SEED + R * 1000 + N)
    am = synthetic.sample_asset_maps(rng, N, R, C)
    cp = synthetic.sample_cam_params(rng, N, nice)
    valid = env.set_layout_from_asset_map(am, cp, budget)
    env.check_errors()
    oenvs, ovalid = oracle_envs(am, cp, cfg, budget)
    assert np.array_equal(valid.cpu().numpy(), ovalid)
    assert np.array_equal(env.tile_codes.cpu().numpy(), np.stack([e.grid for e in oenvs]).astype(np.uint8))
    assert env.budget_spent.cpu().tolist() == [e.info()["spent"] for e in oenvs]
    env.reset()
    ho.reset_all(oenvs)
    assert np.array_equal(u32(env.visibility_bits), np.stack([ho.pack_bits(e.visibility) for e in oenvs]))
    acts = synthetic.sample_actions(rng, T, N)
    out = env.step_many(acts, autoreset=True, want_vis=True)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    assert np.array_equal(out["done"].cpu().numpy(), ref["done"])
    assert np.array_equal(out["status"].cpu().numpy(), ref["status"])
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    assert np.array_equal(out["reward"].cpu().numpy(), ref["reward"])  # float32(double) on both sides
    # final state: positions, ticks, headings, observation tensors
    info = [e.info() for e in oenvs]
    assert env.solver_pos.cpu().tolist() == [[i["solver_r"], i["solver_c"]] for i in info]
    assert env.tick.cpu().tolist() == [i["tick"] for i in info]
    ch = env.cam_heading.cpu().numpy()
    for j, e in enumerate(oenvs):
        assert np.array_equal(ch[j, :info[j]["n_cams"]], e.cam_headings()), j
    state = env.observe().cpu().numpy()
    assert np.array_equal(state, np.stack([e.state_tensor() for e in oenvs]))
    o = env.observation()
    vec = torch.cat([o["solver_position"], o["vault_direction"], o["time_feature"]], 1).cpu().numpy()
    assert np.array_equal(vec, np.stack([np.concatenate(e.obs_vectors()) for e in oenvs]))
    # outcome counters -> architect reward
    rw, sr = env.architect_reward()
    st = ref["status"]
    for j in range(N):
        nv, nd, nt = (st[:, j] == 2).sum(), (st[:, j] == 1).sum(), (st[:, j] == 3).sum()
        rate = nv / (nv + nd + nt) if (nv + nd + nt) else 0.0
        assert sr[j].item() == rate and rw[j].item() == ho.architect_reward(ovalid[j], rate), j
    env.close()


def test_no_autoreset_done_envs_freeze():
    cfg = EnvironmentConfig(max_steps=30)
    env = BatchedHeistEnv(cfg, 64)
    rng = np.random.default_rng(5)
    am, cp = synthetic.sample_asset_maps(rng, 64, 20, 20), synthetic.sample_cam_params(rng, 64)
    env.set_layout_from_asset_map(am, cp, 15)
    oenvs, _ = oracle_envs(am, cp, cfg, 15)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, 45, 64)
    out = env.step_many(acts, autoreset=False, want_vis=True)
    ref = ho.rollout(oenvs, acts, autoreset=False, want_vis=True)
    for k in ["done", "status", "reward"]:
        assert np.array_equal(out[k].cpu().numpy(), ref[k]), k
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    assert (ref["status"][-1] == 4).all()  # everything timed out at tick 30 at the latest


@pytest.mark.parametrize("allow_c,allow_g", [(False, False), (True, False), (True, True)])
def test_decode_validate_matches_oracle_dense_maps(allow_c, allow_g):
    """Dense early-training asset maps (75 % non-zero) and the BFS stress variant (~50 % invalid)."""
    R = C = 32
    N = 768
    rng = np.random.default_rng(77)
    am = synthetic.sample_asset_maps(rng, N, R, C, 0.25, 0.25, 0.25)
    am[N // 2:] = synthetic.sample_asset_maps(rng, N - N // 2, R, C, 0.30, 0.004, 0.002)
    cp = synthetic.sample_cam_params(rng, N)
    budgets = rng.choice([0, 5, 8, 15, 22, 60], N).astype(np.int32)
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C)
    env = BatchedHeistEnv(cfg, N, max_walls=64, max_cams=20, max_guards=12)
    valid = env.set_layout_from_asset_map(am, cp, budgets, allow_c, allow_g).cpu().numpy()
    env.check_errors()
    env.reset()
    vis = u32(env.visibility_bits)
    tiles, spent = env.tile_codes.cpu().numpy(), env.budget_spent.cpu().numpy()
    n_invalid = 0
    for i in range(N):
        walls, cams, guards, _ = ho.decode_layout(am[i], int(budgets[i]), *cp[i])
        e = ho.OracleEnv(R, C, budget=int(budgets[i]))
        v = e.set_layout(walls, cams if allow_c else [], guards if allow_g else [])
        n_invalid += not v
        assert v == valid[i], i
        assert np.array_equal(e.grid.astype(np.uint8), tiles[i]) and e.info()["spent"] == spent[i], i
        e.reset()
        assert np.array_equal(ho.pack_bits(e.visibility), vis[i]), i
    assert n_invalid > 0


def test_bfs_stress_half_invalid():
    R = C = 32
    N = 1024
    rng = np.random.default_rng(99)
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C)
    env = BatchedHeistEnv(cfg, N, max_walls=1024)
    lays, exp = [], []
    for i in range(N):
        p = rng.uniform(0.25, 0.5)
        g = np.zeros((R, C), np.int32)
        g[0, :] = g[-1, :] = g[:, 0] = g[:, -1] = 1
        g[1:-1, 1:-1] = rng.random((R - 2, C - 2)) < p
        g[1, 1], g[R - 2, C - 2] = 2, 3
        walls = [(int(r), int(c)) for r, c in zip(*np.nonzero(g == 1)) if 0 < r < R - 1 and 0 < c < C - 1]
        lays.append((walls, [], []))
        exp.append(ho.bfs(g, (1, 1), (R - 2, C - 2)))
    valid = env.set_layout_explicit(lays, budget=10000).cpu().numpy()
    env.check_errors()
    assert np.array_equal(valid, np.asarray(exp))
    assert 0.1 < np.mean(exp) < 0.9


def test_gae_matches_oracle_large():
    rng = np.random.default_rng(11)
    T, N = 200, 4096
    rew = rng.normal(size=(T, N)).astype(np.float32)
    val = rng.normal(size=(T, N)).astype(np.float32)
    dn = (rng.random((T, N)) < 0.02).astype(np.uint8)
    adv, ret = heist_b200.compute_gae(torch.from_numpy(rew).cuda(), torch.from_numpy(val).cuda(),
                                      torch.from_numpy(dn).cuda())
    oadv, oret = ho.gae(rew, val, dn.astype(np.float32))
    assert np.array_equal(adv.cpu().numpy(), oadv) and np.array_equal(ret.cpu().numpy(), oret)


def test_full_size_config2_properties_and_subset_parity():
    """BASELINE config 2 at full size: 4096 envs, 20x20, random valid layouts, T=200, auto-reset."""
    cfg = EnvironmentConfig()
    N, T = 4096, 200
    env = BatchedHeistEnv(cfg, N)
    am, cp = synthetic.make_valid_workload(env, synthetic.BASE_SEED, 15)
    assert env.valid.all()
    acts = synthetic.sample_actions(np.random.default_rng(synthetic.BASE_SEED + 1), T, N)
    env.reset()
    a = env.step_many(acts, autoreset=True, want_vis=True)
    a = {k: v.clone() for k, v in a.items()}
    # idempotence / determinism: same layouts, reset, same actions -> identical trajectory
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    b = env.step_many(acts, autoreset=True, want_vis=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # size-independent invariants
    st, dn, rw = a["status"], a["done"].bool(), a["reward"]
    assert ((st == 0) == ~dn).all()                  # running <=> not done (auto-reset: never already_done)
    assert (st <= 3).all()
    assert (rw[st == 2] > 8.0).all()                 # vault bonus present
    assert (rw[st == 0].abs() <= 0.26 + 1e-6).all()  # running step: -0.01 +- 0.1 (+ <= 0.15 proximity)
    # first 256 envs bit-exact against the oracle
    sub = 256
    oenvs, ovalid = oracle_envs(am[:sub], cp[:sub], cfg, 15)
    assert ovalid.all()
    ho.reset_all(oenvs)
    ref = ho.rollout(oenvs, acts[:, :sub].copy(), autoreset=True, want_vis=True)
    assert np.array_equal(a["done"][:, :sub].cpu().numpy(), ref["done"])
    assert np.array_equal(a["status"][:, :sub].cpu().numpy(), ref["status"])
    assert np.array_equal(a["reward"][:, :sub].cpu().numpy(), ref["reward"])
    assert np.array_equal(u32(a["vis_bits"][:, :sub]), ref["vis_bits"])


# --------------------------------------------------------------------------------------------
# facade: the reference's own smoke scripts, through the HeistEnvironment-compatible class
# --------------------------------------------------------------------------------------------
def test_facade_reference_smoke_scripts():
    cfg = EnvironmentConfig(grid_rows=10, grid_cols=10, start_pos=(1, 1), vault_pos=(8, 8))
    env = heist_b200.HeistEnvironment(cfg)
    valid = env.set_layout(
        [(3, 3), (3, 4), (3, 5)],
        [{"row": 5, "col": 5, "fov_angle": 60, "heading": 0, "rotation_speed": 15, "vision_range": 4}],
        [{"patrol_path": [(7, 2), (7, 3), (7, 4), (7, 5)], "speed": 1, "vision_range": 3, "fov_angle": 90}])
    assert valid and env.budget.spent == 11
    obs = env.reset()
    assert sorted(obs) == ["occupancy_grid", "solver_position", "time_feature", "vault_direction", "visibility_map"]
    assert int((env.visibility_map.visibility > 0.5).sum()) == 23
    rewards = []
    for _ in range(5):
        obs, reward, done, info = env.step(4)
        rewards.append(reward)
        if done:
            break
    assert rewards == [0.09000000000000001] * 4 + [-0.91]
    assert env.solver_pos == (1, 6) and env.tick == 5 and info == {"status": "detected", "tick": 4}
    assert int((env.visibility_map.visibility > 0.5).sum()) == 21
    assert [c.heading for c in env.cameras] == [75.0] and [g.heading for g in env.guards] == [0.0]
    assert env.step(0)[1:] == (0.0, True, {"status": "already_done"})
    assert env.get_state_tensor().shape == (3, 10, 10)
    assert env.render_text().splitlines()[5] == "#....C...#"
    # test_fixes.py:13-37
    env = heist_b200.HeistEnvironment(EnvironmentConfig(grid_rows=10, grid_cols=10))
    env.set_layout([], [], [])
    env.reset()
    tot = sum(env.step(2)[1] for _ in range(7))
    assert f"{tot:+.3f}" == "+0.630"
    for _ in range(7):
        obs, r, done, info = env.step(4)
        tot += r
    assert f"{tot:+.3f}" == "+11.560" and info["status"] == "vault_reached" and env.vault_reached
    env.reset()
    s = env.get_state_tensor()
    assert f"{s[2].min():.3f}" == "-1.000" and f"{s[2].max():.3f}" == "0.790"


def test_facade_replays_trainer_tapes(golden2):
    """BASELINE config 1 drop-in: the exact call sequence the unmodified AdversarialTrainer._run_one_episode
    (training.py:418-600) made on the reference's HeistEnvironment -- budget.scale_budget, set_layout, per attempt
    reset / get_state_tensor / step / tick, is_level_valid (rewards.py:58), get_environment_state -- replayed on
    the facade: every return value must equal what the reference returned (tests/golden/make_golden_r2.py)."""
    tr = golden2.meta["trainer"]
    cfg = EnvironmentConfig(grid_rows=tr["R"], grid_cols=tr["C"], max_steps=tr["max_steps"])
    env = heist_b200.HeistEnvironment(cfg)
    n_steps = 0
    for k, ep in enumerate(tr["episodes"]):
        for i, c in enumerate(ep["tape"]):
            where = (k, i, c["call"])
            arr = lambda key: golden2.z[f"trainer{k}/{i}/{key}"]

            def check_obs(obs):
                assert sorted(obs) == ["occupancy_grid", "solver_position", "time_feature", "vault_direction", "visibility_map"]
                assert np.array_equal(obs["occupancy_grid"], arr("occ")) and np.array_equal(obs["visibility_map"], arr("vis")), where
                vec = np.concatenate([obs["solver_position"], obs["vault_direction"], obs["time_feature"]])
                assert vec.dtype == np.float32 and np.array_equal(vec, arr("vec")), where

            if c["call"] == "scale_budget":
                env.budget.scale_budget(c["arg"])
                assert (env.budget.spent, env.budget.remaining) == (c["spent"], c["remaining"]), where
            elif c["call"] == "set_layout":
                guards = [{**g, "patrol_path": [tuple(p) for p in g["patrol_path"]]} for g in c["guards"]]
                assert env.set_layout([tuple(w) for w in c["walls"]], c["cameras"], guards) == c["ret"], where
                assert env.budget.spent == c["spent"], where
                assert [len(env.walls), len(env.cameras), len(env.guards)] == c["n_placed"], where
                assert [[w.row, w.col] for w in env.walls] == c["walls_placed"], where
                # (tick is whatever the previous episode left behind -- set_layout does not touch it -- and the
                # fixture keeps a subset of the episodes the trainer ran)
                assert repr(env).split(", tick=")[0] == c["repr"].split(", tick=")[0], where
            elif c["call"] == "is_level_valid":
                assert env.is_level_valid() == c["ret"], where
            elif c["call"] == "reset":
                check_obs(env.reset())
            elif c["call"] == "step":
                obs, r, d, info = env.step(c["arg"])
                assert (r, d, info) == (c["reward"], c["done"], c["info"]), where
                assert env.tick == c["tick_after"], where
                check_obs(obs)
                n_steps += 1
            elif c["call"] == "get_state_tensor":
                st = env.get_state_tensor()
                assert st.dtype == np.float32 and np.array_equal(st, arr("state")), where
            elif c["call"] == "get_environment_state":
                got = json.loads(json.dumps(env.get_environment_state()))   # tuples -> lists, as in the fixture
                assert sorted(got) == sorted(c["ret"]), where
                for key in c["ret"]:
                    assert got[key] == c["ret"][key], (where, key)
    assert n_steps >= 30


def test_ray_directions_match_host_libm_near_multiples_of_30():
    """The directions the kernels use (ray_dir, heist_common.cuh) against the platform libm, which is what the
    reference calls (security.py:71-75): every double within 40 ulp of a multiple of 30 degrees (|a| <= 1440), the
    angles speed 0.1 / 0.3 / 1/3 accumulate to, the window around 0, and the window edges must be EQUAL; generic
    angles (device sincos) are only reported."""
    import ctypes as C
    import math
    lib = heist_b200.load_library()
    env = BatchedHeistEnv(EnvironmentConfig(), 1)   # uploads the table
    angles = []
    for k in range(-48, 49):
        c = k * 30.0
        a = c
        lo = [a := math.nextafter(a, -math.inf) for _ in range(40)]
        a = c
        hi = [a := math.nextafter(a, math.inf) for _ in range(40)]
        angles += [c] + lo + hi + [c - 9.9e-12, c + 9.9e-12, c - 1e-11, c + 1e-11, c - 1.01e-11, c + 1.01e-11, c + 3e-13, c - 7e-13]
    for speed in (0.1, 0.3, 1.0 / 3.0, -0.7):
        h = 0.0
        for t in range(4000):
            h = (h + speed) % 360.0
            for off in (0.0, -30.0, 45.0, -60.0):
                a = h + off
                if abs(a - round(a / 30.0) * 30.0) < 2e-11:
                    angles.append(a)
    n_table = len(angles)
    rng = np.random.default_rng(5)
    angles += list(rng.uniform(-720, 1080, 20000)) + [float(np.float32(x)) for x in rng.uniform(0, 360, 20000)]
    a = torch.tensor(angles, dtype=torch.float64, device="cuda")
    dx, dy = torch.empty_like(a), torch.empty_like(a)
    rc = lib.heist_debug_ray_dirs(0, C.c_void_p(a.data_ptr()), len(angles), C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()), None)
    assert rc == 0
    dx, dy = dx.cpu().numpy(), dy.cpu().numpy()
    ref_dx = np.array([math.cos(math.radians(x)) for x in angles])
    ref_dy = np.array([-math.sin(math.radians(x)) for x in angles])
    def in_window(x):   # the table's windows: [fl(|c| - 1e-11), fl(|c| + 1e-11)] around c = k * 30, |k| <= 48
        c = abs(round(x / 30.0) * 30.0)
        return c <= 1440.0 and c - 1e-11 <= abs(x) <= c + 1e-11

    must = [i for i in range(len(angles)) if in_window(angles[i])]
    assert len(must) >= 97 * 85 and sum(1 for i in must if angles[i] != round(angles[i] / 30.0) * 30.0 and abs(angles[i]) > 1e-9) > 8000
    bad = [(angles[i], dx[i], ref_dx[i], dy[i], ref_dy[i]) for i in must if dx[i] != ref_dx[i] or dy[i] != ref_dy[i]]
    assert not bad, bad[:5]
    rest = np.array([i for i in range(len(angles)) if not in_window(angles[i])])
    generic_diff = int((dx[rest] != ref_dx[rest]).sum() + (dy[rest] != ref_dy[rest]).sum())
    print(f"ray_dir: {len(must)} near-multiple angles equal; other angles differing in the last ulp: {generic_diff} of {2 * len(rest)}")
    env.close()


@pytest.mark.parametrize("R,C,N,T,counts,nice", [
    (20, 20, 4096, 200, None, False),
    (20, 20, 2048, 200, None, True),
    (32, 32, 2048, 100, (2, 4, 2), False),   # config 3: budget 22 = 4 cameras + 2 guards (+2 walls)
    (32, 32, 1024, 100, (1, 7, 0), True),
    (64, 64, 1024, 60, (2, 4, 2), False),
])
def test_cache_and_filtered_march_equal_all_fp64_path(R, C, N, T, counts, nice):
    """Device vs device at full size: the angular visibility cache (default mode) and the fixed-point filter +
    exact fallback (march mode) must give the same bits as routing every sample through the fp64 reference
    arithmetic."""
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C)
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(4242 + R + N)
    am = (synthetic.sample_asset_maps(rng, N, R, C) if counts is None
          else synthetic.sample_asset_maps_exact(rng, N, R, C, *counts))
    cp = synthetic.sample_cam_params(rng, N, nice)
    acts = synthetic.sample_actions(rng, T, N)
    res = []
    for mode in (env.MODE_DEFAULT, env.MODE_MARCH, env.MODE_EXACT):
        env.set_mode(mode)
        env.set_layout_from_asset_map(am, cp, 22 if counts else 15)
        if mode == env.MODE_DEFAULT:   # Architect-decoded layouts are what the cache is built for: all of them
            assert env.cache_stats()[0] == N
        env.reset()
        v0 = env.visibility_bits.clone()
        out = env.step_many(acts, autoreset=True, want_vis=True)
        res.append((v0, {k: v.clone() for k, v in out.items()}, env.cam_heading.clone(), env.env_dyn.clone(),
                    env.guard_heading.clone(), env.guard_idx.clone(), env.visibility_bits.clone()))
    env.check_errors()
    for other in res[1:]:
        assert torch.equal(res[0][0], other[0])
        for k in res[0][1]:
            assert torch.equal(res[0][1][k], other[1][k]), k
        for i in range(2, 7):
            assert torch.equal(res[0][i], other[i]), i
    assert res[0][1]["vis_bits"].ne(0).any()


@pytest.mark.parametrize("R,N,autoreset", [(20, 1024, True), (20, 512, False), (64, 256, True)])
def test_cached_path_launch_shapes_agree(R, N, autoreset):
    """The table-driven path has several launch shapes -- pipelined chunks (auto-reset, T > 32), sequential chunks,
    single ticks, with the maps built in the caller's trajectory or in scratch.  They must agree with each other
    and carry state across launches (envs that are done when a launch begins included)."""
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=R, max_steps=40)
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(77 + R)
    am = synthetic.sample_asset_maps(rng, N, R, R)
    cp = synthetic.sample_cam_params(rng, N)
    T = 100
    acts = torch.as_tensor(synthetic.sample_actions(rng, T, N)).cuda()

    def fresh():
        env.set_layout_from_asset_map(am, cp, 15)
        env.reset()

    def state():
        return [x.clone() for x in (env.env_dyn, env.cam_heading, env.guard_heading, env.guard_idx, env.visibility_bits)]

    fresh()
    ref = env.step_many(acts, autoreset=autoreset, want_vis=True)
    ref = {k: v.clone() for k, v in ref.items()}
    ref_state = state()
    # (1) no trajectory buffer: maps built in scratch, same outputs and same final state
    fresh()
    out = env.step_many(acts, autoreset=autoreset, want_vis=False)
    for k in ("reward", "done", "status"):
        assert torch.equal(out[k], ref[k]), k
    for a, b in zip(state(), ref_state):
        assert torch.equal(a, b)
    # (2) the same rollout in uneven pieces (33 + 1 + 7 + 59 ticks)
    fresh()
    t0 = 0
    for n in (33, 1, 7, 59):
        o = env.step_many(acts[t0:t0 + n], autoreset=autoreset, want_vis=True)
        for k in ref:
            assert torch.equal(o[k], ref[k][t0:t0 + n]), (k, t0)
        t0 += n
    for a, b in zip(state(), ref_state):
        assert torch.equal(a, b)
    # (3) against the ray-march on the same inputs
    env.set_mode(env.MODE_MARCH)
    fresh()
    o = env.step_many(acts, autoreset=autoreset, want_vis=True)
    for k in ref:
        assert torch.equal(o[k], ref[k]), k
    for a, b in zip(state(), ref_state):
        assert torch.equal(a, b)
    env.check_errors()


@pytest.mark.parametrize("seed", range(6))
def test_cached_path_vs_ray_march_ragged_shapes(seed):
    """Soak: ragged grids, odd batch sizes, dense asset maps, nice / generic angles, both auto-reset modes, one call or
    uneven pieces -- the table-driven path and the ray-march must agree on every output and on the final state."""
    rng = np.random.default_rng(1000 + seed)
    R, C = [(20, 20), (32, 32), (64, 64), (9, 40), (50, 11), (33, 33)][seed]
    N, T = int(rng.choice([1, 31, 257])), int(rng.choice([40, 97]))
    budget, max_steps = int(rng.choice([8, 15, 22])), int(rng.choice([25, 200]))
    autoreset, nice = seed % 4 < 2, bool(seed % 2)
    pieces = [T] if seed % 3 else [T // 3, 1, T - T // 3 - 1]
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, architect_budget=budget, max_steps=max_steps)
    env = BatchedHeistEnv(cfg, N)
    am = synthetic.sample_asset_maps(rng, N, R, C, p_wall=0.04, p_cam=0.02, p_guard=0.01)
    cp = synthetic.sample_cam_params(rng, N, nice)
    acts = torch.as_tensor(synthetic.sample_actions(rng, T, N)).cuda()
    res = []
    for mode in (env.MODE_DEFAULT, env.MODE_MARCH):
        env.set_mode(mode)
        env.set_layout_from_asset_map(am, cp, budget)
        env.reset()
        outs, t0 = [], 0
        for n in pieces:
            o = env.step_many(acts[t0:t0 + n], autoreset=autoreset, want_vis=True)
            outs.append({k: v.clone() for k, v in o.items()})
            t0 += n
        res.append((outs, [x.clone() for x in (env.env_dyn, env.cam_heading, env.guard_heading, env.guard_idx,
                                               env.visibility_bits)]))
    for a, b in zip(res[0][0], res[1][0]):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    env.check_errors()


def test_pipelined_step_many_is_graph_capturable():
    """The pipelined launch forks onto internal streams and joins back: once its scratch buffers exist it can be
    captured into a CUDA graph on the caller's stream and replayed."""
    R, N, T = 20, 1024, 64
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=R)
    env = BatchedHeistEnv(cfg, N)
    am, cp = synthetic.make_valid_workload(env, 1, 15)
    acts = torch.as_tensor(synthetic.sample_actions(np.random.default_rng(2), T, N)).cuda()
    out = {"reward": torch.empty((T, N), device="cuda"), "done": torch.empty((T, N), dtype=torch.uint8, device="cuda"),
           "status": torch.empty((T, N), dtype=torch.uint8, device="cuda"),
           "vis_bits": torch.empty((T, N, R, 1), dtype=torch.int32, device="cuda")}
    env.reset()
    ref = {k: v.clone() for k, v in env.step_many(acts, autoreset=True, out=out).items()}   # also sizes the buffers
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            env.step_many(acts, autoreset=True, out=out)
    torch.cuda.synchronize()
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    for v in out.values():
        v.zero_()
    g.replay()
    torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(out[k], ref[k]), k


@pytest.mark.parametrize("T,autoreset,mode", [(100, True, 0), (20, True, 0), (70, False, 0), (70, True, 2)])
def test_step_many_host_matches_step_many(T, autoreset, mode):
    """Host-buffer rollouts (copies riding the pipelined launch, or bracketing the sequential / ray-march ones)."""
    N = 512
    cfg = EnvironmentConfig(max_steps=40)
    env = BatchedHeistEnv(cfg, N)
    env.set_mode(mode)
    rng = np.random.default_rng(11)
    am, cp = synthetic.sample_asset_maps(rng, N, 20, 20), synthetic.sample_cam_params(rng, N)
    acts = torch.as_tensor(synthetic.sample_actions(rng, T, N))
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    ref = {k: v.cpu() for k, v in env.step_many(acts, autoreset=autoreset, want_vis=True).items()}
    ref_dyn = env.env_dyn.clone()
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    vis = torch.empty((T, N, env.R, env.W), dtype=torch.int32, device="cuda")
    out = env.step_many_host(acts.pin_memory(), autoreset=autoreset, vis_out=vis)
    torch.cuda.synchronize()
    for k in ("reward", "done", "status"):
        assert torch.equal(out[k], ref[k]), k
    assert torch.equal(vis.cpu(), ref["vis_bits"]) and torch.equal(env.env_dyn, ref_dyn)
    env.check_errors()


def test_cache_coverage_and_fallback_mix():
    """Assets outside the cache's range (vision_range > 7, fov > 180, more than 4 guards) leave their env to the
    ray-march kernel; both kernels then serve one batch.  HEIST_NO_VIS_CACHE=1 disables the cache altogether."""
    import os
    cfg = EnvironmentConfig(max_steps=30)
    N = 64
    env = BatchedHeistEnv(cfg, N, max_cams=8, max_guards=8, max_path=8)
    lays = []
    for i in range(N):
        cams = [{"row": 5, "col": 5 + (i % 7), "fov_angle": 60.0 + i, "heading": 10.0 * i, "rotation_speed": 7.5, "vision_range": 6}]
        if i % 4 == 1:
            cams.append({"row": 12, "col": 12, "fov_angle": 90.0, "heading": 0.0, "rotation_speed": 15.0, "vision_range": 9})
        if i % 4 == 2:
            cams.append({"row": 12, "col": 12, "fov_angle": 200.0, "heading": 0.0, "rotation_speed": 15.0, "vision_range": 3})
        guards = [{"patrol_path": [(15, 3), (15, 4), (15, 5), (14, 5)], "speed": 1, "vision_range": 4, "fov_angle": 90.0}]
        if i % 4 == 3:
            guards = guards * 5
        lays.append(([(8, 8), (8, 9)], cams, guards))
    env.set_layout_explicit(lays, budget=np.full(N, 100, np.int32))
    cached, nbytes = env.cache_stats()
    assert cached == N // 4 and nbytes > 0
    oenvs = []
    for w, c, g in lays:
        e = ho.OracleEnv(20, 20, max_steps=30, budget=100)
        e.set_layout(w, c, g)
        oenvs.append(e)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(np.random.default_rng(3), 70, N)
    out = env.step_many(acts, autoreset=True, want_vis=True)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    for k in ("done", "status", "reward"):
        assert np.array_equal(out[k].cpu().numpy(), ref[k]), k
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    env.check_errors()
    # the same handle with layouts the cache covers completely (the ray-march launches are then skipped), and back
    for lays2, want in ((lays[::4] * 4, N), (lays, N // 4)):
        env.set_layout_explicit(lays2, budget=np.full(N, 100, np.int32))
        assert env.cache_stats()[0] == want
        oenvs = []
        for w, c, g in lays2:
            e = ho.OracleEnv(20, 20, max_steps=30, budget=100)
            e.set_layout(w, c, g)
            oenvs.append(e)
        env.reset()
        ho.reset_all(oenvs)
        for t in range(3):   # single ticks first: they decide from the asynchronous coverage count
            rew, done, status = env.step(acts[t])
            env.reset(mask=done)
        ref = ho.rollout(oenvs, acts[:3], autoreset=True, want_vis=True)
        assert np.array_equal(u32(env.visibility_bits), ref["vis_bits"][2])
        out = env.step_many(acts[3:], autoreset=True, want_vis=True)
        ref = ho.rollout(oenvs, acts[3:], autoreset=True, want_vis=True)
        for k in ("done", "status", "reward"):
            assert np.array_equal(out[k].cpu().numpy(), ref[k]), k
        assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    env.check_errors()
    os.environ["HEIST_NO_VIS_CACHE"] = "1"
    try:
        env2 = BatchedHeistEnv(cfg, 8)
    finally:
        del os.environ["HEIST_NO_VIS_CACHE"]
    assert env2.cache_stats() == (0, 0)


def test_fused_tick_graph_survives_layout_changes():
    """heist_step_observe is one fused kernel for table-driven envs.  A CUDA graph captured while every env was
    table-driven must stay correct after a set_layout that leaves some envs to the ray-march (the captured launch
    carries the ray-march and its observe pass, which exit at once when no env needs them)."""
    cfg = EnvironmentConfig(max_steps=25)
    N, T = 32, 40
    env = BatchedHeistEnv(cfg, N, max_cams=8, max_guards=8, max_path=8)

    def layouts(mixed):
        lays = []
        for i in range(N):
            cams = [{"row": 5, "col": 5 + (i % 7), "fov_angle": 60.0 + i, "heading": 15.0 * i, "rotation_speed": 15.0, "vision_range": 6}]
            if mixed and i % 3 == 1:   # beyond the cache's range: this env is ray-marched
                cams.append({"row": 12, "col": 12, "fov_angle": 90.0, "heading": 0.0, "rotation_speed": 15.0, "vision_range": 9})
            guards = [{"patrol_path": [(15, 3), (15, 4), (15, 5), (14, 5)], "speed": 1, "vision_range": 4, "fov_angle": 90.0}]
            lays.append(([(8, 8), (8, 9)], cams, guards))
        return lays

    acts = torch.as_tensor(synthetic.sample_actions(np.random.default_rng(11), T, N)).cuda()
    a_static = torch.zeros(N, dtype=torch.int8, device="cuda")
    state = torch.empty((N, 3, 20, 20), dtype=torch.float32, device="cuda")
    env.set_layout_explicit(layouts(False), budget=np.full(N, 100, np.int32))
    assert env.cache_stats()[0] == N
    env.reset()
    for _ in range(3):   # warm-up: resolves the asynchronous coverage count
        env.step_observe(a_static, autoreset=True, state_out=state)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rew, done, status, _ = env.step_observe(a_static, autoreset=True, state_out=state)
    for mixed in (False, True, False):
        lays = layouts(mixed)
        env.set_layout_explicit(lays, budget=np.full(N, 100, np.int32))
        assert env.cache_stats()[0] == (N - len([i for i in range(N) if i % 3 == 1]) if mixed else N)
        env.reset()
        oenvs = []
        for w, c, gd in lays:
            e = ho.OracleEnv(20, 20, max_steps=25, budget=100)
            e.set_layout(w, c, gd)
            oenvs.append(e)
        ho.reset_all(oenvs)
        for t in range(T):
            a_static.copy_(acts[t])
            g.replay()
            ref = ho.rollout(oenvs, acts[t:t + 1].cpu().numpy(), autoreset=True, want_vis=True)
            assert np.array_equal(rew.cpu().numpy(), ref["reward"][0]) and np.array_equal(status.cpu().numpy(), ref["status"][0]), (mixed, t)
            assert np.array_equal(u32(env.visibility_bits), ref["vis_bits"][0]), (mixed, t)
            assert np.array_equal(state.cpu().numpy(), np.stack([e.state_tensor() for e in oenvs])), (mixed, t)
    env.check_errors()


def test_tables_only_mode_is_one_kernel_per_tick_and_reports_uncovered_layouts():
    """HEIST_MODE_TABLES: the caller guarantees cache coverage, a tick is one kernel even inside a CUDA graph; a layout
    that breaks the guarantee is a reported (sticky) error, not a silent divergence."""
    cfg = EnvironmentConfig(max_steps=30)
    N, T = 64, 45
    env = BatchedHeistEnv(cfg, N)
    env.set_mode(env.MODE_TABLES)
    rng = np.random.default_rng(99)
    am, cp = synthetic.sample_asset_maps(rng, N, 20, 20), synthetic.sample_cam_params(rng, N)
    env.set_layout_from_asset_map(am, cp, 15)
    env.check_errors()
    oenvs, _ = oracle_envs(am, cp, cfg, 15)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, T, N)
    a_static = torch.zeros(N, dtype=torch.int8, device="cuda")
    state = torch.empty((N, 3, 20, 20), dtype=torch.float32, device="cuda")
    n0 = env.launch_count()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.cuda.graph(g, stream=side):
        rew, done, status, _ = env.step_observe(a_static, autoreset=True, state_out=state)
    assert env.launch_count() - n0 == 1   # one kernel captured
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    for t in range(T):
        a_static.copy_(torch.as_tensor(acts[t]))
        g.replay()
        assert np.array_equal(rew.cpu().numpy(), ref["reward"][t]) and np.array_equal(status.cpu().numpy(), ref["status"][t]), t
        assert np.array_equal(u32(env.visibility_bits), ref["vis_bits"][t]), t
    assert np.array_equal(state.cpu().numpy(), np.stack([e.state_tensor() for e in oenvs]))
    # a camera beyond the cache's range breaks the guarantee: reported
    lays = [([], [{"row": 5, "col": 5, "fov_angle": 60.0, "heading": 0.0, "rotation_speed": 15.0, "vision_range": 9}], [])] * N
    env.set_layout_explicit(lays, budget=np.full(N, 100, np.int32))
    with pytest.raises(RuntimeError, match="not covered"):
        env.check_errors()
    env.close()


def test_debug_bounds_build_reports_no_out_of_range_access():
    """compute-sanitizer is closed on the GPU pool, so the march is also run from a -DHEIST_DEBUG_BOUNDS build that
    range-checks every cell-map access (tests/sanitizer_small.py: four grid classes, resets, exact-path rays)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, HEIST_B200_DEBUG="1")
    res = subprocess.run([sys.executable, os.path.join(root, "tests", "sanitizer_small.py")], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 4


def _wild_layout(rng, R, C):
    """Explicit layouts far outside the Architect's ranges: wide/narrow fov, long ranges, negative and
    fractional speeds, many cameras, long custom patrols with strides, guards on walls/borders."""
    walls = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(int(rng.integers(0, 40)))]
    cams = []
    for _ in range(int(rng.integers(0, 13))):
        kind = rng.integers(0, 4)
        fov = [float(rng.uniform(1, 359)), float(rng.choice([10, 15, 180, 270, 359, 360])), float(np.float32(rng.uniform(30, 120))),
               float(rng.choice([60, 90, 120]))][kind]
        cams.append({"row": int(rng.integers(-1, R + 1)), "col": int(rng.integers(-1, C + 1)), "fov_angle": fov,
                     "heading": float(rng.choice([0, 90, 45.5, -30, 725.25, float(rng.uniform(-400, 800))])),
                     "rotation_speed": float(rng.choice([0, 15, -15, 0.1, 359.9, -720.5, float(rng.uniform(-50, 50))])),
                     "vision_range": int(rng.choice([0, 1, 2, 3, 6, 6, 9, 13]))})
    guards = []
    for _ in range(int(rng.integers(0, 5))):
        L = int(rng.integers(1, 17))
        path = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(L)]
        guards.append({"patrol_path": path, "speed": int(rng.choice([1, 1, 2, 3, 5, -1, -2, 0])),
                       "vision_range": int(rng.choice([0, 1, 4, 4, 7, 10])),
                       "fov_angle": float(rng.choice([90.0, 90.0, 45.0, 200.0, 13.7, 360.0]))})
    return walls, cams, guards


@pytest.mark.parametrize("R,C,N,T,seed", [(20, 20, 96, 60, 1), (33, 47, 48, 40, 2), (64, 64, 32, 30, 3), (7, 64, 32, 40, 4),
                                           (64, 5, 32, 40, 5)])
def test_wild_explicit_layouts_match_oracle(R, C, N, T, seed):
    rng = np.random.default_rng(1000 + seed)
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=17, start_pos=(1, 1))
    env = BatchedHeistEnv(cfg, N, max_walls=64, max_cams=16, max_guards=8, max_path=16)
    lays = [_wild_layout(rng, R, C) for _ in range(N)]
    budgets = rng.choice([3, 15, 40, 200], N).astype(np.int32)
    valid = env.set_layout_explicit(lays, budget=budgets).cpu().numpy()
    env.check_errors()
    oenvs = []
    for (w, c, g), b in zip(lays, budgets):
        e = ho.OracleEnv(R, C, max_steps=17, budget=int(b))
        oenvs.append(e)
        assert e.set_layout(w, c, g) == valid[len(oenvs) - 1], len(oenvs) - 1
    assert np.array_equal(env.tile_codes.cpu().numpy(), np.stack([e.grid for e in oenvs]).astype(np.uint8))
    assert env.budget_spent.cpu().tolist() == [e.info()["spent"] for e in oenvs]
    env.reset()
    ho.reset_all(oenvs)
    assert np.array_equal(u32(env.visibility_bits), np.stack([ho.pack_bits(e.visibility) for e in oenvs]))
    acts = synthetic.sample_actions(rng, T, N)
    res = []
    for exact in (False, True):
        env.set_exact_only(exact)
        env.set_layout_explicit(lays, budget=budgets)
        env.reset()
        out = env.step_many(acts, autoreset=True, want_vis=True)
        res.append({k: v.clone() for k, v in out.items()})
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    for out in res:
        assert np.array_equal(out["done"].cpu().numpy(), ref["done"])
        assert np.array_equal(out["status"].cpu().numpy(), ref["status"])
        assert np.array_equal(out["reward"].cpu().numpy(), ref["reward"])
        assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    info = [e.info() for e in oenvs]
    ch, gh, gx = env.cam_heading.cpu().numpy(), env.guard_heading.cpu().numpy(), env.guard_idx.cpu().numpy()
    for j, e in enumerate(oenvs):
        assert np.array_equal(ch[j, :info[j]["n_cams"]], e.cam_headings()), j
        gs, ghead = e.guards_state()
        assert np.array_equal(gx[j, :info[j]["n_guards"]], gs[:, 2]) and np.array_equal(gh[j, :info[j]["n_guards"]], ghead), j
    assert np.array_equal(env.observe().cpu().numpy(), np.stack([e.state_tensor() for e in oenvs]))


@pytest.mark.parametrize("R,C,T", [(20, 20, 200), (40, 64, 136)])
def test_guard_heavy_rollout_on_the_pipelined_path(R, C, T):
    """Up to four guards per env with wide cones that overlap each other and the camera cones, all table-driven, T
    ticks through the pipelined launch: the guards are OR-ed into the finished camera rows with atomics (k_finish_or,
    one- and two-word rows); maps, rewards and final guard states equal the oracle's."""
    rng = np.random.default_rng(4242 + R)
    N = 96
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=23, start_pos=(1, 1))
    env = BatchedHeistEnv(cfg, N, max_walls=32, max_cams=4 if R == 20 else 3, max_guards=4, max_path=12)   # (odd capacity: shared-memory alignment)
    lays = []
    for _ in range(N):
        walls = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(int(rng.integers(0, 25)))]
        cams = [{"row": int(rng.integers(0, R)), "col": int(rng.integers(0, C)), "fov_angle": float(rng.uniform(30, 120)),
                 "heading": float(rng.uniform(0, 360)), "rotation_speed": float(rng.uniform(5, 35)),
                 "vision_range": int(rng.choice([3, 6, 7]))} for _ in range(int(rng.integers(0, 4)))]
        guards = []
        for _ in range(int(rng.integers(1, 5))):
            r0, c0 = int(rng.integers(2, R - 2)), int(rng.integers(2, C - 2))   # patrols close together: cones overlap
            path = [(min(R - 1, max(0, r0 + int(rng.integers(-2, 3)))), min(C - 1, max(0, c0 + int(rng.integers(-2, 3)))))
                    for _ in range(int(rng.integers(1, 13)))]
            guards.append({"patrol_path": path, "speed": int(rng.choice([1, 1, 2, 3, -1])), "vision_range": int(rng.choice([4, 6, 7])),
                           "fov_angle": float(rng.choice([90.0, 120.0, 180.0, 57.3]))})
        lays.append((walls, cams, guards))
    env.set_layout_explicit(lays, budget=np.full(N, 1000, np.int32))
    env.check_errors()
    assert env.cache_stats()[0] == N   # every env on the table-driven path
    oenvs = []
    for w, c, g in lays:
        e = ho.OracleEnv(R, C, max_steps=23, budget=1000)
        e.set_layout(w, c, g)
        oenvs.append(e)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, T, N)
    out = env.step_many(acts, autoreset=True, want_vis=True)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    env.check_errors()   # (every state the guards went through had its cone built: no ERR_STATE)
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])
    assert np.array_equal(out["status"].cpu().numpy(), ref["status"])
    assert np.array_equal(out["reward"].cpu().numpy(), ref["reward"])
    assert np.array_equal(u32(env.visibility_bits), np.stack([ho.pack_bits(e.visibility) for e in oenvs]))
    gh, gx = env.guard_heading.cpu().numpy(), env.guard_idx.cpu().numpy()
    for j, e in enumerate(oenvs):
        gs, ghead = e.guards_state()
        ng = e.info()["n_guards"]
        assert np.array_equal(gx[j, :ng], gs[:, 2]) and np.array_equal(gh[j, :ng], ghead), j
    env.close()


@pytest.mark.parametrize("N", [1, 2, 3, 5, 7, 130])
def test_env_counts_not_multiple_of_cta(N):
    cfg = EnvironmentConfig(max_steps=40)
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(N)
    am, cp = synthetic.sample_asset_maps(rng, N, 20, 20), synthetic.sample_cam_params(rng, N)
    env.set_layout_from_asset_map(am, cp, 15)
    oenvs, _ = oracle_envs(am, cp, cfg, 15)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, 70, N)
    out = env.step_many(acts, autoreset=True, want_vis=True)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    assert np.array_equal(out["status"].cpu().numpy(), ref["status"])
    assert np.array_equal(out["reward"].cpu().numpy(), ref["reward"])
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"])


def test_errors_are_reported_not_swallowed():
    import ctypes as C
    from heist_b200 import _ffi
    lib = _ffi.load()
    # bad construction parameters -> negative code + message
    p = _ffi.HeistParams(grid_rows=80, grid_cols=20, max_steps=200, start_row=1, start_col=1, vault_row=18, vault_col=18,
                         architect_budget=15, max_walls=8, max_cams=8, max_guards=4, max_path=8,
                         reward_vault=10.0, reward_detection=-1.0, reward_step=-0.01)
    h = C.c_void_p()
    assert lib.heist_create(C.byref(p), 4, 0, C.byref(h)) == -3 and b"outside" in lib.heist_last_error()
    p.grid_rows, p.max_cams, p.max_guards = 20, 30, 8
    assert lib.heist_create(C.byref(p), 4, 0, C.byref(h)) == -5
    with pytest.raises(RuntimeError, match="heist_create"):
        BatchedHeistEnv(EnvironmentConfig(grid_rows=2, grid_cols=2), 1)
    # a budget that can buy more assets than the per-env lists hold is refused up front (the decode would otherwise
    # drop assets the reference keeps): host check in the wrapper and in heist_decode_validate (scalar budget) ...
    env = BatchedHeistEnv(EnvironmentConfig(), 8, max_walls=4)
    am = np.zeros((8, 20, 20), np.int8)
    am[:, 5, 2:12] = 1
    cp8 = np.tile(np.float32([60, 15, 0]), (8, 1))
    with pytest.raises(ValueError, match="capacities"):
        env.set_layout_from_asset_map(am, cp8, budget=15)
    am_d, cp_d = torch.as_tensor(am).cuda(), torch.as_tensor(cp8).cuda()
    assert lib.heist_decode_validate(env._h, C.c_void_p(am_d.data_ptr()), C.c_void_p(cp_d.data_ptr()), None, 1, 1, None, None) == -12
    assert b"capacities" in lib.heist_last_error()
    # ... and with a per-env device budget the kernel raises the sticky flag instead of truncating silently
    bud = torch.full((8,), 15, dtype=torch.int32, device="cuda")
    assert lib.heist_decode_validate(env._h, C.c_void_p(am_d.data_ptr()), C.c_void_p(cp_d.data_ptr()), C.c_void_p(bud.data_ptr()),
                                     1, 1, None, None) == 0
    with pytest.raises(RuntimeError, match="capacity"):
        env.check_errors()
    env.check_errors()  # the flag is cleared once reported
    # host-side validation of explicit layouts
    with pytest.raises(ValueError, match="capacity"):
        env.set_layout_explicit([([(1, 1)] * 5, [], [])] * 8)
    with pytest.raises(ValueError, match="outside"):
        env.set_layout_explicit([([], [], [{"patrol_path": [(25, 3)]}])] * 8)
    with pytest.raises(KeyError):
        heist_b200.HeistEnvironment(EnvironmentConfig(grid_rows=10, grid_cols=10)).step(7)


@pytest.mark.parametrize("speed", [1, 3, -1, 8, 0])
def test_guard_cones_cover_every_reachable_state_and_only_those(speed):
    """The guards' cones are tabulated per (waypoint, heading slot) for the pairs a patrol can reach (k_build_cache):
    strides that skip waypoints, steps that are no move (heading kept), resets at every phase of the patrol -- single
    ticks and rollouts equal the oracle with no state error; a (waypoint, heading) pair written by hand into the state
    that no patrol reaches is reported, not served from an unbuilt table."""
    cfg = EnvironmentConfig(max_steps=7)
    N, T = 32, 70
    env = BatchedHeistEnv(cfg, N, max_path=12)
    rng = np.random.default_rng(99 + speed)
    lays = []
    for i in range(N):
        r0, c0 = int(rng.integers(4, 15)), int(rng.integers(4, 15))
        ring = [(r0 - 1, c0 - 1), (r0 - 1, c0), (r0 - 1, c0 + 1), (r0, c0 + 1), (r0 + 1, c0 + 1), (r0 + 1, c0), (r0 + 1, c0 - 1), (r0, c0 - 1)]
        if i % 3 == 1:
            ring[2] = ring[1]; ring[5] = ring[4]          # no-move steps
        if i % 3 == 2:
            ring = ring[:int(rng.integers(1, 8))] + [(r0 + 2, c0 - 2)]
        lays.append(([(9, 9)], [], [{"patrol_path": ring, "speed": speed, "vision_range": 4, "fov_angle": 90.0}]))
    env.set_layout_explicit(lays, budget=np.full(N, 100, np.int32))
    assert env.cache_stats()[0] == N
    oenvs = []
    for w, c, g in lays:
        e = ho.OracleEnv(20, 20, max_steps=7, budget=100)
        e.set_layout(w, c, g)
        oenvs.append(e)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, T, N)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    for t in range(8):   # single ticks (k_walk) ...
        _, done, _ = env.step(acts[t])
        env.reset(mask=done)
        assert np.array_equal(u32(env.visibility_bits), ref["vis_bits"][t]), t
    out = env.step_many(acts[8:], autoreset=True, want_vis=True)   # ... then the rollout kernels from that state
    env.check_errors()
    assert np.array_equal(u32(out["vis_bits"]), ref["vis_bits"][8:])
    assert np.array_equal(out["status"].cpu().numpy(), ref["status"][8:])
    if speed == 1:   # waypoint 4 of a ring is always entered heading south (270); heading 90 (a slot of the path) never gets there
        env.set_layout_explicit(lays[:1] * N, budget=np.full(N, 100, np.int32))
        env.reset()
        env.guard_idx[:, 0] = 4
        env.guard_heading[:, 0] = 90.0
        env.step(acts[0])
        with pytest.raises(RuntimeError, match="guard state"):
            env.check_errors()
    env.close()


@pytest.mark.parametrize("R,C", [(5, 4), (8, 8), (7, 12), (20, 20), (33, 36), (40, 64), (64, 64), (64, 4), (6, 16)])
def test_fused_tick_dense_state_on_odd_grid_shapes(R, C):
    """heist_step_observe writes the (3, R, C) state from the fused tick kernel with incrementally advanced
    (channel, cell, row, column) indices: grids whose channels are shorter than one warp trip (5x4: 15 float4 in all),
    end inside a trip (7x12, 33x36), or take two lanes / two words per row (40x64, 64x64) -- every tick's state equals
    the oracle's get_state_tensor (environment.py:347-374), with auto-resets on the way."""
    cfg = EnvironmentConfig(grid_rows=R, grid_cols=C, max_steps=9, start_pos=(1, 1), vault_pos=(R - 2, C - 2))
    N, T = 40, 24
    env = BatchedHeistEnv(cfg, N, max_path=8)
    rng = np.random.default_rng(R * 100 + C)
    lays = []
    for _ in range(N):
        walls = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(int(rng.integers(0, 6)))]
        cams = [{"row": int(rng.integers(0, R)), "col": int(rng.integers(0, C)), "fov_angle": float(rng.uniform(30, 120)),
                 "heading": float(rng.uniform(0, 360)), "rotation_speed": float(rng.uniform(5, 35)), "vision_range": int(rng.integers(1, 7))}
                for _ in range(int(rng.integers(0, 3)))]
        path = [(int(rng.integers(0, R)), int(rng.integers(0, C))) for _ in range(int(rng.integers(1, 6)))]
        guards = [{"patrol_path": path, "speed": 1, "vision_range": 3, "fov_angle": 90.0}] if rng.random() < 0.6 else []
        lays.append((walls, cams, guards))
    env.set_layout_explicit(lays, budget=np.full(N, 1000, np.int32))
    env.check_errors()
    assert env.cache_stats()[0] == N
    oenvs = []
    for w, c, g in lays:
        e = ho.OracleEnv(R, C, max_steps=9, budget=1000, start=(1, 1), vault=(R - 2, C - 2))
        e.set_layout(w, c, g)
        oenvs.append(e)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, T, N)
    state = torch.empty(N, 3, R, C, device="cuda")
    for t in range(T):
        rew, done, status, _ = env.step_observe(acts[t], autoreset=True, state_out=state)
        ref = ho.rollout(oenvs, acts[t:t + 1], autoreset=True, want_vis=False)
        assert np.array_equal(status.cpu().numpy(), ref["status"][0]), t
        assert np.array_equal(state.cpu().numpy(), np.stack([e.state_tensor() for e in oenvs])), t
    env.check_errors()
    env.close()


def test_single_tick_api_with_masked_resets_matches_oracle():
    """heist_step + heist_reset(mask) at batch scale, the pattern of a policy-in-the-loop driver."""
    cfg = EnvironmentConfig(max_steps=25)
    N, T = 512, 60
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(2718)
    am, cp = synthetic.sample_asset_maps(rng, N, 20, 20), synthetic.sample_cam_params(rng, N, nice=True)
    env.set_layout_from_asset_map(am, cp, 15)
    oenvs, _ = oracle_envs(am, cp, cfg, 15)
    env.reset()
    ho.reset_all(oenvs)
    acts = synthetic.sample_actions(rng, T, N)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    for t in range(T):
        rew, done, status = env.step(acts[t])
        assert np.array_equal(rew.cpu().numpy(), ref["reward"][t]) and np.array_equal(status.cpu().numpy(), ref["status"][t])
        if t % 7 == 0:  # un-reset state of finished envs is observable too: done flag stays until reset
            assert torch.equal(env.done, done)
        env.reset(mask=done)
        assert np.array_equal(u32(env.visibility_bits), ref["vis_bits"][t]), t
