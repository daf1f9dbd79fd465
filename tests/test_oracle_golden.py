"""Pin the CPU oracle (oracle/heist_oracle.c) against fixtures generated from the reference.

Bit-exact for visibility / positions / done / status / grid / BFS / decode; rewards are
compared as float64 with ==; state tensors as float32 with ==; GAE with == (normalised
advantages at 1e-5 relative, torch's reduction order is not sequential).
"""
import hashlib

import numpy as np
import pytest

from oracle import heist_oracle as ho


def make_env(golden, name):
    t = golden.traces[name]
    walls, cams, guards, budget = golden.layout(name)
    e = ho.OracleEnv(t["R"], t["C"], max_steps=t["max_steps"], budget=budget)
    valid = e.set_layout(walls, cams, guards)
    return e, valid, t


def test_all_traces_bit_exact(golden):
    assert len(golden.traces) >= 40
    _check_traces(golden)


def test_round2_traces_bit_exact(golden2):
    """Headings a few ulp off multiples of 30 degrees (speed 0.1, 1/3), and the 32x32 / 64x64 budget-22 layouts."""
    assert sum(n.startswith("tie") for n in golden2.traces) == 6 and sum(n.startswith("big") for n in golden2.traces) == 8
    assert all(v >= 1 for v in golden2.meta["tie_near_ticks"].values())
    _check_traces(golden2)


def test_round3_patrol_traces_bit_exact(golden3):
    """Guard patrols with strides, no-move steps, one- and two-waypoint paths, up to four guards, short episodes whose
    resets fall on every phase of a patrol (security.py:145-159, environment.py:205-208) -- recorded from the reference."""
    assert sum(n.startswith("patrol") for n in golden3.traces) == 10 and all(t["valid"] for t in golden3.traces.values())
    assert sum(int(golden3.arr(n, "done").sum()) for n in golden3.traces if n.startswith("patrol")) >= 300
    # ... and six grids that are not the usual squares, get_state_tensor recorded after every step
    assert sum(n.startswith("shape") for n in golden3.traces) == 6
    assert all(len(golden3.arr(n, "state_t")) == len(golden3.arr(n, "actions")) for n in golden3.traces if n.startswith("shape"))
    _check_traces(golden3)


def test_trainer_tapes(golden2):
    """The call sequence AdversarialTrainer._run_one_episode made on the reference env (training.py:418-600):
    the oracle reproduces every recorded return value."""
    tr = golden2.meta["trainer"]
    assert len(tr["episodes"]) >= 3 and sum(len(e["tape"]) for e in tr["episodes"]) >= 300
    e, budget = None, 15
    for k, ep in enumerate(tr["episodes"]):
        for i, c in enumerate(ep["tape"]):
            where = (k, i, c["call"])
            arr = lambda key: golden2.z[f"trainer{k}/{i}/{key}"]
            if c["call"] == "scale_budget":
                budget = c["arg"]
            elif c["call"] == "set_layout":
                e = ho.OracleEnv(tr["R"], tr["C"], max_steps=tr["max_steps"], budget=budget)
                guards = [{**g, "patrol_path": [tuple(p) for p in g["patrol_path"]]} for g in c["guards"]]
                assert e.set_layout([tuple(w) for w in c["walls"]], c["cameras"], guards) == c["ret"], where
                info = e.info()
                assert info["spent"] == c["spent"] and [info["n_walls"], info["n_cams"], info["n_guards"]] == c["n_placed"], where
            elif c["call"] == "is_level_valid":
                assert e.is_level_valid() == c["ret"], where
            elif c["call"] in ("reset", "step"):
                if c["call"] == "reset":
                    e.reset()
                else:
                    r, d, st = e.step(c["arg"])
                    assert (r, d) == (c["reward"], c["done"]), where
                    assert ["running", "detected", "vault_reached", "timeout", "already_done"][st] == c["info"]["status"], where
                    assert e.info()["tick"] == c["tick_after"], where
                s = e.state_tensor()
                assert np.array_equal(s[0], arr("occ")) and np.array_equal(s[1], arr("vis")), where
                assert np.array_equal(np.concatenate(e.obs_vectors()), arr("vec")), where
            elif c["call"] == "get_state_tensor":
                assert np.array_equal(e.state_tensor(), arr("state")), where
            elif c["call"] == "get_environment_state":
                st = c["ret"]
                info = e.info()
                assert [info["solver_r"], info["solver_c"]] == st["solver_pos"] and info["tick"] == st["tick"], where
                assert np.array_equal(e.grid, np.asarray(st["grid"])), where
                assert np.array_equal(e.visibility, np.asarray(st["visibility"], np.float32)), where
                assert np.array_equal(e.cam_headings(), [cam["heading"] for cam in st["cameras"]]), where


def _check_traces(golden):
    for name in golden.traces:
        e, valid, t = make_env(golden, name)
        assert valid == t["valid"], name
        info = e.info()
        assert info["spent"] == t["spent"], name
        assert [info["n_walls"], info["n_cams"], info["n_guards"]] == t["n_placed"], name
        assert np.array_equal(e.grid.astype(np.int8), golden.arr(name, "grid")), name
        e.reset()
        assert np.array_equal(ho.pack_bits(e.visibility), golden.arr(name, "vis0")), name
        acts = golden.arr(name, "actions")
        g = {k: golden.arr(name, k) for k in ["reward", "done", "status", "pos", "tick", "vis", "vis_post",
                                              "cam_heading", "guard_idx", "guard_heading", "state_t", "state",
                                              "obs_vec"]}
        si = 0
        for ti, a in enumerate(acts):
            r, d, st = e.step(int(a))
            info = e.info()
            assert r == g["reward"][ti], (name, ti, r, g["reward"][ti])
            assert d == bool(g["done"][ti]) and st == g["status"][ti], (name, ti)
            assert (info["solver_r"], info["solver_c"]) == tuple(g["pos"][ti]), (name, ti)
            assert info["tick"] == g["tick"][ti], (name, ti)
            assert np.array_equal(ho.pack_bits(e.visibility), g["vis"][ti]), (name, ti)
            nc, ng = info["n_cams"], info["n_guards"]
            assert np.array_equal(e.cam_headings(), g["cam_heading"][ti, :nc]), (name, ti)
            gs, gh = e.guards_state()
            assert np.array_equal(gs[:, 2], g["guard_idx"][ti, :ng]), (name, ti)
            assert np.array_equal(gh, g["guard_heading"][ti, :ng]), (name, ti)
            if si < len(g["state_t"]) and g["state_t"][si] == ti:
                assert np.array_equal(e.state_tensor(), g["state"][si]), (name, ti)
                assert np.array_equal(np.concatenate(e.obs_vectors()), g["obs_vec"][si]), (name, ti)
                si += 1
            if d:
                e.reset()
            assert np.array_equal(ho.pack_bits(e.visibility), g["vis_post"][ti]), (name, ti)


def test_reference_smoke_script_known_answers(golden):
    # test_sanity.py:21-41 and test_fixes.py:13-37 printed values
    e, valid, _ = make_env(golden, "sanity")
    assert valid and e.info()["spent"] == 11
    e.reset()
    assert int(e.visibility.sum()) == 23
    rows = [int("".join("1" if v > 0.5 else "0" for v in row), 2) for row in e.visibility]
    assert rows == [0, 0, 0, 0x2, 0xE, 0x3E, 0x7E, 0xF2, 0x70, 0]
    out = [e.step(4) for _ in range(5)]
    assert [o[0] for o in out] == [0.09000000000000001] * 4 + [-0.91]
    assert out[-1][1] and out[-1][2] == 1
    info = e.info()
    assert (info["solver_r"], info["solver_c"], info["tick"]) == (1, 6, 5)
    assert int(e.visibility.sum()) == 21
    e, _, _ = make_env(golden, "fixes")
    e.reset()
    tot = sum(e.step(2)[0] for _ in range(7))
    assert f"{tot:+.3f}" == "+0.630"
    res = [e.step(4) for _ in range(7)]
    tot += sum(r[0] for r in res)
    assert f"{tot:+.3f}" == "+11.560" and res[-1][2] == 2
    e.reset()
    s = e.state_tensor()
    assert f"{s[2].min():.3f}" == "-1.000" and f"{s[2].max():.3f}" == "0.790"


def test_survey_kat20_hash(golden):
    e, valid, _ = make_env(golden, "kat20")
    assert valid and e.info()["spent"] == 16
    e.reset()
    h, tot, episodes = hashlib.sha256(), 0, 0
    for _ in range(400):
        v = e.visibility > 0.5
        h.update(np.packbits(v).tobytes())
        tot += int(v.sum())
        _, d, _ = e.step(0)
        if d:
            episodes += 1
            e.reset()
    assert h.hexdigest() == golden.meta["kat20_sha256"] == \
        "513d958ef974b21eade9aa03015912517dece7e4b253561dd48067fbad01e8b6"
    assert tot == golden.meta["kat20_sum"] == 28938 and episodes == 2
    assert e.cam_headings().tolist() == golden.meta["kat20_final_headings"] == [21.5, 80.0]


def test_decode_and_set_layout(golden):
    for k, d in enumerate(golden.meta["decode"]):
        am = golden.z[f"decode{k}/asset_map"]
        fov, speed, heading = d["params"]
        walls, cams, guards, _ = ho.decode_layout(am, d["budget"], fov, speed, heading)
        assert [list(w) for w in walls] == d["walls"], k
        assert cams == d["cameras"], k
        assert [[list(p) for p in g["patrol_path"]] for g in guards] == d["guards"], k
        e = ho.OracleEnv(d["H"], d["W"], budget=d["budget"])
        valid = e.set_layout(walls, cams if d["allow_cameras"] else [], guards if d["allow_guards"] else [])
        assert valid == d["valid"] and e.info()["spent"] == d["spent"], k
        assert np.array_equal(e.grid.astype(np.int8), golden.z[f"decode{k}/grid"]), k
        e.reset()
        assert np.array_equal(ho.pack_bits(e.visibility), golden.z[f"decode{k}/vis0"]), k


def test_bfs(golden):
    grids, dims, ans = golden.z["bfs/grids"], golden.z["bfs/dims"], golden.z["bfs/answer"]
    assert 0 < ans.sum() < len(ans)
    for g, (R, C, sr, sc, gr, gc), a in zip(grids, dims, ans):
        assert ho.bfs(g[:R, :C].astype(np.int32), (sr, sc), (gr, gc)) == bool(a)


def test_gae_returns(golden):
    for k in range(golden.meta["gae_cases"]):
        g = {key: golden.z[f"gae{k}/{key}"] for key in ["rew", "val", "done", "adv", "ret", "norm"]}
        adv, ret = ho.gae(g["rew"], g["val"], g["done"])
        assert np.array_equal(adv, g["adv"]), k      # bit-exact fp32 in the reference's op order
        assert np.array_equal(ret, g["ret"]), k
        np.testing.assert_allclose(ho.normalize(adv), g["norm"], rtol=1e-5, atol=1e-6)


def test_gae_columns_are_independent():
    rng = np.random.default_rng(3)
    rew, val = rng.normal(size=(50, 4)).astype(np.float32), rng.normal(size=(50, 4)).astype(np.float32)
    dn = (rng.random((50, 4)) < 0.1).astype(np.float32)
    adv, ret = ho.gae(rew, val, dn)
    for j in range(4):
        a1, r1 = ho.gae(rew[:, j].copy(), val[:, j].copy(), dn[:, j].copy())
        assert np.array_equal(a1, adv[:, j]) and np.array_equal(r1, ret[:, j])


def test_architect_reward(golden):
    for s, rv, ri in zip(golden.z["arch/solve_rate"], golden.z["arch/reward_valid"], golden.z["arch/reward_invalid"]):
        assert ho.architect_reward(True, s) == rv and ho.architect_reward(False, s) == ri


def test_batch_rollout_matches_single_env_loop(golden):
    names = [n for n in golden.traces if n.startswith("rand20x20")][:6]
    envs, singles = [], []
    for n in names:
        e, _, _ = make_env(golden, n)
        e2, _, _ = make_env(golden, n)
        envs.append(e)
        singles.append(e2)
    T = 48
    acts = np.stack([golden.arr(n, "actions")[:T] for n in names], axis=1)
    ho.reset_all(envs, n_threads=2)
    out = ho.rollout(envs, acts, autoreset=True, want_vis=True, n_threads=3)
    for j, n in enumerate(names):
        assert np.array_equal(out["reward64"][:, j], golden.arr(n, "reward")[:T])
        assert np.array_equal(out["done"][:, j], golden.arr(n, "done")[:T])
        assert np.array_equal(out["status"][:, j], golden.arr(n, "status")[:T])
        assert np.array_equal(out["vis_bits"][:, j], golden.arr(n, "vis_post")[:T])
    assert out["live"] == T * len(names)
