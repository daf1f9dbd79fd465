"""GPU tests of the callers either side of the env path: on-device Architect sampling, policy-in-loop
rollouts into a packed buffer, minibatch re-expansion, PPO update with the gradient all-reduce hook."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import heist_b200  # noqa: E402
from heist_b200 import BatchedHeistEnv, EnvironmentConfig, ppo, synthetic  # noqa: E402
from oracle import heist_oracle as ho  # noqa: E402


class TinyPolicy(torch.nn.Module):
    """Stand-in for SolverNetwork (same forward contract, networks.py:65-131)."""

    def __init__(self, rows, cols):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 8, 3, padding=1)
        self.pool = torch.nn.AdaptiveAvgPool2d(4)
        self.pi = torch.nn.Linear(128, 5)
        self.v = torch.nn.Linear(128, 1)

    def forward(self, state, hidden=None):
        x = self.pool(torch.relu(self.conv(state))).flatten(1)
        return self.pi(x), self.v(x), hidden


class FixedPolicy(torch.nn.Module):
    """Deterministic: one-hot logits from a fixed hash of the state, so sampling cannot diverge."""

    def forward(self, state):
        key = (state[:, 1].sum((1, 2)) * 7 + state[:, 2].argmax(-1).sum(-1)).long() % 5
        return torch.nn.functional.one_hot(key, 5).float() * 1e4 - 5e3, torch.zeros(state.shape[0], device=state.device)


def _env(N=96, T=40, seed=3):
    cfg = EnvironmentConfig(max_steps=30)
    env = BatchedHeistEnv(cfg, N)
    rng = np.random.default_rng(seed)
    am, cp = synthetic.sample_asset_maps(rng, N, 20, 20), synthetic.sample_cam_params(rng, N)
    env.set_layout_from_asset_map(am, cp, 15)
    env.reset()
    return env, cfg, am, cp


def test_packed_rollout_reexpands_to_the_observed_states():
    env, _, _, _ = _env()
    buf = ppo.PackedRollout(env, 40)
    dense = []
    orig = buf.record_state

    def spy():
        dense.append(env.observe().clone())
        orig()

    buf.record_state = spy
    torch.manual_seed(0)
    ppo.collect_rollout(env, TinyPolicy(20, 20).cuda(), buf)
    idx = torch.arange(40 * env.num_envs, device="cuda")
    assert torch.equal(buf.states(idx), torch.cat(dense))
    assert buf.dones.any() and buf.t == 40


def test_policy_in_loop_rollout_matches_oracle():
    env, cfg, am, cp = _env(N=64)
    buf = ppo.PackedRollout(env, 50)
    state, _, stats = ppo.collect_rollout(env, FixedPolicy(), buf)
    acts = buf.actions.cpu().numpy().astype(np.int8)
    oenvs = []
    for i in range(64):
        walls, cams, guards, _ = ho.decode_layout(am[i], 15, *cp[i])
        e = ho.OracleEnv(20, 20, max_steps=30)
        e.set_layout(walls, cams, guards)
        oenvs.append(e)
    ho.reset_all(oenvs)
    ref = ho.rollout(oenvs, acts, autoreset=True, want_vis=True)
    assert np.array_equal(buf.rewards.cpu().numpy(), ref["reward"])
    assert np.array_equal(buf.dones.cpu().numpy(), ref["done"])
    assert np.array_equal(state.cpu().numpy(), np.stack([e.state_tensor() for e in oenvs]))
    st = ref["status"]
    assert [int(stats[k]) for k in ("vault", "detected", "timeout")] == [(st == 2).sum(), (st == 1).sum(), (st == 3).sum()]
    assert len(np.unique(acts)) > 1


def test_ppo_update_steps_the_policy():
    env, _, _, _ = _env(N=128)
    policy = TinyPolicy(20, 20).cuda()
    opt = torch.optim.Adam(policy.parameters(), lr=1e-3)
    buf = ppo.PackedRollout(env, 32)
    before = [p.detach().clone() for p in policy.parameters()]
    torch.manual_seed(1)
    for _ in range(2):
        ppo.collect_rollout(env, policy, buf)
        m = ppo.ppo_update(policy, opt, buf, epochs=2, minibatch=1024)
        assert all(torch.isfinite(m[k]) for k in ("solver_policy_loss", "solver_value_loss", "solver_entropy"))
        assert m["updates"] == 2 * 4 and buf.t == 0
    assert any(not torch.equal(a, b) for a, b in zip(before, policy.parameters()))


def test_architect_sampling_feeds_the_decode_kernel(golden):
    d = golden.meta["decode"][2]
    am = torch.from_numpy(golden.z["decode2/asset_map"].astype(np.int64)).cuda()
    logits = torch.full((1, 4, d["H"], d["W"]), -1e4, device="cuda")
    logits.scatter_(1, am.view(1, 1, d["H"], d["W"]), 1e4)
    sampled, logp = ppo.architect_sample(logits, temperature=1.3)
    assert torch.equal(sampled[0].long(), am) and abs(logp.item()) < 1e-3
    env = BatchedHeistEnv(EnvironmentConfig(grid_rows=d["H"], grid_cols=d["W"]), 1, max_guards=8)
    cp = ppo.camera_params_tensor({"fov": torch.tensor([[d["params"][0]]]), "speed": torch.tensor([[d["params"][1]]]),
                                   "heading": torch.tensor([[d["params"][2]]])})
    valid = env.set_layout_from_asset_map(sampled, cp, d["budget"], d["allow_cameras"], d["allow_guards"])
    assert bool(valid.item()) == d["valid"]
    assert np.array_equal(env.tile_codes[0].cpu().numpy().astype(np.int8), golden.z["decode2/grid"])
    # statistical sanity of the sampler itself: empirical frequencies follow softmax(logits / T)
    lg = torch.tensor([0.0, 1.0, 2.0, -1.0], device="cuda").view(1, 4, 1, 1).expand(1, 4, 200, 200).contiguous()
    s, _ = ppo.architect_sample(lg, temperature=2.0)
    freq = torch.bincount(s.flatten().long(), minlength=4).float() / s.numel()
    assert torch.allclose(freq, torch.softmax(torch.tensor([0.0, 1.0, 2.0, -1.0]) / 2.0, 0).cuda(), atol=0.01)


def test_calls_are_cuda_graph_capturable():
    """No host sync or allocation inside the ABI calls: a tick + observe + GAE replays from a CUDA graph."""
    env, _, _, _ = _env(N=64)
    acts = torch.randint(0, 5, (64,), dtype=torch.int8, device="cuda")
    state = torch.empty(64, 3, 20, 20, device="cuda")
    snap = {k: getattr(env, k).clone() for k in ("env_dyn", "visibility_bits", "cam_heading", "guard_heading", "guard_idx")}
    r0, d0, s0, st0 = env.step_observe(acts, True)
    for k, v in snap.items():               # rewind, then run the same tick from a graph
        getattr(env, k).copy_(v)
    rew = torch.zeros(1, 64, device="cuda")
    done = torch.zeros(1, 64, dtype=torch.uint8, device="cuda")
    stat = torch.zeros(1, 64, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            env.step_many(acts.view(1, 64), autoreset=True, out={"reward": rew, "done": done, "status": stat})
            env.observe(out=state)
            adv, ret = heist_b200.compute_gae(rew, torch.zeros_like(rew), done)
    for k, v in snap.items():
        getattr(env, k).copy_(v)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(rew[0], r0) and torch.equal(done[0].bool(), d0) and torch.equal(stat[0], s0)
    assert torch.equal(state, st0) and torch.equal(ret, rew)
