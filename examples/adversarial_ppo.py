#!/usr/bin/env python
"""BASELINE config 5 in batched form: Architect layout generation -> Solver rollouts (policy in the loop) -> GAE ->
PPO update with the gradient all-reduce, one process per GPU.

    python examples/adversarial_ppo.py --iters 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/adversarial_ppo.py

The networks are plain PyTorch (out of scope of the B200 rebuild): with the reference on sys.path
(`--reference /path/to/reference`) its own SolverNetwork / ArchitectNetwork are used unchanged; otherwise small
stand-ins with the same forward contracts.  Mirrors AdversarialTrainer._run_one_episode (training.py:418-600):
curriculum budget -> generate_layout -> curriculum filter -> set_layout -> invalid layouts get -1 ->
`solver_episodes` attempts per layout -> architect reward from the solve rate -> both agents update.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heist_b200  # noqa: E402
from heist_b200 import ppo  # noqa: E402


class SmallSolver(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(3, 32, 3, padding=1), nn.Conv2d(32, 64, 3, padding=1)
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.fc = nn.Linear(1024, 256)
        self.pi, self.v = nn.Linear(256, 5), nn.Linear(256, 1)

    def forward(self, x, hidden=None):
        x = F.relu(self.fc(self.pool(F.relu(self.c2(F.relu(self.c1(x))))).flatten(1)))
        return self.pi(x), self.v(x), hidden


class SmallArchitect(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = nn.Sequential(nn.Conv2d(1, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU())
        self.dec = nn.Conv2d(32, 4, 1)
        self.glob = nn.Linear(32, 64)
        self.value, self.cam = nn.Linear(64, 1), nn.Linear(64, 3)

    def forward(self, grid):
        f = self.enc(grid)
        g = F.relu(self.glob(f.mean((2, 3))))
        s = torch.sigmoid(self.cam(g))
        params = {"fov": s[:, 0:1] * 90 + 30, "speed": s[:, 1:2] * 30 + 5, "heading": s[:, 2:3] * 360}  # networks.py:232-236
        return self.dec(f), self.value(g), params


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024, help="layouts (envs) per GPU")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ticks", type=int, default=64)
    ap.add_argument("--budget", type=int, default=15)
    ap.add_argument("--reference", default=None)
    a = ap.parse_args()
    rank, world, local = heist_b200.dist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    torch.manual_seed(1234 + rank)
    cfg = heist_b200.EnvironmentConfig()
    if a.reference:
        sys.path.insert(0, a.reference)
        from heist_architect.networks import ArchitectNetwork, SolverNetwork
        solver, architect = SolverNetwork(20, 20, 5).to(dev), ArchitectNetwork(20, 20).to(dev)
    else:
        solver, architect = SmallSolver().to(dev), SmallArchitect().to(dev)
    for net in (solver, architect):  # same initial weights on every rank
        for p in net.parameters():
            if world > 1:
                torch.distributed.broadcast(p.data, 0)
    opt_s = torch.optim.Adam(solver.parameters(), lr=1e-3)
    opt_a = torch.optim.Adam(architect.parameters(), lr=3e-4)
    env = heist_b200.BatchedHeistEnv(cfg, a.envs, device=dev)
    buf = ppo.PackedRollout(env, a.ticks)
    base = torch.zeros((a.envs, 1, 20, 20), device=dev)
    base[:, 0, 0, :] = base[:, 0, -1, :] = base[:, 0, :, 0] = base[:, 0, :, -1] = 0.2
    base[:, 0, 1, 1], base[:, 0, 18, 18] = 0.4, 0.6
    for it in range(a.iters):
        t0 = time.time()
        # --- Architect: sample -> decode -> validate, all on device (networks.py:241-322, training.py:456-470) ---
        logits, a_value, cam = architect(base)
        asset_map, a_logp = ppo.architect_sample(logits.detach(), temperature=max(0.5, 2.0 - 1.5 * it / max(a.iters, 1)))
        valid = env.set_layout_from_asset_map(asset_map, ppo.camera_params_tensor(cam).detach(), a.budget)
        env.reset()
        # --- Solver: rollouts with the policy in the loop, packed buffer, GAE, PPO (training.py:515-562) ---
        _, _, stats = ppo.collect_rollout(env, solver, buf)
        m = ppo.ppo_update(solver, opt_s, buf, epochs=3, minibatch=8192)
        # --- Architect reward from per-layout solve rates (rewards.py:43-73), value-only update as in the reference
        #     (architect.py:105-141: the policy term is detached there) ---
        a_rew, solve_rate = env.architect_reward()
        a_loss = F.mse_loss(a_value.view(-1), a_rew.float())
        opt_a.zero_grad(set_to_none=True)
        a_loss.backward()
        heist_b200.dist.allreduce_gradients(architect.parameters())
        nn.utils.clip_grad_norm_(architect.parameters(), 0.5)
        opt_a.step()
        cnt = heist_b200.dist.allreduce_sum(torch.stack([valid.sum(), stats["vault"], stats["detected"], stats["timeout"]]))
        torch.cuda.synchronize()
        if rank == 0:
            steps = a.ticks * a.envs * world
            print(f"iter {it}: {steps} env-steps on {world} GPU(s) in {time.time() - t0:.2f}s | valid layouts {int(cnt[0])}"
                  f" | episodes vault/detected/timeout {int(cnt[1])}/{int(cnt[2])}/{int(cnt[3])} | solver loss "
                  f"pi {m['solver_policy_loss']:.4f} v {m['solver_value_loss']:.4f} H {m['solver_entropy']:.3f} | "
                  f"architect reward {a_rew.mean():.3f}", flush=True)
    if world > 1:
        # every rank must hold identical weights after identical all-reduced updates
        w = torch.cat([p.detach().flatten() for p in solver.parameters()])
        ref = w.clone()
        torch.distributed.broadcast(ref, 0)
        assert torch.equal(w, ref), "ranks diverged"
        torch.distributed.destroy_process_group()
    if rank == 0:
        print("ok")


if __name__ == "__main__":
    main()
