#!/usr/bin/env python
"""BASELINE config 5 in batched form: Architect layout generation -> Solver rollouts (policy in the loop) -> GAE ->
PPO update with the gradient all-reduce, one process per GPU.

    python examples/adversarial_ppo.py --iters 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/adversarial_ppo.py

The networks are plain PyTorch (out of scope of the B200 rebuild): with the reference on sys.path
(`--reference /path/to/reference`) its own SolverNetwork / ArchitectNetwork are used unchanged; otherwise
heist_b200.nets stand-ins with the same layer shapes and forward contracts.  The loop itself is heist_b200.loop.  Mirrors AdversarialTrainer._run_one_episode (training.py:418-600):
curriculum budget -> generate_layout -> curriculum filter -> set_layout -> invalid layouts get -1 ->
`solver_episodes` attempts per layout -> architect reward from the solve rate -> both agents update.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heist_b200  # noqa: E402
from heist_b200 import nets  # noqa: E402
from heist_b200.loop import AdversarialLoop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024, help="layouts (envs) per GPU")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ticks", type=int, default=64)
    ap.add_argument("--budget", type=int, default=15)
    ap.add_argument("--reference", default=None)
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph tick")
    ap.add_argument("--amp", action="store_true", help="bf16 autocast inside the two networks")
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
    else:   # stand-ins with the reference's layer shapes (550 150 / 407 464 parameters)
        solver, architect = nets.SolverNet().to(dev), nets.ArchitectNet().to(dev)
    for net in (solver, architect):  # same initial weights on every rank
        for p in net.parameters():
            if world > 1:
                torch.distributed.broadcast(p.data, 0)
    env = heist_b200.BatchedHeistEnv(cfg, a.envs, device=dev)
    loop = AdversarialLoop(env, solver, architect, ticks=a.ticks, budget=a.budget, graph_tick=not a.eager,
                           amp_dtype=torch.bfloat16 if a.amp else None)
    for it in range(a.iters):
        stats, ms = loop.iteration(temperature=max(0.5, 2.0 - 1.5 * it / max(a.iters, 1)))   # training.py:451
        cnt = heist_b200.dist.allreduce_sum(torch.stack([stats["valid"], stats["vault"], stats["detected"], stats["timeout"]]))
        if rank == 0:
            steps = a.ticks * a.envs * world
            print(f"iter {it}: {steps} env-steps on {world} GPU(s) in {ms['total']:.1f} ms (layout {ms['architect_layout']:.1f}, "
                  f"rollout {ms['rollout']:.1f}, ppo {ms['ppo_update']:.1f}, architect {ms['architect_update']:.1f}) | valid layouts "
                  f"{int(cnt[0])} | episodes vault/detected/timeout {int(cnt[1])}/{int(cnt[2])}/{int(cnt[3])} | solver loss "
                  f"pi {float(stats['solver_policy_loss']):.4f} v {float(stats['solver_value_loss']):.4f} H {float(stats['solver_entropy']):.3f} | "
                  f"architect reward {float(stats['architect_reward']):.3f}", flush=True)
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
