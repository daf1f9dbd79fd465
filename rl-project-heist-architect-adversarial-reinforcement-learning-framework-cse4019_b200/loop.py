"""BASELINE config 5 in batched form: the full adversarial loop around the env hot path, one process per GPU.

Mirrors AdversarialTrainer._run_one_episode (training.py:418-600) for a whole batch of layouts per rank:
  curriculum budget -> Architect forward, sample, decode + BFS-validate on device (networks.py:241-322,
  training.py:456-470) -> reset -> Solver rollouts with the policy in the loop (training.py:515-533), auto-reset ->
  GAE + clipped PPO over minibatches of re-expanded packed states (solver.py:112-217) with the gradient all-reduce
  between backward() and clip_grad_norm_ (solver.py:195-199) -> architect reward from the per-layout solve rates
  (rewards.py:43-73) and the Architect's value-only update (architect.py:105-141: the policy term is detached in the
  reference) with its own all-reduce (architect.py:138-141).

The networks are plain PyTorch modules passed in by the caller (the reference's own, or heist_b200.nets stand-ins of
the same shapes); everything between them is the library.  `iteration()` returns device-timed phase durations so
that bench.py can say what limits the loop.
"""
import torch
import torch.nn.functional as F

from . import dist as hdist
from . import nets as hnets
from . import ppo


class AdversarialLoop:
    def __init__(self, env, solver, architect, ticks=64, budget=15, solver_lr=1e-3, architect_lr=3e-4, epochs=3,
                 minibatch=8192, group=None, graph_tick=True, amp_dtype=None, allow_cameras=True, allow_guards=True):
        self.env, self.solver, self.architect = env, solver, architect
        self.ticks, self.budget, self.epochs, self.minibatch, self.group = ticks, budget, epochs, minibatch, group
        self.allow_cameras, self.allow_guards, self.amp_dtype = allow_cameras, allow_guards, amp_dtype
        cfg = env.config
        self.opt_s = torch.optim.Adam(solver.parameters(), lr=solver_lr)        # training.py:143, solver.py:52
        self.opt_a = torch.optim.Adam(architect.parameters(), lr=architect_lr)  # architect.py:41
        self.bucket_s = hdist.GradBucket(solver.parameters(), group)
        self.bucket_a = hdist.GradBucket(architect.parameters(), group)
        self.buf = ppo.PackedRollout(env, ticks)
        self.grid_in = hnets.empty_grid_input(env.num_envs, env.R, env.C, cfg.start_pos, cfg.vault_pos, env.device)
        self.tick = None
        self.graphed = False
        if graph_tick and not env.cache_warning:   # Architect-decoded layouts are always covered by the visibility tables
            env.set_mode(env.MODE_TABLES)
        if graph_tick:
            try:   # policy forward + sampling + env tick + dense state in ONE CUDA graph (ppo.GraphedTick)
                self.tick = ppo.GraphedTick(env, solver, autoreset=True, amp_dtype=amp_dtype)
                self.graphed = True
            except Exception as e:   # e.g. a policy module that cannot be captured: eager ticks, same results
                self.graph_error = f"{type(e).__name__}: {e}"
                self.tick = None
        self.iters = 0

    def _events(self, n):
        return [torch.cuda.Event(enable_timing=True) for _ in range(n)]

    def iteration(self, temperature=1.0):
        """One layout batch: returns (stats dict of device tensors / floats, phase milliseconds dict)."""
        env = self.env
        ev = self._events(5)
        ev[0].record()
        # --- Architect: forward, sample, decode, validate -- no device->host hop ---
        logits, a_value, cam = self.architect(self.grid_in)
        asset_map, _ = ppo.architect_sample(logits.detach(), temperature)
        valid = env.set_layout_from_asset_map(asset_map, ppo.camera_params_tensor(cam).detach(), self.budget,
                                              self.allow_cameras, self.allow_guards)
        env.reset()
        ev[1].record()
        # --- Solver rollouts, policy in the loop ---
        _, _, stats = ppo.collect_rollout(env, self.solver, self.buf, tick=self.tick)
        ev[2].record()
        # --- GAE + PPO with the all-reduce overlapped with the next minibatch's state expansion ---
        m = ppo.ppo_update(self.solver, self.opt_s, self.buf, epochs=self.epochs, minibatch=self.minibatch, group=self.group,
                           bucket=self.bucket_s, amp_dtype=self.amp_dtype)
        ev[3].record()
        # --- Architect: reward from solve rates, value-only update ---
        a_rew, solve_rate = env.architect_reward()
        a_loss = F.mse_loss(a_value.view(-1), a_rew.float())
        self.bucket_a.zero()
        a_loss.backward()
        self.bucket_a.allreduce()
        torch.nn.utils.clip_grad_norm_(self.architect.parameters(), 0.5)
        self.opt_a.step()
        ev[4].record()
        torch.cuda.synchronize(env.device)
        ms = {"architect_layout": ev[0].elapsed_time(ev[1]), "rollout": ev[1].elapsed_time(ev[2]),
              "ppo_update": ev[2].elapsed_time(ev[3]), "architect_update": ev[3].elapsed_time(ev[4]),
              "total": ev[0].elapsed_time(ev[4])}
        self.iters += 1
        out = {"valid": valid.sum(), "vault": stats["vault"], "detected": stats["detected"], "timeout": stats["timeout"],
               "architect_reward": a_rew.mean(), "solve_rate": solve_rate.mean(), "updates": m.get("updates", 0), **m}
        return out, ms

    def time_allreduce(self, reps=20):
        """Device-timed microseconds per gradient all-reduce of the two buckets (0.0 for a single process)."""
        if hdist._world(self.group) == 1:
            return {"solver_us": 0.0, "architect_us": 0.0, "solver_bytes": self.bucket_s.nbytes, "architect_bytes": self.bucket_a.nbytes}
        res = {}
        for name, b in (("solver", self.bucket_s), ("architect", self.bucket_a)):
            for _ in range(3):
                b.allreduce()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                b.allreduce()
            e.record()
            torch.cuda.synchronize()
            res[f"{name}_us"] = 1e3 * s.elapsed_time(e) / reps
            res[f"{name}_bytes"] = b.nbytes
            b.zero()
        return res
