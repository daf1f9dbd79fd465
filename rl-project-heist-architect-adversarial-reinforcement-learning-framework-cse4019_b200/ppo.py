"""Batched host-side glue for the callers either side of the env hot path.

The policy / value networks stay PyTorch and are passed in by the caller (the reference's SolverNetwork /
ArchitectNetwork work unchanged: same forward signatures).  What is here is the plumbing the reference does
with Python lists and `.cpu().numpy()` hops, restated on device tensors:

* `architect_sample`      - ArchitectNetwork.generate_layout's sampling (networks.py:259-274, 320): softmax/T,
                            Categorical sample per cell, summed log-prob -- no device->host hop; the sampled asset
                            map goes straight into `BatchedHeistEnv.set_layout_from_asset_map`.
* `PackedRollout`         - SolverAgent's rollout lists (agents/solver.py:58-64, 94-104) as time-major [T, N]
                            tensors; states are kept PACKED (visibility bitmap + solver position, 84 B instead of
                            4 800 B for 20x20) and re-expanded per minibatch by `heist_expand_states`.
* `collect_rollout`       - the trainer's inner loop (training.py:515-533) for the whole batch with the policy in
                            the loop: select_action (solver.py:75-99) -> env.step -> store_transition.
* `ppo_update`            - SolverAgent.update (solver.py:112-217): GAE (CUDA scan), advantage normalisation,
                            clipped PPO over minibatches; gradients are all-reduced across ranks between backward()
                            and clip_grad_norm_ (solver.py:195-199).
"""
import torch
import torch.nn.functional as F

from . import dist as hdist
from .rollout import compute_gae, normalize_advantages


def architect_sample(placement_logits, temperature=1.0, generator=None):
    """placement_logits [N, K, R, C] (K = asset types + 1) -> (asset_map int8 [N,R,C], total_log_prob [N]).

    networks.py:255-274: logits / T, softmax over asset types per cell, one Categorical sample per cell,
    log-probs summed over cells (:320)."""
    n, k, r, c = placement_logits.shape
    logp = F.log_softmax(placement_logits / temperature, dim=1)              # [N,K,R,C]
    flat = logp.permute(0, 2, 3, 1).reshape(n * r * c, k)
    sampled = torch.multinomial(flat.exp(), 1, generator=generator).squeeze(1)  # [N*R*C]
    total = flat.gather(1, sampled[:, None]).view(n, r * c).sum(1)
    return sampled.view(n, r, c).to(torch.int8), total


def camera_params_tensor(cam_params):
    """The network's {"fov","speed","heading"} dict of [N,1] tensors (networks.py:232-236) -> [N,3] float32 in
    the order heist_decode_validate expects."""
    return torch.cat([cam_params["fov"].view(-1, 1), cam_params["speed"].view(-1, 1),
                      cam_params["heading"].view(-1, 1)], dim=1).float().contiguous()


class PackedRollout:
    """Time-major [T, N] transition store with packed states."""

    def __init__(self, env, T):
        self.env, self.T, self.N = env, T, env.num_envs
        dev = env.device
        f = dict(dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((T, self.N), **f)
        self.values = torch.zeros((T, self.N), **f)
        self.log_probs = torch.zeros((T, self.N), **f)
        self.actions = torch.zeros((T, self.N), dtype=torch.int64, device=dev)
        self.dones = torch.zeros((T, self.N), dtype=torch.uint8, device=dev)
        self.vis_bits = torch.zeros((T, self.N, env.R, env.W), dtype=torch.int32, device=dev)
        self.pos = torch.zeros((T, self.N), dtype=torch.int32, device=dev)
        self.env_idx = torch.arange(self.N, dtype=torch.int32, device=dev).repeat(T, 1)
        self.t = 0

    def record_state(self):
        """Packed copy of what the policy is about to see (states.append(state), solver.py:94)."""
        self.vis_bits[self.t].copy_(self.env.visibility_bits)
        self.pos[self.t].copy_(self.env.env_dyn[:, 0])

    def record_packed(self, vis_bits, pos):
        """Same, from copies taken elsewhere (GraphedTick)."""
        self.vis_bits[self.t].copy_(vis_bits)
        self.pos[self.t].copy_(pos)

    def record_action(self, actions, log_probs, values):
        self.actions[self.t] = actions
        self.log_probs[self.t] = log_probs
        self.values[self.t] = values

    def record_outcome(self, rewards, dones):
        """store_transition (solver.py:101-104)."""
        self.rewards[self.t] = rewards
        self.dones[self.t] = dones
        self.t += 1

    def clear(self):
        self.t = 0

    def states(self, flat_idx):
        """Dense [B,3,R,C] states of the transitions with flat index t*N + n."""
        vb = self.vis_bits.view(-1, self.env.R, self.env.W)[flat_idx]
        return self.env.expand_states(vb, self.pos.view(-1)[flat_idx], self.env_idx.view(-1)[flat_idx])


def _call_policy(policy, state, hidden):
    """SolverNetwork.forward(state, hidden) -> (logits, value, hidden) (networks.py:65-131); plain
    feed-forward modules returning (logits, value) are accepted too."""
    out = policy(state, hidden) if _takes_hidden(policy) else policy(state)
    if len(out) == 3:
        return out
    return out[0], out[1], None


def _takes_hidden(policy):
    import inspect
    try:
        return len(inspect.signature(policy.forward).parameters) >= 2
    except (TypeError, ValueError, AttributeError):
        return False


def _mask_hidden(hidden, keep):
    """solver.reset() drops the LSTM state at the start of every attempt (training.py:517)."""
    if hidden is None:
        return None
    if isinstance(hidden, (tuple, list)):
        return tuple(_mask_hidden(h, keep) for h in hidden)
    return hidden * keep.view(1, -1, 1).to(hidden.dtype)


def _sample_actions(logits, generator=None):
    """Categorical sample + log-prob of the sample (solver.py:87-99).  Gumbel-max instead of torch.multinomial: the same
    distribution from one uniform draw per logit, no host sync, safe inside CUDA-graph capture."""
    logp_all = F.log_softmax(logits.float(), dim=-1)
    u = torch.rand(logp_all.shape, device=logp_all.device, generator=generator).clamp_(1e-12, 1.0 - 1e-7)
    action = torch.argmax(logp_all - torch.log(-torch.log(u)), dim=-1)
    return action, logp_all.gather(1, action[:, None]).squeeze(1)


@torch.no_grad()
def collect_rollout(env, policy, buffer, T=None, generator=None, hidden=None, tick=None):
    """T ticks of every env with the policy in the loop (training.py:515-533 batched, auto-reset on done).
    Returns (state, hidden, stats) where state is the dense observation after the last tick.
    `tick`: a GraphedTick built for (env, policy) -- the whole tick then replays from one CUDA graph."""
    T = T or buffer.T
    ended = torch.zeros(3, dtype=torch.int64, device=env.device)  # vault, detected, timeout (training.py:535-540)
    if tick is not None:
        tick.begin(hidden)
        for _ in range(T):
            tick.replay()
            buffer.record_packed(tick.vis_before, tick.pos_before)
            buffer.record_action(tick.action, tick.log_prob, tick.value)
            buffer.record_outcome(tick.reward, tick.done)
            ended += tick.ended
        return tick.state, tick.hidden, {"vault": ended[0], "detected": ended[1], "timeout": ended[2]}
    state = env.observe()
    for _ in range(T):
        buffer.record_state()
        logits, value, hidden = _call_policy(policy, state, hidden)
        action, logp = _sample_actions(logits, generator)
        buffer.record_action(action, logp, value.view(-1))
        reward, done, status, state = env.step_observe(action.to(torch.int8), autoreset=True, state_out=state)
        buffer.record_outcome(reward, done)
        hidden = _mask_hidden(hidden, ~done)
        ended += torch.stack([(status == 2).sum(), (status == 1).sum(), (status == 3).sum()])
    return state, hidden, {"vault": ended[0], "detected": ended[1], "timeout": ended[2]}


class GraphedTick:
    """One policy-in-the-loop tick -- packed copy of the state the policy sees, policy forward, action sampling,
    heist_step_observe (tick + auto-reset + dense next state), LSTM-state reset for finished episodes, outcome
    counts -- captured ONCE into a CUDA graph and replayed per tick (training.py:523-529 for the whole batch).
    The library's single-tick path neither allocates nor synchronises, so its kernels are captured like torch's.
    Outputs live in static tensors: action, log_prob, value, reward, done, status, state, hidden, vis_before,
    pos_before, ended."""

    def __init__(self, env, policy, autoreset=True, amp_dtype=None, warmup=3):
        self.env, self.policy = env, policy
        N, dev = env.num_envs, env.device
        self.state = env.observe()
        self.hidden = None
        self._takes_hidden = _takes_hidden(policy)
        self.action = torch.zeros(N, dtype=torch.int64, device=dev)
        self.log_prob = torch.zeros(N, dtype=torch.float32, device=dev)
        self.value = torch.zeros(N, dtype=torch.float32, device=dev)
        self.vis_before = torch.zeros((N, env.R, env.W), dtype=torch.int32, device=dev)
        self.pos_before = torch.zeros(N, dtype=torch.int32, device=dev)
        self.ended = torch.zeros(3, dtype=torch.int64, device=dev)
        self._act8 = torch.zeros(N, dtype=torch.int8, device=dev)
        self.amp_dtype = amp_dtype

        def body():
            self.vis_before.copy_(env.visibility_bits)
            self.pos_before.copy_(env.env_dyn[:, 0])
            with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
                logits, value, hid = _call_policy(policy, self.state, self.hidden)
            a, lp = _sample_actions(logits)
            self.action.copy_(a); self.log_prob.copy_(lp); self.value.copy_(value.view(-1).float())
            self._act8.copy_(a)
            r, d, st, _ = env.step_observe(self._act8, autoreset=autoreset, state_out=self.state)
            if hid is not None:
                hid = _mask_hidden(hid, ~d)
                if self.hidden is None:
                    self.hidden = tuple(torch.zeros_like(h) for h in hid) if isinstance(hid, (tuple, list)) else torch.zeros_like(hid)
                for dst, src in zip(self.hidden if isinstance(self.hidden, tuple) else (self.hidden,),
                                    hid if isinstance(hid, (tuple, list)) else (hid,)):
                    dst.copy_(src)
            self.ended.copy_(torch.stack([(st == 2).sum(), (st == 1).sum(), (st == 3).sum()]))
            return r, d, st

        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):   # warm-up outside capture (cuDNN autotune, lazy allocations, static hidden)
                for _ in range(warmup):
                    self.reward, self.done, self.status = body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.reward, self.done, self.status = body()
        torch.cuda.synchronize(dev)

    def begin(self, hidden=None):
        """Start a rollout from the env's current state (the warm-up ticks above advanced the envs: callers reset
        or re-layout the env after construction)."""
        self.env.observe(out=self.state)
        if self.hidden is not None:
            for i, h in enumerate(self.hidden if isinstance(self.hidden, tuple) else (self.hidden,)):
                if hidden is None:
                    h.zero_()
                else:
                    h.copy_(hidden[i] if isinstance(hidden, (tuple, list)) else hidden)

    def replay(self):
        self.graph.replay()


def minibatch_plan(n, minibatch, epochs, group=None, device="cpu", generator=None):
    """Index tensors of the minibatches this rank runs: `epochs` permutations of its n transitions, each cut into
    the SAME number of slices on every rank -- ceil(max over ranks of n / minibatch) -- because every minibatch ends
    in a collective (the gradient all-reduce); ranks whose shards differ in size (dist.shard_range with
    total % world != 0) then cut slightly smaller slices instead of running fewer iterations and hanging NCCL."""
    n_max, n_min = n, n
    if hdist._world(group) > 1:
        cnt = torch.tensor([n, -n], dtype=torch.int64, device=device)
        hdist.allreduce_max(cnt, group)
        n_max, n_min = int(cnt[0].item()), -int(cnt[1].item())
    if n_max == 0:
        return []
    per_epoch = -(-n_max // minibatch)
    if n_min < per_epoch:
        raise ValueError(f"minibatch_plan: a rank holds {n_min} transitions but every rank must cut {per_epoch} minibatches")
    plan = []
    for _ in range(epochs):
        perm = torch.randperm(n, device=device, generator=generator)
        plan += [perm[(i * n) // per_epoch:((i + 1) * n) // per_epoch] for i in range(per_epoch)]
    return plan


def ppo_update(policy, optimizer, buffer, epochs=3, minibatch=4096, gamma=0.99, gae_lambda=0.95, clip_epsilon=0.2,
               value_coeff=0.5, entropy_coeff=0.05, max_grad_norm=0.5, group=None, generator=None, bucket=None,
               amp_dtype=None):
    """SolverAgent.update (solver.py:112-217) on the packed buffer; returns mean losses as device tensors.

    Multi-rank: every rank issues exactly the same collectives whatever its shard size -- the minibatch COUNT comes
    from the all-reduced maximum of the rank-local transition counts and each rank cuts its own permutation into that
    many slices; the advantage statistics are global (normalize_advantages(group)).  `bucket`: a dist.GradBucket
    over policy.parameters(); its all-reduce overlaps the next minibatch's state expansion."""
    t = buffer.t
    n = t * buffer.N
    world = hdist._world(group)
    plan = minibatch_plan(n, minibatch, epochs, group, buffer.rewards.device, generator)
    if not plan:
        return {}
    adv, ret = compute_gae(buffer.rewards[:t], buffer.values[:t], buffer.dones[:t], gamma, gae_lambda)
    if world > 1:
        adv = normalize_advantages(adv, group)
    elif adv.numel() > 1:
        adv = normalize_advantages(adv)
    adv, ret = adv.reshape(-1), ret.reshape(-1)
    actions, old_logp = buffer.actions[:t].reshape(-1), buffer.log_probs[:t].reshape(-1)
    sums = torch.zeros(3, device=adv.device)
    if bucket is not None:
        bucket.zero()
    states = buffer.states(plan[0])
    for it, idx in enumerate(plan):
        with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
            logits, values, _ = _call_policy(policy, states, None)   # feed-forward re-evaluation (solver.py:171-172)
        logp_all = F.log_softmax(logits.float(), dim=-1)
        new_logp = logp_all.gather(1, actions[idx][:, None]).squeeze(1)
        entropy = -(logp_all.exp() * logp_all).sum(-1).mean()
        ratio = torch.exp(new_logp - old_logp[idx])
        surr = torch.min(ratio * adv[idx], torch.clamp(ratio, 1 - clip_epsilon, 1 + clip_epsilon) * adv[idx])
        policy_loss = -surr.mean()
        value_loss = F.mse_loss(values.view(-1).float(), ret[idx])
        loss = policy_loss + value_coeff * value_loss - entropy_coeff * entropy
        if bucket is None:
            optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if bucket is not None:
            handle = bucket.allreduce_async()
            if it + 1 < len(plan):
                states = buffer.states(plan[it + 1])   # overlaps the all-reduce
            bucket.wait(handle)
        else:
            hdist.allreduce_gradients(policy.parameters(), group)
            if it + 1 < len(plan):
                states = buffer.states(plan[it + 1])
        torch.nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)
        optimizer.step()
        if bucket is not None:
            bucket.zero()
        sums += torch.stack([policy_loss.detach(), value_loss.detach(), entropy.detach()])
    buffer.clear()
    sums /= max(len(plan), 1)
    return {"solver_policy_loss": sums[0], "solver_value_loss": sums[1], "solver_entropy": sums[2], "updates": len(plan)}
