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


@torch.no_grad()
def collect_rollout(env, policy, buffer, T=None, generator=None, hidden=None):
    """T ticks of every env with the policy in the loop (training.py:515-533 batched, auto-reset on done).
    Returns (state, hidden, stats) where state is the dense observation after the last tick."""
    T = T or buffer.T
    state = env.observe()
    ended = torch.zeros(3, dtype=torch.int64, device=env.device)  # vault, detected, timeout (training.py:535-540)
    for _ in range(T):
        buffer.record_state()
        logits, value, hidden = _call_policy(policy, state, hidden)
        probs = F.softmax(logits, dim=-1)
        action = torch.multinomial(probs, 1, generator=generator).squeeze(1)
        logp = torch.log(probs.gather(1, action[:, None]).squeeze(1))
        buffer.record_action(action, logp, value.view(-1))
        reward, done, status, state = env.step_observe(action.to(torch.int8), autoreset=True, state_out=state)
        buffer.record_outcome(reward, done)
        hidden = _mask_hidden(hidden, ~done)
        ended += torch.stack([(status == 2).sum(), (status == 1).sum(), (status == 3).sum()])
    return state, hidden, {"vault": ended[0], "detected": ended[1], "timeout": ended[2]}


def ppo_update(policy, optimizer, buffer, epochs=3, minibatch=4096, gamma=0.99, gae_lambda=0.95, clip_epsilon=0.2,
               value_coeff=0.5, entropy_coeff=0.05, max_grad_norm=0.5, group=None, generator=None):
    """SolverAgent.update (solver.py:112-217) on the packed buffer; returns mean losses as device tensors."""
    t = buffer.t
    if t == 0:
        return {}
    adv, ret = compute_gae(buffer.rewards[:t], buffer.values[:t], buffer.dones[:t], gamma, gae_lambda)
    adv = normalize_advantages(adv, group) if adv.numel() > 1 else adv
    n = t * buffer.N
    adv, ret = adv.reshape(-1), ret.reshape(-1)
    actions, old_logp = buffer.actions[:t].reshape(-1), buffer.log_probs[:t].reshape(-1)
    sums = torch.zeros(3, device=adv.device)
    updates = 0
    for _ in range(epochs):
        perm = torch.randperm(n, device=adv.device, generator=generator)
        for s in range(0, n, minibatch):
            idx = perm[s:s + minibatch]
            states = buffer.states(idx)
            logits, values, _ = _call_policy(policy, states, None)   # feed-forward re-evaluation (solver.py:171-172)
            logp_all = F.log_softmax(logits, dim=-1)
            new_logp = logp_all.gather(1, actions[idx][:, None]).squeeze(1)
            entropy = -(logp_all.exp() * logp_all).sum(-1).mean()
            ratio = torch.exp(new_logp - old_logp[idx])
            surr = torch.min(ratio * adv[idx], torch.clamp(ratio, 1 - clip_epsilon, 1 + clip_epsilon) * adv[idx])
            policy_loss = -surr.mean()
            value_loss = F.mse_loss(values.view(-1), ret[idx])
            loss = policy_loss + value_coeff * value_loss - entropy_coeff * entropy
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            hdist.allreduce_gradients(policy.parameters(), group)
            torch.nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)
            optimizer.step()
            sums += torch.stack([policy_loss.detach(), value_loss.detach(), entropy.detach()])
            updates += 1
    buffer.clear()
    sums /= max(updates, 1)
    return {"solver_policy_loss": sums[0], "solver_value_loss": sums[1], "solver_entropy": sums[2], "updates": updates}
