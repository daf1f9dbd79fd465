"""Solver rollout buffer in time-major [T, N] layout with a CUDA GAE/returns scan.

Reference: SolverAgent's Python lists and _compute_gae (agents/solver.py:58-64, 94-104, 134-147,
228-244).  Column n of the buffer is env n's flat concatenation of episodes, which is exactly the
single flat buffer of the reference when N = 1.
"""
import ctypes as C

import torch

from . import _ffi


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def compute_gae(rewards, values, dones, gamma=0.99, gae_lambda=0.95):
    """rewards, values [T] or [T,N] float32, dones uint8/bool (CUDA) -> (advantages, returns) float32.

    Bit-identical to the reference's fp32 op order; the last row's next value is 0 (solver.py:235-236)."""
    if not rewards.is_cuda:
        raise RuntimeError("compute_gae: tensors must be on a CUDA device (no CPU fallback)")
    rew = rewards.contiguous().float()
    val = values.contiguous().float()
    dn = dones.contiguous().to(torch.uint8)
    T = rew.shape[0]
    n = 1 if rew.dim() == 1 else rew.shape[1]
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    lib = _ffi.load()
    stream = C.c_void_p(torch.cuda.current_stream(rew.device).cuda_stream)
    _ffi.check(lib.heist_gae(_ptr(rew), _ptr(val), _ptr(dn), T, n, float(gamma), float(gae_lambda), _ptr(adv),
                             _ptr(ret), rew.device.index or 0, stream), "heist_gae")
    return adv, ret


def normalize_advantages(adv, group=None):
    """(A - mean) / (std_unbiased + 1e-8) over all elements (solver.py:146-147); with a process group the
    statistics are global over ranks (sum, sum of squares, count all-reduced)."""
    if adv.numel() <= 1 and group is None:
        return adv
    if group is None:
        return (adv - adv.mean()) / (adv.std() + 1e-8)
    import torch.distributed as dist
    a64 = adv.double()
    stats = torch.stack([a64.sum(), (a64 * a64).sum(), torch.tensor(float(adv.numel()), device=adv.device,
                                                                       dtype=torch.float64)])
    dist.all_reduce(stats, group=group)
    n = stats[2]
    mean = stats[0] / n
    var = (stats[1] - n * mean * mean) / (n - 1)
    return ((adv - mean.float()) / (var.clamp_min(0).sqrt().float() + 1e-8))


class RolloutBuffer:
    """Preallocated [T, N] transition store (reward/value/logp f32, action i8, done u8) plus packed
    per-step visibility bitmaps, solver positions and ticks from which states can be re-expanded."""

    def __init__(self, T, num_envs, device, rows=None, words=None):
        self.T, self.N, self.device = T, num_envs, torch.device(device)
        f = dict(dtype=torch.float32, device=self.device)
        self.rewards = torch.zeros((T, num_envs), **f)
        self.values = torch.zeros((T, num_envs), **f)
        self.log_probs = torch.zeros((T, num_envs), **f)
        self.actions = torch.zeros((T, num_envs), dtype=torch.int8, device=self.device)
        self.dones = torch.zeros((T, num_envs), dtype=torch.uint8, device=self.device)
        self.vis_bits = (torch.zeros((T, num_envs, rows, words), dtype=torch.int32, device=self.device)
                         if rows else None)
        self.t = 0

    def add(self, actions, log_probs, values, rewards, dones, vis_bits=None):
        """select_action + store_transition for one tick of the whole batch (solver.py:94-104)."""
        t = self.t
        self.actions[t] = actions
        self.log_probs[t] = log_probs
        self.values[t] = values
        self.rewards[t] = rewards
        self.dones[t] = dones
        if vis_bits is not None and self.vis_bits is not None:
            self.vis_bits[t] = vis_bits
        self.t += 1

    def full(self):
        return self.t >= self.T

    def clear(self):
        self.t = 0

    def compute_returns(self, gamma=0.99, gae_lambda=0.95, normalize=True, group=None):
        t = self.t
        adv, ret = compute_gae(self.rewards[:t], self.values[:t], self.dones[:t], gamma, gae_lambda)
        if normalize:
            adv = normalize_advantages(adv, group)
        return adv, ret
