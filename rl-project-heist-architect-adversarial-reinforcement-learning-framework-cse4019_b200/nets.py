"""Stand-in policy / value networks with the reference's layer shapes, for the full-loop measurements
(BASELINE config 5) when the reference tree is not on the box.

The networks are OUT OF SCOPE of the B200 rebuild (they stay plain PyTorch); the hot path
only needs *something* of the right shape between `observe` and `step`.  With the reference importable its own
`SolverNetwork` / `ArchitectNetwork` plug into `ppo.collect_rollout` / `ppo.ppo_update` unchanged (same forward
contracts); these two reproduce their shapes only -- layer widths from networks.py:37-63 and :159-193, parameter
counts 550 150 and 407 464 (checked by tests/test_host_cpu.py) -- not their initialisation.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _trunk(c_in):
    """3x3 same-padding conv stack c_in -> 32 -> 64 -> 64."""
    return nn.Sequential(nn.Conv2d(c_in, 32, 3, padding=1), nn.ReLU(), nn.Conv2d(32, 64, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(64, 64, 3, padding=1), nn.ReLU())


def _head(d_in, d_out):
    return nn.Sequential(nn.Linear(d_in, 128), nn.ReLU(), nn.Linear(128, d_out))


class SolverNet(nn.Module):
    """forward(state [B,3,R,C], hidden=None) -> (logits [B,5], value [B,1], hidden): conv trunk, 4x4 average pool,
    1024 -> 256, one LSTM cell step 256 -> 128 (the reference feeds sequences of length 1), two MLP heads."""

    def __init__(self, num_actions=5, hidden_dim=256, lstm_hidden=128):
        super().__init__()
        self.trunk = _trunk(3)
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.fc = nn.Linear(64 * 16, hidden_dim)
        self.lstm = nn.LSTM(hidden_dim, lstm_hidden, batch_first=True)
        self.pi = _head(lstm_hidden, num_actions)
        self.v = _head(lstm_hidden, 1)
        self.lstm_hidden = lstm_hidden

    def forward(self, state, hidden=None):
        x = F.relu(self.fc(self.pool(self.trunk(state)).flatten(1)))
        if hidden is None:
            z = x.new_zeros(1, x.shape[0], self.lstm_hidden)
            hidden = (z, z.clone())
        y, hidden = self.lstm(x.unsqueeze(1), hidden)
        y = y.squeeze(1)
        return self.pi(y), self.v(y), hidden


class ArchitectNet(nn.Module):
    """forward(grid [B,1,R,C]) -> (placement_logits [B,4,R,C], value [B,1], {"fov","speed","heading"} each [B,1])
    with the camera-parameter ranges of networks.py:232-236."""

    def __init__(self, hidden_dim=256):
        super().__init__()
        self.enc = _trunk(1)
        self.pool = nn.AdaptiveAvgPool2d(4)
        self.fc = nn.Linear(64 * 16, hidden_dim)
        self.dec = nn.Sequential(nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(), nn.Conv2d(64, 32, 3, padding=1), nn.ReLU(),
                                 nn.Conv2d(32, 4, 1))
        self.v = _head(hidden_dim, 1)
        self.cam = nn.Linear(hidden_dim, 3)   # fov, speed, heading (three 256 -> 1 heads side by side)

    def forward(self, grid):
        f = self.enc(grid)
        g = F.relu(self.fc(self.pool(f).flatten(1)))
        s = torch.sigmoid(self.cam(g))
        params = {"fov": s[:, 0:1] * 90 + 30, "speed": s[:, 1:2] * 30 + 5, "heading": s[:, 2:3] * 360}
        return self.dec(f), self.v(g), params


def empty_grid_input(n, rows, cols, start, vault, device):
    """[n,1,R,C] float32: what the Architect sees -- zeros with START / VAULT marked as tile code / 5
    (architect.py:66-71)."""
    g = torch.zeros((n, 1, rows, cols), device=device)
    g[:, 0, start[0], start[1]] = 2 / 5
    g[:, 0, vault[0], vault[1]] = 3 / 5
    return g
