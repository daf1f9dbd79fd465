"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, host-side
packing logic, and the multi-process (gloo, world_size 2) plumbing.  No kernel is launched here."""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import heist_b200
from heist_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "heist_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(heist_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    names = header_functions()
    assert len(names) >= 14 and "heist_step_many" in names
    lib = _ffi.load()
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_ffi.EXPORTS) == names
    assert lib.heist_abi_version() == 2


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_ffi.HeistParams) == 12 * 4 + 3 * 8
    assert C.sizeof(_ffi.HeistLayoutArrays) == 13 * 8
    assert C.sizeof(_ffi.HeistStateView) == 14 * 8


def test_argument_errors_do_not_need_a_gpu():
    lib = _ffi.load()
    assert lib.heist_create(None, 1, 0, None) == -1
    assert b"null" in lib.heist_last_error()
    assert lib.heist_step(None, None, None, None, None, None, None) == -1
    assert lib.heist_gae(None, None, None, 1, 1, 0.99, 0.95, None, None, 0, None) == -1


def test_product_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        heist_b200.BatchedHeistEnv()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        heist_b200.compute_gae(torch.zeros(4), torch.zeros(4), torch.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rl-project-heist-architect-adversarial-reinforcement-learning-framework-cse4019_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), f


def test_guard_heading_table_matches_reference_formula():
    path = [(7, 2), (7, 3), (7, 4), (7, 5)]
    assert heist_b200.guard_heading_table(path, 1) == [0.0, 0.0, 0.0, 180.0]
    t = heist_b200.guard_heading_table([(3, 3), (3, 3), (5, 4)], 1)
    assert math.isnan(t[0]) and t[1] == math.degrees(math.atan2(-2, 1)) % 360.0
    assert all(math.isnan(x) for x in heist_b200.guard_heading_table([(1, 1)], 1))
    up = heist_b200.guard_heading_table([(5, 5), (4, 5)], 1)
    assert up == [90.0, 270.0]


def test_pack_layouts_shapes_and_capacity():
    env = heist_b200.BatchedHeistEnv.__new__(heist_b200.BatchedHeistEnv)
    env.num_envs, env.R, env.C = 2, 10, 10
    env.max_walls, env.max_cams, env.max_guards, env.max_path = 4, 2, 1, 8
    lay = ([(3, 3), (3, 4)], [{"row": 5, "col": 5}], [{"patrol_path": [(7, 2), (7, 3)]}])
    a = env.pack_layouts([lay, ([], [], [])])
    assert a["n_walls"].tolist() == [2, 0] and a["wall_rc"][0, 1].tolist() == [3, 4]
    assert a["cam_f"][0, 0].tolist() == [60.0, 0.0, 15.0] and a["cam_range"][0, 0] == 6
    assert a["guard_len"][0, 0] == 2 and a["guard_head"][0, 0, :2].tolist() == [0.0, 180.0]
    assert a["guard_speed"][0, 0] == 1 and a["guard_range"][0, 0] == 4 and a["guard_fov"][0, 0] == 90.0
    with pytest.raises(ValueError):
        env.pack_layouts([([(1, 1)] * 5, [], []), ([], [], [])])
    with pytest.raises(ValueError):
        env.pack_layouts([([], [], [{"patrol_path": [(10, 10)]}]), ([], [], [])])


def test_shard_range_partitions():
    for total, world in [(4096, 8), (262144, 4), (10, 3), (7, 8)]:
        spans = [heist_b200.dist.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_synthetic_workloads_are_seeded():
    a = heist_b200.synthetic.sample_asset_maps(np.random.default_rng(1), 4, 20, 20)
    b = heist_b200.synthetic.sample_asset_maps(np.random.default_rng(1), 4, 20, 20)
    assert np.array_equal(a, b) and a[:, 0].sum() == 0 and a[:, :, -1].sum() == 0
    e = heist_b200.synthetic.sample_asset_maps_exact(np.random.default_rng(1), 3, 32, 32, 1, 4, 2)
    assert [(e == k).sum() for k in (1, 2, 3)] == [3, 12, 6]


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import heist_b200
rank, world, _ = heist_b200.dist.init_from_env("gloo")
assert world == 2
lo, hi = heist_b200.dist.shard_range(10, rank, world)
torch.manual_seed(0)
net = torch.nn.Linear(4, 3)
x = torch.arange(8, dtype=torch.float32).view(2, 4) + rank
net(x).sum().backward()
local = [p.grad.clone() for p in net.parameters()]
heist_b200.dist.allreduce_gradients(net.parameters())
gathered = [torch.zeros_like(local[0]) for _ in range(world)]
dist.all_gather(gathered, local[0])
assert torch.allclose(net.weight.grad, sum(gathered) / world)
# global advantage statistics == single-process statistics of the concatenation
g = torch.Generator().manual_seed(7)
full = torch.randn(10, generator=g)
mine = heist_b200.normalize_advantages(full[lo:hi], group=dist.group.WORLD)
ref = (full - full.mean()) / (full.std() + 1e-8)
assert torch.allclose(mine, ref[lo:hi], rtol=1e-5, atol=1e-6)
cnt = heist_b200.dist.allreduce_sum(torch.tensor([hi - lo], dtype=torch.int64))
assert cnt.item() == 10
mx = heist_b200.dist.allreduce_max(torch.tensor([float(rank)]))
assert mx.item() == 1.0
# persistent flat gradient bucket: .grad are views, one all-reduce, no copy-back
net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
for p in net2.parameters():
    dist.broadcast(p.data, 0)
bucket = heist_b200.dist.GradBucket(net2.parameters())
for it in range(2):
    net2(x * (it + 1)).sum().backward()
    mine = bucket.flat.clone()
    bucket.wait(bucket.allreduce_async())
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert torch.allclose(bucket.flat, sum(both) / world)
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() and p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.nbytes for p in net2.parameters())
    assert torch.equal(torch.cat([p.grad.reshape(-1) for p in net2.parameters()]), bucket.flat)
    bucket.zero()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in net2.parameters())
# unequal shards (513 vs 512 envs x 64 ticks, minibatch 8192): both ranks must cut the same number of minibatches
from heist_b200 import ppo
n_local = (513 if rank == 0 else 512) * 64
plan = ppo.minibatch_plan(n_local, 8192, 3, dist.group.WORLD)
counts = torch.tensor([len(plan)]); both = [torch.zeros_like(counts) for _ in range(world)]
dist.all_gather(both, counts)
assert both[0].item() == both[1].item() == 15, both
assert sum(len(i) for i in plan) == 3 * n_local and all(len(i) > 0 for i in plan)
assert sorted(torch.cat(plan[:5]).tolist()) == list(range(n_local))
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_stand_in_networks_have_the_reference_shapes():
    """networks.py:13-131 / :134-239 -> 550 150 and 407 464 parameters; forward contracts."""
    from heist_b200 import nets
    s, a = nets.SolverNet(), nets.ArchitectNet()
    assert sum(p.numel() for p in s.parameters()) == 550150
    assert sum(p.numel() for p in a.parameters()) == 407464
    logits, value, hidden = s(torch.zeros(2, 3, 20, 20))
    assert logits.shape == (2, 5) and value.shape == (2, 1) and hidden[0].shape == (1, 2, 128)
    pl, v, prm = a(nets.empty_grid_input(2, 20, 20, (1, 1), (18, 18), "cpu"))
    assert pl.shape == (2, 4, 20, 20) and v.shape == (2, 1) and sorted(prm) == ["fov", "heading", "speed"]
    assert 30 <= float(prm["fov"].min()) and float(prm["fov"].max()) <= 120


def test_host_libm_small_argument_identities():
    """heist_common.cuh ray_dir takes cos(x) = 1.0 and sin(x) = x for |x| <= radians(1e-11) (the window around
    0 degrees cannot be tabulated): pin that the platform libm the reference calls (math.cos / math.sin) agrees."""
    import math
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-1, 1, 200000) * 10.0 ** rng.uniform(-320, -12.7, 200000), [0.0, -0.0, 5e-324, 1.7453292519943295e-13]])
    assert all(math.cos(x) == 1.0 and math.sin(x) == x for x in xs.tolist())


def test_ppo_host_helpers_on_cpu():
    from heist_b200 import ppo
    logits = torch.full((2, 4, 5, 6), -1e4)
    want = torch.randint(0, 4, (2, 5, 6))
    logits.scatter_(1, want[:, None], 1e4)
    am, logp = ppo.architect_sample(logits, temperature=0.7)
    assert am.dtype == torch.int8 and torch.equal(am.long(), want) and logp.abs().max() < 1e-3
    h = (torch.ones(1, 3, 4), torch.ones(1, 3, 4))
    kept = ppo._mask_hidden(h, torch.tensor([True, False, True]))
    assert kept[0][0, 1].sum() == 0 and kept[1][0, 0].sum() == 4
    cp = ppo.camera_params_tensor({"fov": torch.tensor([[60.0]]), "speed": torch.tensor([[15.0]]),
                                   "heading": torch.tensor([[90.0]])})
    assert cp.tolist() == [[60.0, 15.0, 90.0]]
