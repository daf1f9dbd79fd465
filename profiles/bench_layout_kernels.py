"""Timing of the Architect-side kernels (decode + set_layout + BFS + slot order) at the BASELINE sizes."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, heist_b200
from heist_b200 import synthetic
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for i in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); tot += s.elapsed_time(e)
    return tot / reps
for R, N, dens, budget in [(20, 4096, (0.03, 0.012, 0.006), 15), (20, 4096, (0.25, 0.25, 0.25), 15), (32, 65536, (0.03, 0.012, 0.006), 22),
                           (32, 65536, (0.30, 0.004, 0.002), 22), (64, 262144, (0.03, 0.012, 0.006), 22)]:
    env = heist_b200.BatchedHeistEnv(heist_b200.EnvironmentConfig(grid_rows=R, grid_cols=R), N)
    rng = np.random.default_rng(R)
    am = torch.from_numpy(synthetic.sample_asset_maps(rng, N, R, R, *dens)).cuda()
    cp = torch.from_numpy(synthetic.sample_cam_params(rng, N)).cuda()
    ms = timeit(lambda: env.set_layout_from_asset_map(am, cp, budget))
    valid = env.valid.float().mean().item()
    print(f"decode+set_layout+bfs {R}x{R} N={N} density={dens}: {ms*1e3:.0f} us -> {N/ms/1e3:.1f} M layouts/s, "
          f"{N*R*R/ms/1e6:.0f} GB/s of asset map, valid {valid:.2f}")
    ms = timeit(lambda: env.reset())
    print(f"   reset: {ms*1e3:.0f} us")
    env.close()
