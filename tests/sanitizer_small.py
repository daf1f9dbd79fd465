import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, heist_b200
from heist_b200 import synthetic
for (R,C,N,T,counts) in [(20,20,96,40,None),(32,32,48,30,(2,4,2)),(64,64,24,20,(2,4,2)),(12,17,40,30,None)]:
    env = heist_b200.BatchedHeistEnv(heist_b200.EnvironmentConfig(grid_rows=R,grid_cols=C,max_steps=25), N)
    rng = np.random.default_rng(R)
    am = synthetic.sample_asset_maps(rng,N,R,C) if counts is None else synthetic.sample_asset_maps_exact(rng,N,R,C,*counts)
    cp = synthetic.sample_cam_params(rng,N, nice=(R==32))
    env.set_layout_from_asset_map(am,cp,22); env.reset()
    out = env.step_many(synthetic.sample_actions(rng,T,N), autoreset=True, want_vis=True)
    r,d,s = env.step(np.zeros(N,np.int8)); st = env.observe(); o = env.observation(); env.architect_reward()
    a,b = heist_b200.compute_gae(out["reward"], torch.zeros_like(out["reward"]), out["done"])
    env.check_errors(); torch.cuda.synchronize(); print(R,C,"ok", int(out["done"].sum()))
