import os, sys; sys.path.insert(0, os.getcwd())
import numpy as np, torch, heist_b200
from heist_b200 import synthetic
def timeit(fn, reps=20):
    flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot=0
    for i in range(reps):
        flush.fill_(i&255); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); tot+=s.elapsed_time(e)
    return tot/reps
for T,N in [(200,4096),(200,65536),(200,262144)]:
    r=torch.randn(T,N,device='cuda'); v=torch.randn(T,N,device='cuda'); d=(torch.rand(T,N,device='cuda')<0.02).to(torch.uint8)
    ms=timeit(lambda: heist_b200.compute_gae(r,v,d)); print(f"gae T={T} N={N}: {ms*1e3:.1f} us, {17*T*N/ms/1e6:.0f} GB/s ({17*T*N/ms/1e6/6554.2:.3f} of HBM peak)")
for R,N in [(20,4096),(20,65536),(32,65536),(64,65536)]:
    env=heist_b200.BatchedHeistEnv(heist_b200.EnvironmentConfig(grid_rows=R,grid_cols=R),N)
    st=torch.empty(N,3,R,R,device='cuda'); ms=timeit(lambda: env.observe(out=st))
    b=N*(12*R*R+R*R+4*R*((R+31)//32)+4); print(f"observe {R}x{R} N={N}: {ms*1e3:.1f} us, {b/ms/1e6:.0f} GB/s ({b/ms/1e6/6554.2:.3f} of HBM peak)")
    env.close()
