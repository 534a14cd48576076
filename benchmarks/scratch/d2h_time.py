"""scratch: time of get_moments_numpy (full-frame D2H) at 8192x2048"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
from helpers import make_config
pkg = importlib.import_module("01-lbm-2d_b200")
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
cfg = make_config(8192, 2048, rho_in=1.01, nu=0.01, cs=0.1, warmup=50, sponge=(8, 16, 8, 8))
s = pkg.LBM2D_MRT_LES(cfg, mask_data=None); s.init(); s.run_step(10); s.synchronize()
for i in range(5):
    t = time.perf_counter(); m = s.get_moments_numpy(); dt = time.perf_counter() - t
    print(f"get_moments_numpy {m.nbytes/1e6:.0f} MB: {dt*1e3:.1f} ms = {m.nbytes/dt/1e9:.1f} GB/s", flush=True)
    if i % 2: del m
v = s.vel.to_numpy(); t = time.perf_counter(); v = s.vel.to_numpy(); print(f"vel {v.nbytes/1e6:.0f} MB {(time.perf_counter()-t)*1e3:.1f} ms")
