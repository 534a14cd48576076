"""scratch: early start on / off on small grids (us per step)"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(sys.path[0], "tests"))
from helpers import make_config
pkg = importlib.import_module("01-lbm-2d_b200")
for nx, ny in [(512, 128), (1024, 256), (1024, 512), (1536, 512), (2048, 512), (1024, 1024), (4096, 256), (4096, 512), (2048, 1024)]:
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.01, cs=0.1, warmup=50, sponge=(8, 16, 8, 8))
    mask = np.zeros((nx, ny), bool); mask[nx // 4:nx // 4 + 20, ny // 2 - 10:ny // 2 + 10] = True
    row = []
    for early in ("0", "1500"):
        os.environ["LBM2D_EARLY_CTAS"] = early
        os.environ["LBM2D_EARLY_MIN_CTAS"] = "0"
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
        s.init(); s.run_step(500); s.synchronize()
        st = torch.cuda.ExternalStream(s.device_view().stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record(st); s.run_step(4000); e1.record(st); s.synchronize(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4000 * 1000)
        row.append(best); s.close()
    gx = (ny + 255) // 256
    print(f"{nx}x{ny} ctas={(nx-2)*gx}: off {row[0]:.2f} us  on {row[1]:.2f} us  ({row[0]/row[1]:.3f}x)", flush=True)
