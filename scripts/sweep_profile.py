"""Where does a configs[4] sweep case spend its host time?  python scripts/sweep_profile.py [n_cases]"""
import cProfile, importlib, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from benchmarks import workloads as W
batch = importlib.import_module("01-lbm-2d_b200.batch")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cases = [(f"p{s}", *W.sweep_case(s)) for s in range(n)]
batch._gpu_runner(*cases[0], "/tmp/swp", 0, None, False)   # warm-up: context, module load
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for name, cfg, mask in cases[1:]:
    batch._gpu_runner(name, cfg, mask, "/tmp/swp", 0, None, False)
pr.disable()
dt = time.perf_counter() - t0
print(f"{(n - 1)} cases, {dt / (n - 1) * 1e3:.1f} ms per case")
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(28)
print("\n".join(l[:150] for l in out.getvalue().splitlines()[4:]))
