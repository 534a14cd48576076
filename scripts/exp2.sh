#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/exp
run() { python bench.py --quick --steps 1000 --warmup 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1000,2))"; }
for i in 1 2; do
  (cd _ab_r01 && LBM2D_EARLY_CTAS=0 run r01_noearly)
  LBM2D_EARLY_CTAS=0 run new_noearly
  (cd _ab_r01 && LBM2D_NO_PDL=1 run r01_nopdl)
  LBM2D_NO_PDL=1 run new_nopdl
done | tee -a gpurun_out/exp/exp2.txt
