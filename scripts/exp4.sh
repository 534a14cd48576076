#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { tag=$1; shift; env "$@" python bench.py --quick --steps 1000 --windows 3 --warmup 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['ms_per_step']*1000,2))"; }
L=$PWD/01-lbm-2d_b200/lib
for i in 1 2; do
  run plain LBM2D_LIB=$L/exp_cur.so
  run peer_kernel LBM2D_LIB=$L/exp_cur.so LBM2D_FORCE_PEER_KERNEL=1
  run peer_kernel_expA LBM2D_LIB=$L/exp_expA.so LBM2D_FORCE_PEER_KERNEL=1
  run peer_kernel_noearly LBM2D_LIB=$L/exp_cur.so LBM2D_FORCE_PEER_KERNEL=1 LBM2D_EARLY_CTAS=0
  run plain_noearly LBM2D_LIB=$L/exp_cur.so LBM2D_EARLY_CTAS=0
done
