#!/bin/bash
# exp.sh ARITH VARIANT...: time lib/exp_VARIANT.so builds on one box (2 rounds, alternating)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/exp
arith=$1; shift
for i in 1 2; do
  for v in "$@"; do
    LBM2D_LIB=$PWD/01-lbm-2d_b200/lib/exp_$v.so python bench.py --quick --steps 1000 --windows 3 --warmup 100 --arith $arith 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$arith $v', round(d['ms_per_step']*1000,2))"
  done
done | tee -a gpurun_out/exp/exp.txt
