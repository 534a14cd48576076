#!/bin/bash
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/small; mkdir -p $O
for g in cylinder sweep_case; do
 for a in strict fast; do
  for pm in 2500 0; do
    LBM2D_PERSIST_MAX_CTAS=$pm timeout 300 python bench.py --workload $g --arith $a --quick --steps 2000 --windows 5 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g $a persist_max=$pm us/step', round(d['ms_per_step']*1000,3), 'MLUPS', round(d['value']), 'launches', d['gpu_launches'])"
  done
 done
done
