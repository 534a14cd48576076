#!/bin/bash
# A/B on one box: round-1 tree (_ab_r01) vs current tree, alternating
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/ab
for i in 1 2 3; do
  (cd _ab_r01 && python bench.py --quick --steps 1000 --warmup 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('r01 ', d['ms_per_step'])")
  python bench.py --quick --steps 1000 --warmup 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new ', d['ms_per_step'])"
done | tee gpurun_out/ab/ab.txt
