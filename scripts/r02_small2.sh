#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/small2
for w in cylinder sweep_case tube_bank; do
  timeout 60 python bench.py --workload $w --quick --steps 2000 --windows 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step']*1000,3),'us/step', round(d['value']),'MLUPS e2e', round(d['e2e']['value']))"
done | tee gpurun_out/small2/out.txt
