#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/small2
for g in graph nograph graph nograph; do
  if [ $g = nograph ]; then export LBM2D_NO_GRAPH=1; else unset LBM2D_NO_GRAPH; fi
  timeout 40 python bench.py --workload cylinder --quick --steps 2000 --windows 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cylinder $g', round(d['ms_per_step']*1000,3),'us/step')"
done | tee gpurun_out/small2/ab.txt
