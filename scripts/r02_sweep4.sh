#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/sweep4
for r in 1 2; do
  for c in 2 3 4; do
    timeout 300 python bench.py --workload sweep --concurrency $c 2> gpurun_out/sweep4/c${c}_$r.err | tail -1 > gpurun_out/sweep4/c${c}_$r.json
    python -c "import json; d=json.load(open('gpurun_out/sweep4/c${c}_$r.json')); print($c, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
  done
done
LBM2D_CASE_TIMING=1 timeout 300 python bench.py --workload sweep --concurrency 1 2>&1 | grep "^\[case" | head -12
LBM2D_CASE_TIMING=1 timeout 300 python bench.py --workload sweep --concurrency 3 2>&1 | grep "^\[case" | tail -8
df -h /tmp | tail -1; nproc; free -g | head -2
