#!/bin/bash
# cases/hour vs cases in flight (bench.py --workload sweep: 32 cases per GPU), two rounds; host profile of one case
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/sweep3
timeout 300 python -m pytest tests/test_gpu_export.py -x -q -m gpu -k "fresh_pinned" 2>&1 | tail -2
for r in 1 2; do
  for c in 2 3 4 5; do
    timeout 300 python bench.py --workload sweep --concurrency $c 2> gpurun_out/sweep3/c${c}_$r.err | tail -1 > gpurun_out/sweep3/c${c}_$r.json
    python -c "import json; d=json.load(open('gpurun_out/sweep3/c${c}_$r.json')); print($c, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
  done
done
timeout 300 python scripts/sweep_profile.py 6 2>&1 | tail -36
