#!/bin/bash
# cases/hour of the configs[4] sweep on ONE GPU as a function of the cases in flight
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/sweepc
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/sweepc/smi.txt
for c in 1 2 3 4 6 8; do
  timeout 300 python bench.py --workload sweep --concurrency $c 2> gpurun_out/sweepc/c$c.err | tail -1 > gpurun_out/sweepc/c$c.json
  python -c "import json; d=json.load(open('gpurun_out/sweepc/c$c.json')); print($c, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
done
for w in cylinder sweep_case tube_bank; do
  timeout 300 python bench.py --workload $w --quick --steps 2000 --windows 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step']*1000,3),'us/step', round(d['value']),'MLUPS e2e', round(d['e2e']['value']))"
done
