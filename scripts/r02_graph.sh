#!/bin/bash
# graph replay of launch-bound batches: parity test, then cases/hour vs cases in flight with and without it
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/graph
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graph_replay or api_surface or config1" > gpurun_out/graph/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/graph/pytest.log
for g in graph nograph; do
  for c in 1 2 3 4 6; do
    if [ $g = nograph ]; then export LBM2D_NO_GRAPH=1; else unset LBM2D_NO_GRAPH; fi
    timeout 300 python bench.py --workload sweep --concurrency $c 2> gpurun_out/graph/${g}_c$c.err | tail -1 > gpurun_out/graph/${g}_c$c.json
    python -c "import json; d=json.load(open('gpurun_out/graph/${g}_c$c.json')); print('$g', $c, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
  done
done
unset LBM2D_NO_GRAPH
for w in cylinder sweep_case tube_bank; do
  for g in graph nograph; do
    if [ $g = nograph ]; then export LBM2D_NO_GRAPH=1; else unset LBM2D_NO_GRAPH; fi
    timeout 300 python bench.py --workload $w --quick --steps 2000 --windows 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $g', round(d['ms_per_step']*1000,3),'us/step', round(d['value']),'MLUPS e2e', round(d['e2e']['value']))"
  done
done
