#!/bin/bash
# graph replay + h5lite on the GPU: parity tests, then cases/hour vs cases in flight (64 cases per run, two rounds)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/graph2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_loop.py tests/test_gpu_export.py tests/test_gpu_workloads.py -x -q -m gpu -k "graph_replay or reference or export or configs4 or configs1 or api_surface" > gpurun_out/graph2/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/graph2/pytest.log
for r in 1 2; do
  for c in 2 3 4 5; do
    timeout 300 python 01-lbm-2d_b200/batch.py --sweep 64 --no-resume --out /tmp/sw --concurrency $c 2> gpurun_out/graph2/c${c}_$r.err | tail -1 > gpurun_out/graph2/c${c}_$r.json
    python -c "import json; d=json.load(open('gpurun_out/graph2/c${c}_$r.json')); print($c, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
  done
done
python /tmp/x.py 2>/dev/null
ls -la /tmp/sw | head -5; python 01-lbm-2d_b200/h5lite.py /tmp/sw/sweep_00.h5 | cut -c1-150
