#!/bin/bash
# device memory pool + short GIL switch interval: parity subset, then cases/hour vs cases in flight (A/B against no pool)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/pool
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py tests/test_gpu_reference_loop.py tests/test_gpu_workloads.py tests/test_gpu_viz.py tests/test_gpu_bounce_back.py -x -q -m gpu -k "not 8192 and not urban and not early_start and not unsteady and not very_long and not selftest and not inline" > gpurun_out/pool/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pool/pytest.log
run() { # label conc
  timeout 300 python bench.py --workload sweep --concurrency $2 2> gpurun_out/pool/$1_c$2.err | tail -1 > gpurun_out/pool/$1_c$2.json
  python -c "import json; d=json.load(open('gpurun_out/pool/$1_c$2.json')); print('$1', $2, round(d['value']), round(d['mlups_aggregate']), round(d['wall_s'],2), d['success'])"
}
for r in a b; do
  for c in 1 2 3 4; do run pool$r $c; done
  LBM2D_NO_POOL=1 run nopool$r 3
done
LBM2D_CASE_TIMING=1 timeout 300 python bench.py --workload sweep --concurrency 1 2>&1 | grep "^\[case" | head -8
LBM2D_CASE_TIMING=1 timeout 300 python bench.py --workload sweep --concurrency 3 2>&1 | grep "^\[case" | tail -6
