#!/bin/bash
# round 2, first GPU pass: parity of the new packed strict kernel + strict / fast throughput
mkdir -p gpurun_out/r1
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L > gpurun_out/r1/smi.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or config1 or awkward or random_bc or very_long or nan or early_start_is_bit" > gpurun_out/r1/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1/pytest.log
for arith in strict fast; do
  timeout 300 python bench.py --quick --steps 1000 --warmup 100 --arith $arith > gpurun_out/r1/bench_$arith.json 2> gpurun_out/r1/bench_$arith.err
done
tail -3 gpurun_out/r1/pytest.log
cat gpurun_out/r1/bench_*.json
