#!/bin/bash
# 8-GPU pass: slab parity tests (world 2 and 4) + weak-scaling bench at N = 1, 2, 4, 8 (driver's K = 20, W = 5)
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/n8; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 600 python bench.py --quick --steps 20 --warmup 5 > $O/bench_1.json 2> $O/bench_1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --quick --steps 20 --warmup 5 > $O/bench_$n.json 2> $O/bench_$n.err; fi
  echo "bench $n exit $?"; tail -1 $O/bench_$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['windows']['per_rank_median_ms'], d['e2e']['value'], d['config']['parallelism'][:60])"
done
