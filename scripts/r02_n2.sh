#!/bin/bash
# 2-GPU pass: slab parity tests (peer-memory and NCCL halo paths) + weak-scaling bench
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/n2; mkdir -p $O
N=${1:-2}
nvidia-smi -L > $O/smi.txt; nvidia-smi topo -m >> $O/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -15 $O/pytest.log
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --quick --steps 20 --warmup 5 > $O/bench_1.json 2> $O/bench_1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --quick --steps 20 --warmup 5 > $O/bench_$n.json 2> $O/bench_$n.err; fi
  echo "bench $n exit $?"; tail -1 $O/bench_$n.json | cut -c1-900; tail -3 $O/bench_$n.err
done
