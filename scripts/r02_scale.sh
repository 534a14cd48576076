#!/bin/bash
# the driver's scaling protocol: full bench.py at N = 1, 2, 4, 8 (K = 20, W = 5), one 8-GPU box
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/scale; mkdir -p $O
nvidia-smi -L > $O/smi.txt
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_1.json 2> $O/bench_1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_$n.json 2> $O/bench_$n.err; fi
  echo "bench $n exit $?"; tail -1 $O/bench_$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['windows']['per_rank_median_ms'], 'alt', d['alt_arith'] and round(d['alt_arith']['mlups']), 'e2e', round(d['e2e']['value']), d['e2e_reference_writer_path'] and round(d['e2e_reference_writer_path']['value']), d.get('slab_parity',{}).get('result'))"
done
