#!/bin/bash
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/n8q; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --quick --steps 20 --warmup 5 > $O/bench_8.json 2> $O/bench_8.err
echo "bench exit $?"; tail -1 $O/bench_8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['slab_parity']['result'], d['windows']['per_rank_median_ms'])"
