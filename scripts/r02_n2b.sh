#!/bin/bash
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/n2b; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-alt > $O/bench_2.json 2> $O/bench_2.err
echo "bench exit $?"; tail -2 $O/bench_2.err
tail -1 $O/bench_2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e_reference_writer_path'] and round(d['e2e_reference_writer_path']['value']), d['slab_parity']['result'])"
timeout 600 python -m pytest tests/test_gpu_slab.py -x -q -k "monolithic and auto and 2" 2>&1 | tail -2
