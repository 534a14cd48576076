#!/bin/bash
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/n2exp; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --quick --steps 200 --windows 5 --warmup 5 2>$O/$tag.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', d['ms_per_step'], d['windows']['per_rank_median_ms'], d['e2e']['value'])"; }
python bench.py --quick --steps 200 --windows 5 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1', d['ms_per_step'])"
run peer A=1



run peer A=1
CUDA_VISIBLE_DEVICES=1 python bench.py --quick --steps 200 --windows 5 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1_gpu1', d['ms_per_step'])"
runk() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --quick --steps 20 --warmup 5 2>$O/$tag.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag K=20', d['ms_per_step'], d['windows']['per_rank_median_ms'], d['e2e']['value'])"; }
runk peer A=1

python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n1 K=20', d['ms_per_step'])"
