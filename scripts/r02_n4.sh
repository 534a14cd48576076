#!/bin/bash
# N-GPU pass: slab parity tests (peer-memory and NCCL halo paths, distributed export) + weak-scaling bench incl. e2e
cd "$GRAFT_REPO_ROOT"; N=${1:-4}; O=gpurun_out/n$N; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -k "monolithic or fallback" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --no-alt > $O/bench_$N.json 2> $O/bench_$N.err
echo "bench exit $?"; tail -3 $O/bench_$N.err
tail -1 $O/bench_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e_reference_writer_path'] and round(d['e2e_reference_writer_path']['value']), d['slab_parity']['result'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 scripts/e2e_probe.py 2>&1 | grep -v "^\[\|\*\*\*\|OMP" | tail -4
