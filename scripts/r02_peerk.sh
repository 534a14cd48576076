#!/bin/bash
# one GPU: the PEER instantiation of the step kernel (slab code compiled in, no neighbour) against the plain one
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/peerk
for i in 1 2 3; do
  for v in plain peer; do
    if [ $v = peer ]; then export LBM2D_FORCE_PEER_KERNEL=1; else unset LBM2D_FORCE_PEER_KERNEL; fi
    python bench.py --quick --steps 1000 --windows 3 --warmup 100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step']*1000,2))"
  done
done | tee gpurun_out/peerk/out.txt
