#!/bin/bash
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/e2e1; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 --no-alt > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
tail -1 $O/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), round(d['e2e_reference_writer_path']['value']))"
timeout 300 python bench.py --workload tube_bank --quick --steps 500 --windows 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tube_bank', round(d['ms_per_step']*1000,3),'us/step', round(d['value']),'MLUPS e2e', round(d['e2e']['value']))"
