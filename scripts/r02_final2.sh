#!/bin/bash
# GPU pass after the prologue trims: the whole gpu test-suite, smoke, bench
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/final2; mkdir -p $O
timeout 3000 python -m pytest tests -m gpu -q --durations=5 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
tail -12 $O/pytest.log; tail -3 $O/smoke.log; tail -2 $O/bench.err
tail -1 $O/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], 'alt', round(d['alt_arith']['mlups']), 'e2e', round(d['e2e']['value']), round(d['e2e_reference_writer_path']['value']), d['gpu_launches'])"
