#!/bin/bash
# full GPU pass: the whole gpu test-suite, smoke, bench (both arms)
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/full; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nproc >> $O/smi.txt
timeout 3000 python -m pytest tests -m gpu -q --durations=15 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
tail -25 $O/pytest.log; tail -5 $O/smoke.log; cat $O/bench.json | cut -c1-1500; tail -3 $O/bench.err
