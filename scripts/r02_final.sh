#!/bin/bash
# final GPU pass of round 2: the whole gpu test-suite, smoke, bench (both arms), launch list of the bench command
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/final; mkdir -p $O
nvidia-smi -L > $O/smi.txt; nproc >> $O/smi.txt
timeout 3000 python -m pytest tests -m gpu -q --durations=12 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_strict.csv python bench.py --quick --steps 200 --windows 3 --warmup 20 > $O/ncu1.log 2>&1; echo "launch list rc=$?"
timeout 300 python bench.py --workload sweep > $O/sweep.json 2> $O/sweep.err
tail -22 $O/pytest.log; tail -4 $O/smoke.log; tail -1 $O/bench.json | cut -c1-1200; tail -3 $O/bench.err; tail -1 $O/sweep.json | cut -c1-300
