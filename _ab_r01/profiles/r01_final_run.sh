set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --quick --steps 200 --warmup 20 > gpurun_out/final/plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --quick --steps 200 --warmup 20 > gpurun_out/final/ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 3 -o gpurun_out/final/prof_step -f python bench.py --quick --steps 200 --warmup 20 > gpurun_out/final/ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/final/pytest_gpu.log; cat gpurun_out/final/bench.json
