#!/bin/bash
# ncu --set full of the strict step kernel (after the plain run exits 0)
A=${1:-strict}
OUT=gpurun_out/ncu_$A
mkdir -p $OUT
cd "$GRAFT_REPO_ROOT"
timeout 300 python bench.py --quick --steps 200 --warmup 20 --arith $A > $OUT/plain.json 2> $OUT/plain.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 2 -o $OUT/prof_$A -f python bench.py --quick --steps 200 --warmup 20 --arith $A > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; cat $OUT/plain.json | cut -c1-200
