#!/bin/bash
# Round 2 profile pass (one B200): plain run first (must exit 0), then the ncu launch list and one --set full capture of
# the dominant kernel, for the default (strict) and the optional fast arithmetic.  Outputs: gpurun_out/prof/.
cd "$GRAFT_REPO_ROOT"; O=gpurun_out/prof; mkdir -p $O
CMD="python bench.py --quick --steps 200 --windows 3 --warmup 20"
for A in strict fast; do
  timeout 300 $CMD --arith $A > $O/plain_$A.json 2> $O/plain_$A.err || { echo "plain $A failed"; exit 1; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$A.csv $CMD --arith $A > $O/ncu1_$A.log 2>&1; echo "launch list $A rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 3 -o $O/prof_$A -f $CMD --arith $A > $O/ncu2_$A.log 2>&1; echo "full $A rc=$?"
done
cut -c1-300 $O/plain_strict.json
