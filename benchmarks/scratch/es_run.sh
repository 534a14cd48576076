mkdir -p gpurun_out/es
exec > gpurun_out/es/sweep.txt 2>&1
run() { LBM2D_EARLY_CTAS=$1 timeout 300 python bench.py --quick $2 --steps $3 --warmup 100 2>gpurun_out/es/err.txt | tail -1 > gpurun_out/es/last.json; python -c "import sys,json; d=json.load(open('gpurun_out/es/last.json')); print('$2 early',$1,round(d['value']),round(d['roofline']['frac'],4), round(d['ms_per_step']*1000,2),'us')" || { tail -3 gpurun_out/es/err.txt; head -c 300 gpurun_out/es/last.json; }; }
for e in 0 750 1000 1500 2000 0 1500; do run $e "" 1000; done
for g in 16384x2048 4096x2048 2048x2048 8192x512 2048x8192; do for e in 0 1500; do run $e "--grid $g" 1000; done; done
for w in tube_bank cylinder; do for e in 0 1500; do run $e "--workload $w" 5000; done; done
