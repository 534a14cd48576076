#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/exp
bash scripts/exp.sh strict C2 D
LBM2D_LIB=$PWD/01-lbm-2d_b200/lib/exp_D.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounce_back.py -x -q -m gpu -k "blow_up or nan or golden or random or awkward or bounce or graph_replay" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blow_up or graph_replay" 2>&1 | tail -2
