#!/bin/bash
# round 2, call c (2 GPUs): persistent push kernels (correctness + sweep), pushed MLP
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
timeout 300 python tools/sweep_push.py 2 > $O/r2c_push.log 2>&1
timeout 300 python tools/bench_c4.py 2 > $O/r2c_c4.log 2>&1
echo done
